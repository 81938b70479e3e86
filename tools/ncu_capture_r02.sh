#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE B200; every ncu run is preceded by the identical plain run exiting 0).
# `--set full` of every hot kernel at its BASELINE shape, one launch each, from the per-kernel drivers tools/bench_kernels.py
# (cfg2 shapes: 13 tcgen05 conv layers fwd / dgrad / wgrad, narrow layers, BatchNorm / pool passes, pseudo labels, CE+Dice,
# SGD+EMA; a second run with the loss kernels at cfg4 size) and tools/bench_mid.py (UNet-B warp-MMA kernels incl. the logits head).
# The drivers allocate only the operands of one layer at a time: ncu saves / restores device memory around every replay
# pass (3-4 s per launch even so), so profiling inside the full training step (tens of GB resident) is not practical.
# The .ncu-rep files (0.7 MB per launch) exceed what gpurun brings back: they are converted to the raw-page CSV on the box.
set -u
OUT=gpurun_out
export USTRUN_BENCH_ITERS=0
cap() { # name, kernel regex, command...
  local name=$1 regex=$2; shift; shift
  "$@" > $OUT/r02_plain_$name.log 2>&1 && timeout 330 ncu --set full --clock-control none -k "regex:$regex" -o $OUT/r02_ncu_$name -f "$@" > $OUT/r02_ncu_$name.log 2>&1
  echo "$name: rc=$?"
  ncu -i $OUT/r02_ncu_$name.ncu-rep --page raw --csv > $OUT/r02_ncu_$name.csv 2>/dev/null
  ls -la $OUT/r02_ncu_$name.ncu-rep $OUT/r02_ncu_$name.csv
  rm -f $OUT/r02_ncu_$name.ncu-rep
}
cap tc_cfg2 '^k_tc_' python tools/bench_kernels.py
cap hbm_cfg2 '^k_(bn_|maxpool|conv_first|conv_narrow|head1x1|wgrad_narrow|pseudo|ce_dice_pass|sgd_ema)' python tools/bench_kernels.py narrow
USTRUN_BENCH_STEP_SHAPE=32,288,4 cap step_cfg4 '^k_(pseudo|ce_dice_pass|sgd_ema)' python tools/bench_kernels.py narrow
cap mid_cfg2b '^k_(conv_mid|wgrad_mid)' python tools/bench_mid.py
unset USTRUN_BENCH_ITERS
python tools/bench_kernels.py > $OUT/r02_kernel_microbench_cfg2.txt 2>&1
python tools/bench_mid.py > $OUT/r02_kernel_microbench_cfg2b.txt 2>&1
du -sh $OUT
