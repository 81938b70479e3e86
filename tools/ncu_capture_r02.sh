#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE B200; every ncu run is preceded by the identical plain run exiting 0).
# Captures `--set full` for the kernels VERDICT r01 asked for, on the final code, inside a real SSL step:
#   cfg2 (UNet-A 1x384x384, 8+8): one loss branch (forward + backward) of k_tc_conv, k_tc_wgrad_row, the BatchNorm / pool kernels,
#   the fused SGD+EMA and the one-launch weight packing; cfg4 (32+32 at 288x288, 4 classes): pseudo labels, CE+Dice pass 1 / 2.
# Windows (-s / -c) are placed with the per-step launch counts of the single-lane eager step (profiles/r02_launches_*.csv);
# tools/step_once.py --manifest records the algorithmic work and shape of every profiled call of the LAST step.
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
run() { # name regex skip count workload steps
  local name=$1 regex=$2 skip=$3 count=$4 wl=$5 steps=$6
  python tools/step_once.py --workload $wl --steps $steps > $OUT/ncu_plain_$name.log 2>&1 && \
  $NCU -k "regex:$regex" -s $skip -c $count -o $OUT/r02_ncu_$name -f python tools/step_once.py --workload $wl --steps $steps > $OUT/ncu_$name.log 2>&1
  echo "$name: rc=$?"
}
python tools/step_once.py --workload cfg2 --steps 3 --manifest $OUT/r02_manifest_cfg2.json > $OUT/ncu_plain_manifest_cfg2.log 2>&1
python tools/step_once.py --workload cfg4 --steps 2 --manifest $OUT/r02_manifest_cfg4.json > $OUT/ncu_plain_manifest_cfg4.log 2>&1
python tools/step_once.py --workload cfg2b --steps 3 --manifest $OUT/r02_manifest_cfg2b.json > $OUT/ncu_plain_manifest_cfg2b.log 2>&1
# k_tc_conv: 381 launches per step (5 no-grad forwards x 33, 4 branches x 54); window = the labelled branch, forward + backward
run tc_conv 'k_tc_conv' 894 54 cfg2 3
# row-mode weight gradients: 68 per step (17 per branch)
run tc_wgrad 'k_tc_wgrad' $((2*84)) 21 cfg2 3
# BatchNorm / pool passes: 322 per step (162 bn_act[_pool], 72 reduce, 72 apply, 16 maxpool_bwd); window = the labelled branch
run bn 'k_bn_act|k_bn_bwd_reduce|k_bn_bwd_apply|k_maxpool_bwd' $((2*322+72)) 58 cfg2 3
run opt 'k_sgd_ema|k_pack_multi' 3 3 cfg2 3
# loss kernels at cfg4 size: per step 1 pseudo-label + 4 x (pass1 + finalize + pass2)
run loss 'k_pseudo_label|k_ce_dice_pass' 9 9 cfg4 2
# UNet-B 16/32-channel kernels (warp-level MMA) incl. the logits head: cfg2b, first no-grad forward of step 2
run mid 'k_conv_mid_mma|k_wgrad_mid_mma|k_conv_first_mma' $((2*185)) 40 cfg2b 3
for n in tc_conv tc_wgrad bn opt loss mid; do
  ncu -i $OUT/r02_ncu_$n.ncu-rep --page raw --csv > $OUT/r02_ncu_$n.csv 2>/dev/null
done
ls -la $OUT/r02_ncu_*.csv $OUT/r02_ncu_*.ncu-rep
