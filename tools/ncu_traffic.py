"""Summarise DRAM traffic of the tcgen05 conv kernels from an `ncu --set full --csv --page raw` capture of
tools/bench_kernels.py (USTRUN_BENCH_ITERS=0: one launch per kernel, layers in table order: fwd, dgrad, wgrad).
Writes profiles/r01_ncu_tc_conv_traffic.json, which bench.py reports as roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/ncu_tc_full.csv"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYERS = [("inc.3", 64, 64, 1), ("down1.0", 64, 128, 2), ("down1.3", 128, 128, 2), ("down2.0", 128, 256, 4), ("down2.3", 256, 256, 4),
          ("down3.0", 256, 512, 8), ("down3.3", 512, 512, 8), ("down4.0", 512, 1024, 16), ("down4.3", 1024, 1024, 16),
          ("up1.0", 1024, 512, 8), ("up2.0", 512, 256, 4), ("up3.0", 256, 128, 2), ("up4.0", 128, 64, 1)]
B, H0 = 8, 384
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[i], [r for r in rows[i + 2:] if len(r) == len(rows[i])]
col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
unit = {n: rows[i + 1][col[n]] for n in col}
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
conv = [r for r in data if "k_tc_conv<" in r[col["Kernel Name"]]]
assert len(conv) == 2 * len(LAYERS), f"expected {2 * len(LAYERS)} k_tc_conv launches, got {len(conv)}"
out, tot_d, tot_a, tot_t, tot_f = [], 0.0, 0.0, 0.0, 0.0
for li, (name, cin, cout, d) in enumerate(LAYERS):
    H = H0 // d
    px = B * H * H
    alg = px * cin * 2 + px * cout * 2 + 9 * cin * cout * 2            # one bf16 read of the input, one write of the output, the weights
    flops = 2.0 * px * cin * cout * 9
    for which in (0, 1):                                               # fwd, dgrad (same algorithmic bytes, roles swapped)
        r = conv[2 * li + which]
        dram = to_bytes(r[col["dram__bytes_read.sum"]], unit["dram__bytes_read.sum"]) + to_bytes(r[col["dram__bytes_write.sum"]], unit["dram__bytes_write.sum"])
        t = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3}[unit["gpu__time_duration.sum"]]
        out.append({"layer": name, "pass": ("fwd", "dgrad")[which], "us": round(t * 1e6, 1), "dram_bytes": dram, "algorithmic_bytes": alg, "ratio": round(dram / alg, 2),
                    "tensor_active_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])})
        tot_d += dram; tot_a += alg; tot_t += t; tot_f += flops
res = {"source": os.path.basename(sys.argv[1]), "launches": len(out), "dram_bytes_per_launch": tot_d / len(out), "algorithmic_bytes_per_launch": tot_a / len(out),
       "ratio": tot_d / tot_a, "tflops_under_ncu": tot_f / tot_t / 1e12, "per_launch": out}
json.dump(res, open(os.path.join(ROOT, "profiles", "r01_ncu_tc_conv_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "per_launch"}))
for o in out: print(o)
