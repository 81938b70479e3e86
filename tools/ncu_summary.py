"""Summarise an `ncu --set full` capture (raw page as CSV) of the step's kernels, optionally joined with the launch manifest
that tools/step_once.py --manifest wrote for the same step (algorithmic FLOPs / bytes and the layer shape of every call).

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > gpurun_out/X.csv
    python tools/ncu_summary.py gpurun_out/X.csv [--manifest gpurun_out/manifest.json --cls tc_conv --skip-calls N] [--json out.json]

Per launch: duration, tensor-pipe activity, DRAM bytes read + written (`traffic`), DRAM throughput %, registers, and -- with a
manifest -- achieved TFLOP/s or GB/s on the ALGORITHMIC work and traffic / algorithmic bytes.  Times under ncu are cold-cache
and serialised: the numbers to read are the percentages and the byte counts."""
import argparse
import csv
import json
import re
import sys

UNIT_T = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
UNIT_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
COLS = {"t": "gpu__time_duration.sum", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread", "grid": "launch__grid_size",
        "smem": "launch__shared_mem_per_block_dynamic", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "l2hit": "lts__t_sector_hit_rate.pct"}


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def load(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units = rows[i], rows[i + 1]
    col = {k: (hdr.index(v) if v in hdr else None) for k, v in COLS.items()}
    name_i = hdr.index("Kernel Name")
    out = []
    for r in rows[i + 2:]:
        if len(r) != len(hdr):
            continue
        d = {"kernel": re.sub(r"^void |ustrun::|\(.*$", "", r[name_i])}
        for k, ci in col.items():
            if ci is None:
                d[k] = None
                continue
            v = num(r[ci])
            if v is not None and k == "t":
                v *= UNIT_T.get(units[ci], 1.0)
            if v is not None and k in ("rd", "wr"):
                v *= UNIT_B.get(units[ci], 1.0)
            d[k] = v
        out.append(d)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--manifest", default="")
    ap.add_argument("--cls", default="", help="manifest classes to join, comma separated (e.g. tc_conv or hbm_bn_act)")
    ap.add_argument("--skip-calls", type=int, default=0, help="manifest calls of those classes the capture window skipped")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    launches = load(a.csv)
    calls = []
    if a.manifest:
        m = json.load(open(a.manifest))
        want = set(a.cls.split(","))
        calls = [c for c in m["calls"] if c["cls"] in want][a.skip_calls:]
    li = 0
    rows = []
    for c in calls:
        if li + c["kernels"] > len(launches):
            break
        group = launches[li: li + c["kernels"]]
        li += c["kernels"]
        t = sum(g["t"] for g in group)
        traffic = sum((g["rd"] or 0) + (g["wr"] or 0) for g in group)
        row = {"cls": c["cls"], "entry": c["entry"], "shape": c["meta"], "kernel": group[0]["kernel"], "launches": len(group), "us": t * 1e6,
               "tensor_pipe_pct": max(g["tensor"] or 0 for g in group), "dram_bytes": traffic, "dram_pct": max(g["dram_pct"] or 0 for g in group),
               "regs": group[0]["regs"], "grid": group[0]["grid"], "smem": group[0]["smem"]}
        if c["cls"].startswith("hbm_"):
            row.update(algorithmic_bytes=c["work"], gbs=c["work"] / t / 1e9, traffic_ratio=traffic / c["work"] if c["work"] else None)
        else:
            row.update(flops=c["work"], tflops=c["work"] / t / 1e12)
        rows.append(row)
    if not calls:
        for g in launches:
            rows.append({"kernel": g["kernel"], "us": g["t"] * 1e6, "tensor_pipe_pct": g["tensor"], "dram_bytes": (g["rd"] or 0) + (g["wr"] or 0), "dram_pct": g["dram_pct"],
                         "regs": g["regs"], "grid": g["grid"], "smem": g["smem"], "gbs_dram": ((g["rd"] or 0) + (g["wr"] or 0)) / g["t"] / 1e9})
    for r in rows:
        extra = (f" {r['tflops']:7.1f} TF/s" if "tflops" in r else (f" {r['gbs']:7.0f} GB/s alg, traffic x{r['traffic_ratio']:.2f}" if "gbs" in r else f" {r['gbs_dram']:7.0f} GB/s dram"))
        print(f"{r['kernel'][:34]:34s} {str(r.get('shape') or ''):32s} {r['us']:8.1f} us  tensor {r['tensor_pipe_pct'] or 0:5.1f}%  dram {r['dram_pct'] or 0:5.1f}%  {r['dram_bytes'] / 1e6:8.1f} MB{extra}  regs {r['regs']:.0f} grid {r['grid']:.0f}")
    summ = {}
    if rows:
        tt = sum(r["us"] for r in rows)
        summ = {"launches": len(rows), "us_total": tt, "tensor_pipe_pct_time_weighted": sum((r["tensor_pipe_pct"] or 0) * r["us"] for r in rows) / tt,
                "dram_pct_time_weighted": sum((r["dram_pct"] or 0) * r["us"] for r in rows) / tt, "dram_bytes_per_launch": sum(r["dram_bytes"] for r in rows) / len(rows)}
        if all("flops" in r for r in rows):
            summ["tflops_under_ncu"] = sum(r["flops"] for r in rows) / (tt * 1e-6) / 1e12
        if all("algorithmic_bytes" in r for r in rows):
            ab = sum(r["algorithmic_bytes"] for r in rows)
            summ.update(algorithmic_bytes_per_launch=ab / len(rows), traffic_ratio=sum(r["dram_bytes"] for r in rows) / ab, gbs_algorithmic_under_ncu=ab / (tt * 1e-6) / 1e9)
        print(json.dumps(summ))
    if a.json:
        json.dump({"source": a.csv, "summary": summ, "per_launch": rows}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
