"""Summarise an ncu source page of a warp-specialised kernel: mbarrier try-wait retries per barrier offset,
sample shares, headline metrics.  Usage: python tools/ncu_roles.py report.ncu-rep"""
import collections, csv, subprocess, sys, io, re
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; R = []; name = ""
for r in rows:
    if r and r[0] == "Kernel Name": name = r[1]
    if r and r[0] == "Address": hdr = r; continue
    if hdr and r and r[0].startswith("0x"): R.append(r)
si, sc, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[si]) for r in R)
print(name[:90], "samples", tot)
waits = collections.defaultdict(lambda: [0, 0])
for k, r in enumerate(R):
    m = re.search(r"SYNCS.PHASECHK.TRANS64.TRYWAIT \w+, \[(\w+)\+URZ(\+0x[0-9a-f]+)?\]", r[sc])
    if m:
        off = m.group(2) or "+0x0"
        waits[off][0] += int(r[ie]); waits[off][1] += int(r[si]) + (int(R[k + 1][si]) if k + 1 < len(R) else 0)
for off, (ex, sm) in sorted(waits.items(), key=lambda x: int(x[0], 16)): print(f"  try_wait [{off}] executed {ex:9d}  samples {sm}")
op = collections.Counter()
for r in R:
    t = r[sc].strip().split()
    if not t: continue
    o = t[1] if t[0].startswith("@") else t[0]
    op[o] += int(r[si])
print("  top opcodes by samples:", op.most_common(8))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h = rr[0]
for want in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_sectors_srcunit_tex.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__cycles_active.avg"):
    if want in h:
        print(f"  {want} = {rr[2][h.index(want)]} {rr[1][h.index(want)]}")
