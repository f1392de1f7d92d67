"""Summarise one .ncu-rep (ncu --set full --import-source on): the metrics the DESIGN / profiles notes quote, the stall mix per issued
instruction, the hottest SASS lines and the shared-memory wavefronts by opcode.  usage: python scripts/ncu_summary.py file.ncu-rep [header text]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
print(sys.argv[2] if len(sys.argv) > 2 else rep)
NCU = "/usr/local/cuda/bin/ncu"
raw = subprocess.run([NCU, "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
print(f"kernel: {m.get('Kernel Name')}  grid: {m.get('Grid Size')}  block: {m.get('Block Size')}")
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sass__inst_executed_register_spilling", "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    if k in m and m[k] != "":
        print(f"  {k} = {m[k]} {u.get(k, '')}")
for k in sorted(m):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        try:
            v = float(m[k].replace(",", ""))
        except ValueError:
            continue
        if v >= 0.05:
            print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:>22s} = {v:.3f}")
src = subprocess.run([NCU, "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hi = next((i for i, r in enumerate(srows) if r and r[0] == "Address"), None)
if hi is not None:
    h = srows[hi]
    srows = srows[hi:]
    ci = {n: i for i, n in enumerate(h)}
    samp = ci["# Samples"] if "# Samples" in ci else ci["Warp Stall Sampling (All Samples)"]
    srcc = ci.get("Source")
    tot = 0; lines = []
    for r in srows[1:]:
        try:
            s = int(r[samp].replace(",", ""))
        except (ValueError, IndexError):
            continue
        tot += s; lines.append((s, r[srcc]))
    lines.sort(reverse=True)
    print(f"hottest SASS lines (of {tot} samples):")
    for s, t in lines[:25]:
        print(f"  {100.0 * s / max(tot, 1):5.1f}%  {t.strip()[:110]}")
    ops = collections.Counter()
    for s, t in lines:
        op = t.strip().split(" ")[0] if not t.strip().startswith("@") else t.strip().split(" ")[1]
        ops[op.split(".")[0]] += s
    print("samples by opcode family:", ", ".join(f"{k} {100.0 * v / max(tot, 1):.1f}%" for k, v in ops.most_common(14)))
