#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into the handful of numbers DESIGN.md / profiles/*.md cite.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_register_spilling",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    for r in rows:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"), " grid:", d.get("Grid Size"), " block:", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]} {u[k]}")
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                v = float(d[k] or 0)
                if v >= 0.05:
                    print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:>22s} = {v:.3f}")
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hi = next(i for i, r in enumerate(rows) if r and r[0].strip() == "Address")   # row 0 is the kernel name
        rows = rows[hi:]
        h = rows[0]
        print("source columns:", [c for c in h][:12])
        def col(name):
            for i, c in enumerate(h):
                if c.strip() == name:
                    return i
            return None
        si = col("# Samples") or col("Warp Stall Sampling (All Samples)") or col("Warp Stall Sampling (All Cycles)")
        srci = col("Source")
        if si is None:
            for i, c in enumerate(h):
                if "Sampl" in c:
                    si = i
                    break
        body = [r for r in rows[1:] if len(r) > si and r[si].replace('.', '').isdigit()]
        body.sort(key=lambda r: -float(r[si]))
        tot = sum(float(r[si]) for r in body) or 1
        for r in body[:n]:
            print(f"  {float(r[si]) / tot * 100:5.1f}%  {r[srci][:150] if srci is not None else r[:3]}")
        # shared-memory wavefronts per opcode (where the LSU pipe's time goes)
        wi, ei = col("L1 Wavefronts Shared"), col("Instructions Executed")
        if wi is not None and srci is not None:
            by = {}
            for r in rows[1:]:
                if len(r) > wi and r[wi].isdigit() and int(r[wi]) > 0:
                    op = r[srci].split()[0] if not r[srci].strip().startswith("@") else r[srci].split()[1]
                    w, e = by.get(op, (0, 0))
                    by[op] = (w + int(r[wi]), e + int(r[ei]))
            tw = sum(w for w, _ in by.values()) or 1
            print("shared-memory wavefronts by opcode:")
            for op, (w, e) in sorted(by.items(), key=lambda kv: -kv[1][0]):
                print(f"  {op:<16s} {w:>14d} wavefronts ({w / tw * 100:5.1f}%)  {e:>12d} warp instrs  {w / max(e, 1):.2f}/instr")


if __name__ == "__main__":
    main()
