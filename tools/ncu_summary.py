#!/usr/bin/env python
"""Summarise `ncu --set full --import-source on` reports (brought back in gpurun_out/) into small text files for profiles/.

    python tools/ncu_summary.py gpurun_out/r02a_ncu_<scene>.ncu-rep [...] --out profiles/r02a

Per report: <out>_<scene>_metrics.csv   (the raw page, selected rows: time, instruction counts, issue / pipe utilisation,
                                         occupancy, divergence, DRAM bytes, local-memory sectors, warp-stall reasons)
            <out>_<scene>_opcodes.txt   (executed warp-instructions per opcode, from the source page's SASS view)
and one table <out>_summary.md over all of them.  Needs `ncu` (reads reports only; no GPU)."""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "smsp__sass_branch_targets_threads_divergent.sum",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "sm__sass_inst_executed_op_local_ld.sum",
    "sm__sass_inst_executed_op_local_st.sum",
]


STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio$")
SCALE = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return [(h, u, v) for h, u, v in zip(hdr, units, vals)]


def opcode_histogram(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    while rows and "Source" not in rows[0]:      # leading "Kernel Name", ... lines
        rows = rows[1:]
    if not rows:
        return None, 0
    hdr = rows[0]
    try:
        i_src = hdr.index("Source")
        i_cnt = next(k for k, h in enumerate(hdr) if h.strip() in ("# Warp Instructions Executed", "Warp Instructions Executed", "Instructions Executed"))
    except (ValueError, StopIteration):
        return None, 0
    hist = collections.Counter()
    for r in rows[1:]:
        if len(r) <= max(i_src, i_cnt):
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", r[i_src])
        if not m:
            continue
        try:
            hist[m.group(1)] += int(float(r[i_cnt].replace(",", "")))
        except ValueError:
            pass
    return hist, sum(hist.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    table = []
    for rep in args.reports:
        scene = re.sub(r".*_ncu_", "", os.path.basename(rep)).replace(".ncu-rep", "")
        rows = raw_page(rep)
        d = {h: v for h, u, v in rows}
        unit = {h: u for h, u, v in rows}
        with open(f"{args.out}_{scene}_metrics.csv", "w") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit", "value"])
            for h, u, v in rows:
                if h in KEEP or STALL.match(h):
                    w.writerow([h, u, v])
        hist, total = opcode_histogram(rep)
        if hist:
            with open(f"{args.out}_{scene}_opcodes.txt", "w") as f:
                f.write(f"# pe_render_kernel, {scene}: warp-level instructions executed per opcode (ncu source page, one launch); total {total}\n")
                for op, n in hist.most_common():
                    f.write(f"{op:12s} {n:12d} {100.0 * n / total:6.2f}%\n")

        def g(k, default="nan"):
            try:
                return float(d.get(k, default).replace(",", "")) * SCALE.get(unit.get(k, ""), 1.0)      # times in ms, bytes in MB
            except ValueError:
                return float("nan")
        px = g("launch__grid_size") * g("launch__block_size")
        stalls = sorted(((float(v), STALL.match(h).group(1)) for h, u, v in rows if STALL.match(h) and STALL.match(h).group(1) != "selected"), reverse=True)[:5]
        table.append((scene, g("gpu__time_duration.sum"), g("smsp__inst_executed.sum"), g("smsp__inst_executed.sum") * 32 / px if px else float("nan"),
                      g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("smsp__thread_inst_executed_per_inst_executed.ratio"),
                      g("sm__warps_active.avg.pct_of_peak_sustained_active"), g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                      g("dram__bytes_write.sum"), g("launch__registers_per_thread"),
                      g("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum") + g("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
                      ", ".join(f"{h} {v:.2f}" for v, h in stalls)))
    with open(f"{args.out}_summary.md", "w") as f:
        f.write("| scene | kernel ms (ncu) | warp-instr | instr / thread | issue slots used % | active threads / instr | warps active % of max | FMA pipe % | DRAM write (MB) | regs | local sectors | warps stalled per issued instruction, top reasons |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in table:
            f.write(f"| {r[0]} | {r[1]:.3f} | {r[2]:.4g} | {r[3]:.0f} | {r[4]:.1f} | {r[5]:.2f} | {r[6]:.1f} | {r[7]:.1f} | {r[8]:.1f} | {r[9]:.0f} | {r[10]:.3g} | {r[11]} |\n")
    print(open(f"{args.out}_summary.md").read())


if __name__ == "__main__":
    sys.exit(main())
