"""Summarises an .ncu-rep capture (made on the GPU box, read here) into a small text file for profiles/.

usage: python profiles/summarize.py gpurun_out/<name>.ncu-rep profiles/<name>.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__cycles_active.avg", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    lines = []
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr = raw[0]
    units = raw[1] if len(raw) > 2 and not raw[1][0].isdigit() else None
    for row in raw[1:]:
        if not row or not row[0].isdigit():
            continue
        d = dict(zip(hdr, row))
        lines.append("== launch id %s: %s  grid %s block %s" % (d.get("ID"), d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d:
                u = dict(zip(hdr, units)).get(k, "") if units else ""
                lines.append("  %-70s %s %s" % (k, d[k], u))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, agg, ops = None, [], collections.Counter()
    total_samples = 0
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) >= 8 and r[0].isdigit():
            try:
                agg.append((int(r[7]), int(r[6]) if r[6].isdigit() else 0, cur, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
        elif len(r) >= 8 and r[0] == "" and r[2].startswith("0x"):
            try:
                n = int(r[7])
            except ValueError:
                continue
            s = re.sub(r"^@!?U?P\d+\s+", "", r[3].strip())
            ops[s.split()[0] if s else "?"] += n
    if agg:
        tot = sum(a[0] for a in agg)
        total_samples = sum(a[1] for a in agg) or 1
        lines.append("")
        lines.append("== executed warp-instructions by source line (top 30 of %d total)" % tot)
        for a in sorted(agg, reverse=True)[:30]:
            lines.append("  %6.2f%% inst %6.2f%% stall-samples  %s:%d  %s" % (100.0 * a[0] / tot, 100.0 * a[1] / total_samples, a[2], a[3], a[4]))
        lines.append("")
        lines.append("== executed warp-instructions by opcode (top 25)")
        t2 = sum(ops.values()) or 1
        for k, v in ops.most_common(25):
            lines.append("  %-22s %6.2f%%" % (k, 100.0 * v / t2))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
