"""Summarise an .ncu-rep (one `ncu --set full` capture) into a small text file for profiles/.

  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.txt [launch-index]

Writes the headline counters of the chosen launch (default 0), the executed-instruction mix per
opcode and the SASS lines with the most stall samples / shared-memory wavefronts.
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    return list(csv.reader(io.StringIO(out.decode(errors="replace"))))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    lines = ["summary of %s (launch %d)" % (rep, which), ""]
    raw = ncu_csv(rep, "raw")
    hdr, units, rows = raw[0], raw[1], raw[2:]
    ix = {n: i for i, n in enumerate(hdr)}
    lines.append("launches captured: %d" % len(rows))
    for r in rows:
        lines.append("  %s  grid %s block %s  %s %s" % (r[ix["Kernel Name"]][:60], r[ix.get("Grid Size", 0)], r[ix.get("Block Size", 0)],
                                                      r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]]))
    lines.append("")
    r = rows[which]
    for k in KEYS:
        if k in ix:
            lines.append("%-88s %16s %s" % (k, r[ix[k]], units[ix[k]]))
    src = ncu_csv(rep, "source")
    # the source page holds one table per launch; take the first
    h = src[1]
    sx = {n: i for i, n in enumerate(h)}
    ops = collections.Counter()
    samples = collections.Counter()
    total = 0
    per_line = []
    for row in src[2:]:
        if len(row) < len(h) or not row[sx["Instructions Executed"]].isdigit():
            if len(row) < 3:
                break
            continue
        toks = row[sx["Source"]].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        base = op.rstrip(";").split(".")[0]
        n = int(row[sx["Instructions Executed"]])
        s = int(row[sx["# Samples"]] or 0)
        # kernels without shared memory have no such columns
        w = int(row[sx["L1 Wavefronts Shared"]] or 0) if "L1 Wavefronts Shared" in sx else 0
        wi = int(row[sx["L1 Wavefronts Shared Ideal"]] or 0) if "L1 Wavefronts Shared Ideal" in sx else 0
        ops[base] += n
        samples[base] += s
        total += n
        per_line.append((s, n, w, wi, row[sx["Source"]].strip()))
    lines += ["", "executed warp instructions (first launch of the source page): %d" % total]
    for k, v in ops.most_common(30):
        lines.append("  %-10s %12d %6.2f%%  stall samples %7d" % (k, v, 100.0 * v / max(total, 1), samples[k]))
    lines += ["", "SASS lines with most stall samples:  samples / executed / source"]
    for s, n, w, wi, t in sorted(per_line, reverse=True)[:20]:
        lines.append("  %7d %10d  %s" % (s, n, t))
    tw = sum(p[2] for p in per_line)
    ti = sum(p[3] for p in per_line)
    lines += ["", "shared-memory wavefronts: %d (ideal %d)  - top lines: wavefronts / ideal / executed / source" % (tw, ti)]
    for s, n, w, wi, t in sorted(per_line, key=lambda p: -p[2])[:16]:
        lines.append("  %9d %9d %9d  %s" % (w, wi, n, t))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
