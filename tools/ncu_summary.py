#!/usr/bin/env python
"""Summarise ncu outputs brought back from gpurun into small text files under profiles/.
  tools/ncu_summary.py launches gpurun_out/launches_r01.csv
  tools/ncu_summary.py rep gpurun_out/prof_trace_r01.ncu-rep [more.ncu-rep ...]
"""
import collections, csv, re, subprocess, sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
           "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
           "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_imc_miss_per_warp_active.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]; kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0]); total = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn]); v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        agg[name][0] += 1; agg[name][1] += v; total += v
    print("# %s : per-kernel device time (gpu__time_duration.sum, cold-cache serialised launches: compare SHARES)" % path)
    print("%-44s %8s %14s %8s" % ("kernel", "launches", "total_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %8d %14.1f %7.1f%%" % (k[:44], n, t, 100 * t / total))
    print("%-44s %8d %14.1f" % ("TOTAL", sum(v[0] for v in agg.values()), total))


def rep(path):
    if path.endswith(".csv"):                 # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv > x_raw.csv
        out = open(path, errors="replace").read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s" % path)
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        print("## launch id %s  %s" % (r[hdr.index("ID")], r[hdr.index("Kernel Name")]))
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m); print("  %-72s %16s %s" % (m, r[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        for p in sys.argv[2:]:
            rep(p); print()
