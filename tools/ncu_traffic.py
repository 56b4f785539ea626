#!/usr/bin/env python
"""Per-kernel DRAM traffic from an ncu CSV log (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv --log-file x.csv ...`): prints the JSON object bench.py reads from profiles/r0N_dram_traffic.json for one workload.
  tools/ncu_traffic.py gpurun_out/k_traffic_cornell.csv [raw_full.csv]  ->  {"k_trace<0>": {...}, "k_shade<0>": {...}, ...}"""
import collections, csv, json, re, sys


def parse(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]; kid, kn, mn, mu, mv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", "")); u = r[mu]
        if r[mn] == "gpu__time_duration.sum":
            v *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1e-9)
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        per.setdefault(r[kid], {"name": re.sub(r"\(.*", "", r[kn]).replace("void ", "")})[r[mn]] = v
    return per


def main():
    per = parse(sys.argv[1])
    agg = collections.OrderedDict()
    for rec in per.values():
        a = agg.setdefault(rec["name"], {"launches_captured": 0, "busy_launches": 0, "dram_bytes_total": 0.0, "time_s_total": 0.0})
        t = rec.get("gpu__time_duration.sum", 0.0); b = rec.get("dram__bytes_read.sum", 0.0) + rec.get("dram__bytes_write.sum", 0.0)
        a["launches_captured"] += 1; a["busy_launches"] += int(t > 20e-6); a["dram_bytes_total"] += b; a["time_s_total"] += t
    out = collections.OrderedDict()
    for k, a in agg.items():
        if a["time_s_total"] <= 0 or not k.startswith(("k_trace", "k_shadow", "k_shade<", "k_tail", "k_bdpt")):
            continue
        a["dram_bytes_per_launch_all"] = a["dram_bytes_total"] / a["launches_captured"]
        a["dram_GBps"] = a["dram_bytes_total"] / a["time_s_total"] / 1e9
        out[k] = a
    if len(sys.argv) > 2:            # --set full raw page of the same workload: issue-slot figures of the first busy k_trace launch
        rows = list(csv.reader(open(sys.argv[2], errors="replace")))
        hdr = rows[0]
        for r in rows[2:]:
            if len(r) == len(hdr) and "k_trace" in r[hdr.index("Kernel Name")]:
                g = lambda m: float(r[hdr.index(m)].replace(",", ""))
                out["ncu_full"] = {"kernel": re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", ""),
                                   "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                   "threads_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
                                   "dram_throughput_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                                   "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
                                   "shared_bank_conflicts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")}
                # the bound that matters for these kernels: fraction of the lane-issue peak (issue slots busy x lanes active per instruction)
                out["ncu_full"]["lane_issue_frac"] = out["ncu_full"]["issue_active_pct"] / 100.0 * out["ncu_full"]["threads_per_instruction"] / 32.0
                break
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
