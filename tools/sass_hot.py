#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report: executed counts, thread utilisation, stall samples.
  tools/sass_hot.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        break
    try:
        data.append((r[ci["Source"]], int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0)))
    except ValueError:
        pass
tot_i = sum(d[2] for d in data); tot_t = sum(d[3] for d in data); tot_s = sum(d[1] for d in data)
print("instructions executed (warp-level): %d, thread-instructions: %d (avg %.1f threads), samples %d, SASS lines %d" % (tot_i, tot_t, tot_t / max(1, tot_i), tot_s, len(data)))
ops = {}
for src, s, i, t in data:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    o = ops.setdefault(op, [0, 0, 0]); o[0] += i; o[1] += t; o[2] += s
print("-- by opcode: warp-inst share, avg threads, stall-sample share")
for op, (i, t, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:25]:
    print("  %-10s %6.2f%%  thr %5.1f  samples %6.2f%%" % (op, 100.0 * i / tot_i, t / max(1, i), 100.0 * s / max(1, tot_s)))
print("-- top SASS lines by stall samples")
for idx in sorted(range(len(data)), key=lambda k: -data[k][1])[:topn]:
    src, s, i, t = data[idx]
    print("  #%4d %6.2f%% samples  exec %10d thr %5.1f  %s" % (idx, 100.0 * s / max(1, tot_s), i, t / max(1, i), src[:90]))
