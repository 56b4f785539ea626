#!/bin/bash
# round 2, call Z: occupancy / register budget of k_trace.  tg5: global-tree kernel at 5 CTAs/SM (48 regs), shared-tree kernel free
# (72 regs, 3 CTAs/SM); tg6: global at 6 CTAs (40 regs, 24 B spilled); t5: shared-tree at 5 CTAs (48 regs), global free (76 regs, 3 CTAs)
mkdir -p gpurun_out
export TIRAY_ALLOW_MISSING=1
Q="timeout 200 python tools/perf_probe.py --reps 5"
L=gpurun_out/z_probe.log
for lib in libtiray.so libtiray_v_tg5.so libtiray_v_tg6.so libtiray_v_t5.so; do
  for wl in cornell teapot_mc; do $Q --workload $wl --lib $lib 2>&1 | grep -v "libpng\|total light" >> $L; done
done
cut -c1-200 $L
