#!/usr/bin/env python
"""torchrun entry that runs bench.py on every rank, rank 0 under `ncu --metrics gpu__time_duration.sum` (one pass, no kernel
replay, so the NCCL kernel of rank 0 still meets its peers): a per-launch timeline of one rank of an N-GPU step that includes
the ncclReduce.  Usage: python -m torch.distributed.run --nproc-per-node N tools/rank0_ncu.py <log.csv> <bench.py args...>"""
import os
import sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log, args = sys.argv[1], sys.argv[2:]
cmd = [sys.executable, os.path.join(root, "bench.py")] + args
if os.environ.get("RANK", "0") == "0":
    cmd = ["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", log, "-c", "4000"] + cmd
os.execvp(cmd[0], cmd)
