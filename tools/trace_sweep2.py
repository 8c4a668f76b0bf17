import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exec(open(os.path.join(ROOT, "tools", "trace_sweep.py")).read().split("def show")[0])
m1 = t[1]
print("mma per-job cycle breakdown: waits | setup | 2xMMA | commit1 | commit2+bookkeeping")
for i in list(range(0, 30)) + list(range(200, 215)):
    e = m1[i]
    print(i, int(e[0]), int(e[1] >> 32), int(e[1] & 0xffffffff), int(e[2] >> 32), int(e[2] & 0xffffffff))
