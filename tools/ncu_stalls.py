"""Per-instruction stall sampling of one .ncu-rep (captured with --import-source on): the top stalled SASS lines and the
stall reasons aggregated over a range of lines.  usage: ncu_stalls.py rep.ncu-rep [first_line last_line] > out.txt"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1])
h, rows = rows[1], rows[2:]
tot = sum(int(r[2]) for r in rows if r[2].isdigit())
print(f"total samples {tot}, {len(rows)} SASS lines")
print("-- 40 most sampled lines: line, samples, not-issued samples, executions, SASS")
for i in sorted(sorted(range(len(rows)), key=lambda i: -int(rows[i][2]) if rows[i][2].isdigit() else 0)[:40]):
    r = rows[i]
    print(f"{i:6d} {r[2]:>7s} {r[3]:>7s} {r[5]:>10s}  {r[1].strip()[:100]}")
lo = int(sys.argv[2]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.Counter(); per = collections.defaultdict(collections.Counter); mix = collections.Counter()
for i in range(lo, min(hi, len(rows))):
    r = rows[i]
    op = r[1].strip().split()
    o = (op[1] if op[0].startswith("@") and len(op) > 1 else op[0]).split(".")[0]
    mix[o] += int(r[5]) if r[5].isdigit() else 0
    for c in cols:
        v = int(r[c]) if r[c].isdigit() else 0
        agg[h[c]] += v; per[h[c]][o] += v
t = sum(agg.values()) or 1
print(f"-- stall reasons over SASS lines [{lo}, {hi})")
for k, v in agg.most_common(10):
    print(f"{k:24s} {v:8d} {100 * v / t:5.1f}%   {per[k].most_common(5)}")
print("-- executed instruction mix of the range:", mix.most_common(12))
