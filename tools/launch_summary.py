"""Summarise the ncu launch list of tools/gpu_r2_ncu.sh (gpurun_out/r02_launches.csv: two forwards of bench.py, cfg2) per kernel.
usage: python tools/launch_summary.py [csv] [n_forwards] > profiles/rNN_launch_summary.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_launches.csv"
nfwd = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with open(path) as f:
    recs = list(csv.DictReader([l for l in f if not l.startswith("==")]))
tot, cnt = collections.OrderedDict(), collections.Counter()
for x in recs:
    n = re.sub(r"\(.*", "", x["Kernel Name"]).replace("m2t::<unnamed>::", "").replace("void ", "").replace("m2t::", "")
    tot[n] = tot.get(n, 0) + float(x["Metric Value"])
    cnt[n] += 1
T = sum(tot.values())
print(f"{nfwd} forwards of `python bench.py --steps 2 --warmup 3 --no-cpu` (cfg2, x4, 16x3x128x128): {len(recs) // nfwd} launches per forward, "
      f"{T / 1e3 / nfwd:.1f} us summed per forward;")
print("ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn|ffconv|tail|branch_prep|head_conv -s 153 -c 102 "
      "(cold caches, serialised: compare shares, not absolutes)")
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v / 1e3 / nfwd:9.1f} us {100 * v / T:6.1f}%  x{cnt[n] // nfwd:<3d} {v / 1e3 / cnt[n]:7.1f} us each  {n}")
