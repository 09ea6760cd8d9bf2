"""Summarise gpurun_out/clip_launches.csv (ncu gpu__time_duration launch list of tools/bench_clip.py) per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/clip_launches.csv"
with open(path) as f:
    recs = list(csv.DictReader([l for l in f if not l.startswith("==")]))
names = [x["Kernel Name"] for x in recs]
start = [i for i, n in enumerate(names) if "resize" in n][-1]
recs = recs[start:]
tot, cnt = collections.OrderedDict(), collections.Counter()
for x in recs:
    n = re.sub(r"\(.*", "", x["Kernel Name"]).replace("m2t::<unnamed>::", "").replace("void ", "").replace("m2t::", "")
    tot[n] = tot.get(n, 0) + float(x["Metric Value"])
    cnt[n] += 1
T = sum(tot.values())
print(f"one forward: {len(recs)} launches, {T / 1e3:.1f} us summed (ncu: cold caches, serialised)")
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v / 1e3:9.1f} us {100 * v / T:5.1f}%  x{cnt[n]:3d}  {n}")
if "-v" in sys.argv:
    for x in recs:
        print(f"{float(x['Metric Value']) / 1e3:8.1f} us  {x['Kernel Name'][:60]}")
