"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python tools/summarise_launches.py gpurun_out/<tag>_launches.csv > profiles/<tag>_launches_summary.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
ours = sum(v for k, v in tot.items() if "rfd::" in k)
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {total / 1e3:.2f} ms of GPU time; rfd:: kernels {100 * ours / total:.1f} % of it")
print("# share %   total us   launches   kernel")
for k, v in tot.most_common():
    print(f"{100 * v / total:7.2f} {v:11.1f} {cnt[k]:8d}   {k}")
