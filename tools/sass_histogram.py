"""cuobjdump -sass opcode histogram of librfdnet_b200.so, per kernel: the tcgen05 / TMEM / bulk-copy / cluster evidence
(UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, ACQBULK / PREEXIT = griddepcontrol, REDUX = redux.sync, UCGABAR = barrier.cluster).
usage: python tools/sass_histogram.py > profiles/<tag>_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rfdnet_b200", "librfdnet_b200.so")
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UBLKPF", "SYNCS", "UCGABAR", "ACQBULK", "PREEXIT", "REDUX",
       "LDGSTS", "FENCE", "HMMA", "FFMA", "DFMA", "DADD", "DMUL", "MUFU", "ATOM", "ATOMG", "RED", "ATOMS", "LDS", "STS", "LDG", "STG",
       "SHFL", "VOTE", "BAR", "CCTL", "MEMBAR", "ERRBAR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
tot = collections.Counter()
print(f"# {os.path.relpath(LIB, ROOT)}: {len(hist)} kernels, arch sm_100a")
print("# kernel | instructions | " + " ".join(KEY))
for k, c in hist.items():
    tot.update(c)
    cols = " ".join(f"{op}={c[op]}" for op in KEY if c[op])
    print(f"{k} | {sum(c.values())} | {cols}")
print("# whole library: " + " ".join(f"{op}={tot[op]}" for op in KEY if tot[op]))
