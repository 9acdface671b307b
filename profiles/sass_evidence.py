"""SASS evidence for the built library, here on the CPU box:  python profiles/sass_evidence.py > profiles/r02_sass_evidence.txt

Per kernel of libsoccer_b200.so (sm_100a; cuobjdump -xelf + nvdisasm): instruction count and the counts of
  UBLKCP      cp.async.bulk -- the TMA bulk copies that fill the shared-memory table (stage_table)
  SYNCS       mbarrier operations (the completion barrier of those copies)
  STL / LDL   local-memory stores / loads = register spills (the out-of-line slip walk's call frame in the *_slipi / slip_i kernels)
  IMAD.WIDE   64-bit products (Philox)        DADD  fp64 adds (the reference's cumulative walk)       REDUX  warp reductions
"""
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gym_soccer_littman94_b200", "libsoccer_b200.so")
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True, check=True)
cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", os.path.join(td, cub)], capture_output=True, text=True, check=True).stdout
print(__doc__)
KEYS = ("UBLKCP", "SYNCS", "STL", "LDL", "IMAD.WIDE", "DADD", "REDUX")
cur, stats = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    if line.startswith("\t.section"):
        cur = None
    if cur:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins = m.group(2).strip()
            op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
            c = stats[cur]
            c["n"] += 1
            for key in KEYS:
                if op.startswith(key):
                    c[key] += 1
names = subprocess.run(["c++filt"] + list(stats), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':84s} {'instrs':>6s} " + " ".join(f"{k:>9s}" for k in KEYS))
for (k, c), nm in zip(stats.items(), names):
    nm = re.sub(r"\(.*", "", nm).replace("void soccer::", "")
    print(f"{nm[:84]:84s} {c['n']:6d} " + " ".join(f"{c[k]:9d}" for k in KEYS))
