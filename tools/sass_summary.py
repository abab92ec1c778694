#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what the hot kernels are built from (tensor core, tensor memory,
TMA, packed FP32, XU), read from the built library with cuobjdump (runs without a GPU).

    python tools/sass_summary.py [lib] > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "cvx_proj_b200", "libapap_b200.so")
KEYS = ("UTCHMMA UTCBAR STTM LDTM UBLKCP UTMALDG UTMASTG UTMAPF SYNCS FFMA2 FADD2 FMUL2 FFMA MUFU DFMA DADD DMUL LDS STS LDG STG "
        "ATOMG RED LDL STL").split()
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = {}
kernels = collections.OrderedDict()
cur = None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
    if m and cur is not None:
        tok = m.group(1).split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        cur[op.split(".")[0]] += 1
        cur["_n"] += 1
names = list(kernels)
pretty = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
print(f"# {os.path.basename(lib)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a; static counts, loops not weighted)")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, STTM / LDTM = tcgen05.st / ld (tensor memory), UBLKCP = cp.async.bulk (1-D TMA),")
print("# UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (tensor-map TMA), SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed FP32x2,")
print("# MUFU = XU pipe, DFMA / DADD / DMUL = FP64, LDG / STG = plain global access, LDS / STS = shared memory, LDL / STL = local (spills)")
rows = []
for mangled, name in zip(names, pretty or names):
    c = kernels[mangled]
    short = re.sub(r"\(.*", "", name).replace("apap::", "")
    rows.append((short, c))
for short, c in sorted(rows):
    print(f"{short:34s} instr={c['_n']:<5d} " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
