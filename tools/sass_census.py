#!/usr/bin/env python
"""tools/sass_census.py [lib.so] -> JSON: per kernel of libisv_b200.so, the SASS opcode census that backs the claims in
DESIGN.md (which kernels issue FP64 tensor-core DMMA, which only scalar DFMA, where 128-bit shared-memory accesses and
TMA / bulk-copy instructions appear).  Runs `cuobjdump -sass` (works without a GPU)."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "is_vins_b200", "libisv_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
FAMILIES = {"DMMA": r"^DMMA", "DFMA": r"^DFMA", "DMUL": r"^DMUL", "DADD": r"^DADD", "MUFU": r"^MUFU",
            "LDS.128": r"^LDS.*\.128", "LDS.64": r"^LDS.*\.64", "STS.128": r"^STS.*\.128", "STS.64": r"^STS.*\.64",
            "LDG": r"^LDG", "STG": r"^STG", "SHFL": r"^SHFL", "RED/ATOM": r"^(RED|ATOM)", "LDGSTS (cp.async)": r"^LDGSTS",
            "UTMALDG (TMA load)": r"^UTMALDG", "UBLKCP (bulk copy)": r"^UBLKCP", "SYNCS (mbarrier)": r"^SYNCS",
            "HMMA/IMMA/QMMA": r"^(HMMA|IMMA|QMMA)", "UTCMMA (tcgen05)": r"^UTC.*MMA"}
res = {}
cur = None
arch = None
for line in txt.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("isv::", "")
        cur = res.setdefault(name, {"arch": arch, "instructions": 0, "ops": collections.Counter()})
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["instructions"] += 1
        for fam, pat in FAMILIES.items():
            if re.match(pat, op):
                cur["ops"][fam] += 1
out = {k: {"arch": v["arch"], "instructions": v["instructions"], **{f: v["ops"].get(f, 0) for f in FAMILIES if v["ops"].get(f, 0)}}
       for k, v in sorted(res.items())}
print(json.dumps({"library": os.path.relpath(lib, ROOT), "kernels": out}, indent=1))
