#!/usr/bin/env python
"""Per-kernel digest of the SASS in libptk.so (cuobjdump -sass): instruction count, registers are in the ncu
summaries; here the mnemonics that show what a kernel is made of - fp64 arithmetic, global/shared/local memory
operations, atomics, warp collectives, and the Blackwell/Hopper async-copy instructions (UBLKCP = TMA bulk copy,
SYNCS = mbarrier).  usage: python profiles/sass_digest.py [libptk.so] > profiles/r2_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "ptudes_lab_b200", "csrc", "libptk.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
GROUPS = [("fp64", r"^D(ADD|MUL|FMA|SETP|MNMX)"), ("mufu", r"^MUFU"), ("ld.global", r"^LDG"), ("st.global", r"^STG"),
          ("ld.shared", r"^LDS"), ("st.shared", r"^STS"), ("local (spill)", r"^(LDL|STL)"), ("atomics", r"^(ATOMG|ATOM|RED)"),
          ("shfl/vote/match/redux", r"^(SHFL|VOTE|MATCH|REDUX)"), ("barrier", r"^BAR"), ("TMA bulk copy", r"^UBLKCP"),
          ("mbarrier", r"^SYNCS"), ("membar/fence", r"^(MEMBAR|FENCE)"), ("call", r"^CALL")]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for name, pat in GROUPS:
            if re.match(pat, op):
                counts[cur][name] += 1
print(f"# cuobjdump -sass digest of {os.path.basename(lib)} (sm_100a); columns = static instruction counts")
for fn, c in counts.items():
    if "ptk" not in fn:
        continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"{name}: {c['total']} instructions; " + ", ".join(f"{k} {c[k]}" for k, _ in GROUPS if c[k]))
