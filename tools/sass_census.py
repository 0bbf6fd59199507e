#!/usr/bin/env python
"""tools/sass_census.py [libcvb200.so] -- per kernel: how many tcgen05 / TMEM / TMA / cp.async / reduction instructions the SASS of
the built library holds (cuobjdump -sass; works without a GPU).  The evidence B200_PROFILING.md asks for."""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "canonicalvoting_b200", "_C", "libcvb200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "LDGSTS", "SYNCS", "REDG", "RED.", "ATOMG", "ATOMS", "HMMA", "DFMA", "BAR.SYNC", "SHFL", "ELECT", "UCGABAR"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter()
        continue
    if kern and "/*" in ln:
        for mn in MNEMONICS:
            if re.search(r"\b" + re.escape(mn.rstrip(".")) + r"\b", ln):
                counts[kern][mn.rstrip(".")] += 1
print("SASS census of %s (sm_100a): instruction counts per kernel" % os.path.basename(so))
for k, c in counts.items():
    if c:
        print("%-58s %s" % (k[-58:], "  ".join("%s %d" % (mn, n) for mn, n in c.items())))
