"""tools/sanitizer_summary.py LOG -- group the errors of a compute-sanitizer log by (kind, kernel, innermost python frames)."""
import collections
import re
import sys

blocks, cur = [], None
for ln in open(sys.argv[1], errors="replace"):
    ln = ln.rstrip("\n")
    m = re.match(r"^========= (Uninitialized|Invalid|Race|Program hit|Barrier|Potential|Error).*", ln)
    if m:
        cur = {"kind": ln[10:70], "kernel": "", "py": []}
        blocks.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"^=========\s+at (.*)", ln)
    if m and not cur["kernel"]:
        cur["kernel"] = re.sub(r"<.*", "", m.group(1))[:70]
    m = re.match(r"^=========\s+Host Frame: (\S+) in (\S+\.py:\d+)", ln)
    if m:
        cur["py"].append("%s@%s" % (m.group(1), m.group(2)))
cnt = collections.Counter((b["kind"], b["kernel"], " < ".join(b["py"][:3])) for b in blocks)
for (kind, kern, py), c in cnt.most_common():
    print("%5d  %s | %s | %s" % (c, kind, kern, py))
print("total", len(blocks))
