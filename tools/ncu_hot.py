#!/usr/bin/env python
"""tools/ncu_hot.py -- per-SOURCE-LINE hot spots of one kernel from an .ncu-rep.

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep <kernel-regex> [--so canonicalvoting_b200/_C/libcvb200.so] [--top 25]

`ncu --page source --csv` only exports the SASS view; this joins it with the line table
of the cubin (nvdisasm -g; compile with -lineinfo) and aggregates stall samples and
executed instructions per CUDA source line.  Works on the CPU box (no GPU needed)."""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile


def line_table(so, kernel_re):
    """SASS offset -> (file, line) for the first function whose mangled name matches."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    table = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, inside, loc = None, False, None
        for ln in out.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                inside = re.search(kernel_re, m.group(1)) is not None
                loc = None
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                table[int(m.group(1), 16)] = (loc, m.group(2).strip())
        if table:
            break
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--so", default="canonicalvoting_b200/_C/libcvb200.so")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--launch", type=int, default=0, help="which matching launch in the report")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", "regex:" + a.kernel],
                         capture_output=True, text=True).stdout
    # the csv holds one block per launch, each starting with a "Kernel Name" row
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(row)
    blk = blocks[a.launch]
    hdr = blk[0]
    ia, isamp, iinst, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    table = line_table(a.so, a.kernel)
    base = None
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot_s = tot_i = 0
    for r in blk[1:]:
        if len(r) <= max(ia, isamp, iinst):
            continue
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        off = addr - base
        s = int(r[isamp] or 0)
        n = int(r[iinst] or 0)
        loc = table.get(off, (None, ""))[0]
        agg[loc][0] += s
        agg[loc][1] += n
        agg[loc][2] += 1
        tot_s += s
        tot_i += n
    print("kernel %s: %d stall samples, %d warp instructions executed, %d SASS lines" % (a.kernel, tot_s, tot_i, len(blk) - 1))
    src_cache = {}
    print("%7s %7s %6s  %s" % ("samp%", "inst%", "#sass", "source"))
    for loc, (s, n, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        text = ""
        if loc:
            for root in ("canonicalvoting_b200/csrc", "."):
                p = os.path.join(root, loc[0])
                if os.path.exists(p):
                    src_cache.setdefault(p, open(p).read().splitlines())
                    if loc[1] - 1 < len(src_cache[p]):
                        text = src_cache[p][loc[1] - 1].strip()[:100]
                    break
        print("%6.1f%% %6.1f%% %6d  %s:%s  %s" % (100.0 * s / max(tot_s, 1), 100.0 * n / max(tot_i, 1), k,
                                                 loc[0] if loc else "?", loc[1] if loc else "?", text))


if __name__ == "__main__":
    main()
