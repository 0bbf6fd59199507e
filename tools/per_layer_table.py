#!/usr/bin/env python
"""tools/per_layer_table.py LAUNCHES.csv -- per-layer times of the convolution program from an ncu launch list
(`--metrics gpu__time_duration.sum`) of bench.py: the sc_conv_persist_kernel launches of every complete step are matched to
the program order MinkUNetEngine.build emits for MinkUNet34C (stem, 4 x [stride-2 conv + BasicBlocks], 4 x [transposed conv +
BasicBlocks on the concatenation], final) and the median over the steps is printed per layer and per level.  ncu times are
cold-cache and serialised (no programmatic-launch overlap): compare shares, not absolutes."""
import csv
import sys

import numpy as np

LEVEL_ROWS = {1: 50000, 2: 17001, 4: 4090, 8: 940, 16: 196}          # C2 scene


def program():
    prog = [("conv0p1s1 (stem 5^3, 4-channel gather)", 1, 3, 32, 125)]

    def blocks(name, ts, cin, planes, n):
        out = []
        for i in range(n):
            c = cin if i == 0 else planes
            out.append(("%s.%d.conv1" % (name, i), ts, c, planes, 27))
            if c != planes:
                out.append(("%s.%d.downsample 1^3" % (name, i), ts, c, planes, 1))
            out.append(("%s.%d.conv2" % (name, i), ts, planes, planes, 27))
        return out
    for conv, ts, c, block, planes, n in (("conv1p1s2", 2, 32, "block1", 32, 2), ("conv2p2s2", 4, 32, "block2", 64, 3),
                                          ("conv3p4s2", 8, 64, "block3", 128, 4), ("conv4p8s2", 16, 128, "block4", 256, 6)):
        prog.append((conv + " 2^3 stride 2", ts, c, c, 8))
        prog += blocks(block, ts, c, planes, n)
    for conv, ts, cin, cout, block, cat, n in (("convtr4p16s2", 8, 256, 256, "block5", 384, 2), ("convtr5p8s2", 4, 256, 128, "block6", 192, 2),
                                               ("convtr6p4s2", 2, 128, 96, "block7", 128, 2), ("convtr7p2s2", 1, 96, 96, "block8", 128, 2)):
        prog.append((conv + " transposed 2^3", ts, cin, cout, 8))
        prog += blocks(block, ts, cat, cout, n)
    prog.append(("final 1^3", 1, 96, 64, 1))
    return prog


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    k, v = h.index("Kernel Name"), h.index("Metric Value")
    us = [float(r[v]) / 1e3 for r in rows[1:] if "sc_conv_persist" in r[k]]
    prog = program()
    steps = np.array(us[:len(us) // len(prog) * len(prog)]).reshape(-1, len(prog))
    med = np.median(steps, 0)
    print("%d complete steps, %d launches per step, %.1f us per step in total" % (len(steps), len(prog), med.sum()))
    per_level = {}
    for (name, ts, cin, cout, k3), t in zip(prog, med):
        per_level.setdefault(ts, [0, 0.0])
        per_level[ts][0] += 1
        per_level[ts][1] += t
    for ts in sorted(per_level):
        n, t = per_level[ts]
        print("level stride %2d (%6d voxels): %2d launches %7.1f us %5.1f %%" % (ts, LEVEL_ROWS[ts], n, t, 100 * t / med.sum()))
    for (name, ts, cin, cout, k3), t in zip(prog, med):
        print("%-40s stride %2d  %3d -> %3d  K^3 = %3d  %7.1f us" % (name, ts, cin, cout, k3, t))


if __name__ == "__main__":
    main(sys.argv[1])
