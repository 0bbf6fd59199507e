#!/usr/bin/env python
"""tools/conv_probe.py [C2|C5] -- which side bounds the persistent tcgen05 convolution?  (GPU box)

Times single convolutions of the U-Net's shapes on the real kernel maps of a synthetic scene, in normal operation and
with parts of the kernel switched off (cvb200_sc_set_conv_debug): no gather copies / no zero-fill copies / no MMA /
no weight TMA.  CUDA events around 20 back-to-back launches, L2 warm (the feature matrix and the weights of one layer fit)."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from canonicalvoting_b200 import _lib, synthetic  # noqa: E402
from canonicalvoting_b200.sparse.coords import CoordinateManager, _ptr, _stream  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
L = _lib.load()
sc = synthetic.make_config(wl, seed=0)
coords = torch.cat([torch.zeros(len(sc["coords"]), 1, dtype=torch.int32), torch.from_numpy(sc["coords"])], 1).cuda()
cm = CoordinateManager(coords)
for ts in (1, 2, 4, 8):
    cm.down(ts)
g = torch.Generator().manual_seed(0)


def run(table, n_in, cin, cout, reps=20):
    n_out, k3 = table.shape
    x = torch.randn(n_in, cin, generator=g).cuda()
    wt = (torch.randn(k3, cout, cin, generator=g) * 0.05).cuda()
    out = torch.empty(n_out, cout, device="cuda")

    def go():
        rc = L.cvb200_sc_conv_forward_tc(_ptr(x), n_in, cin, _ptr(wt), cout, _ptr(table), n_out, k3, None, _ptr(out), _stream())
        _lib.check(rc, "conv")
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        go()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


cases_all = [("L1 3^3 96->96", cm.kernel_map(1, 3), cm.levels[1].n, 96, 96),
         ("L1 3^3 128->96", cm.kernel_map(1, 3), cm.levels[1].n, 128, 96),
         ("L2 3^3 96->96", cm.kernel_map(2, 3), cm.levels[2].n, 96, 96),
         ("L4 3^3 128->128", cm.kernel_map(4, 3), cm.levels[4].n, 128, 128),
         ("L8 3^3 256->256", cm.kernel_map(8, 3), cm.levels[8].n, 256, 256),
         ("L16 3^3 256->256", cm.kernel_map(16, 3), cm.levels[16].n, 256, 256),
         ("L1 up 2^3 96->96", cm._down[1]["up_table"], cm.levels[2].n, 96, 96)]
cases = cases_all if "all" in sys.argv else cases_all[:1] + cases_all[2:3] + cases_all[4:5]
modes = [(0, "normal"), (13, "barriers only"), (1, "no gather"), (4, "no MMA"), (8, "no weight TMA"), (9, "MMA only"), (5, "weights only"), (12, "gather only"), (2, "no zero-fill copies")]
for impl in (3,):
    for name, table, n_in, cin, cout in cases:
        pairs = int((table >= 0).sum())
        line = "impl %d %-18s rows %6d pairs %8d:" % (impl, name, table.shape[0], pairs)
        for mask, label in (modes if impl == 3 else modes[:1]):
            L.cvb200_sc_set_conv_debug(max(mask, 0))
            us = run(table, n_in, cin, cout)
            line += "  %s %.1f us" % (label, us)
            if mask == 0:
                line += " (%.1f TF/s alg.)" % (2.0 * pairs * cin * cout / us / 1e6)
        L.cvb200_sc_set_conv_debug(0)
        print(line, flush=True)

# pieces per split tile: planner's choice (0) vs forced
for name, table, n_in, cin, cout in []:
    line = "ks sweep %-18s:" % name
    for ks in (0, 1, 2, 4, 8, 12, 16, 24, 32):
        L.cvb200_sc_set_conv_debug(ks << 8)
        line += "  ks=%d %.1f us" % (ks, run(table, n_in, cin, cout))
    L.cvb200_sc_set_conv_debug(0)
    print(line, flush=True)
