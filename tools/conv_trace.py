#!/usr/bin/env python
"""tools/conv_trace.py [mask] -- clock64 timeline of CTA 0 of the persistent convolution (first k-blocks), GPU box."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from canonicalvoting_b200 import _lib, synthetic  # noqa: E402
from canonicalvoting_b200.sparse.coords import CoordinateManager, _ptr, _stream  # noqa: E402

mask = int(sys.argv[1]) if len(sys.argv) > 1 else 0
level = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = _lib.load()
sc = synthetic.make_config("C2", seed=0)
coords = torch.cat([torch.zeros(len(sc["coords"]), 1, dtype=torch.int32), torch.from_numpy(sc["coords"])], 1).cuda()
cm = CoordinateManager(coords)
for ts in (1, 2, 4, 8):
    cm.down(ts)
table = cm.kernel_map(level, 3)
n, cin, cout = cm.levels[level].n, {1: 96, 2: 96, 4: 128, 8: 256, 16: 256}[level], {1: 96, 2: 96, 4: 128, 8: 256, 16: 256}[level]
x = torch.randn(n, cin).cuda()
wt = (torch.randn(27, cout, cin) * 0.05).cuda()
out = torch.empty(n, cout, device="cuda")
trace = torch.zeros(3 * 768 + 16, dtype=torch.int64, device="cuda")
L.cvb200_sc_set_conv_debug(mask)


def go():
    _lib.check(L.cvb200_sc_conv_forward_tc(_ptr(x), n, cin, _ptr(wt), cout, _ptr(table), n, 27, None, _ptr(out), _stream()), "conv")


for _ in range(3):
    go()
torch.cuda.synchronize()
L.cvb200_sc_set_conv_trace(_ptr(trace))
go()
torch.cuda.synchronize()
L.cvb200_sc_set_conv_trace(None)
tc = trace.cpu()
st = tc[3 * 768:]
t = tc[:3 * 768].view(3, 256, 3)
t0 = int(t[t > 0].min())
names = ["MMA ", "GATH", "WTMA"]
print("mask", mask, " columns: role  k-block: [wait start, wait end, issue done] relative cycles; waited = end - start")
for i in list(range(0, 8)) + list(range(40, 52)):
    print("kb %3d  " % i + "   ".join("%s %7d %7d %7d (w %5d)" % (names[r], t[r, i, 0] - t0, t[r, i, 1] - t0, t[r, i, 2] - t0, t[r, i, 1] - t[r, i, 0]) for r in range(3)))
for r in range(3):
    m = 50 if r == 1 else 200
    d = (t[r, 1:m, 2] - t[r, :m - 1, 2]).float()
    print(names[r], "cycles per own k-block: mean %.0f  median %.0f" % (d.mean(), d.median()), " mean wait %.0f" % (t[r, :m, 1] - t[r, :m, 0]).float().mean(), "(GATH = warp 2: every 4th k-block)" if r == 1 else "")
for r in range(3):
    wv = (t[r, :, 1] - t[r, :, 0])
    top = torch.argsort(wv, descending=True)[:6].tolist()
    print(names[r], "longest waits:", ", ".join("kb %d: %d cyc (at %d)" % (i, int(wv[i]), int(t[r, i, 0] - t0)) for i in sorted(top)))
print("last stamp:", int(t.max() - t0), "cycles")

lab = ["entry", "setup done", "griddep wait done", "epilogue: acc_full seen", "epilogue: partial sums issued + bar", "epilogue: counter known", "epilogue done", "CTA done"]
print("CTA 0 milestones (cycles after entry):", ", ".join("%s %d" % (lab[i], int(st[i] - st[0])) for i in range(8) if int(st[i]) > 0))
print("first MMA wait end %d, last commit %d (k-blocks traced %d)" % (int(t[0, 0, 1] - st[0]), int(t[0][t[0][:, 2] > 0][-1, 2] - st[0]), int((t[0][:, 2] > 0).sum())))
