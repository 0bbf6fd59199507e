#!/usr/bin/env python
"""tools/time_unet.py [C2|C5] [fp32|tf32] -- steady-state timing of MinkUNet34C forward on one synthetic scene (GPU box)."""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from canonicalvoting_b200 import sparse as ME  # noqa: E402
from canonicalvoting_b200 import synthetic  # noqa: E402
from canonicalvoting_b200.minkunet import MinkUNet34C, decode_heads  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
ME.set_forward_mode(mode)
sc = synthetic.make_config(wl, seed=0)
coords = torch.cat([torch.zeros(len(sc["coords"]), 1, dtype=torch.int32), torch.from_numpy(sc["coords"])], 1)
feats = torch.from_numpy(sc["feats"]) * 2 - 1
torch.manual_seed(0)
model = MinkUNet34C(3, 64).cuda().eval()
cd, fd = coords.cuda(), feats.cuda()


def step():
    with torch.no_grad():
        out = model(ME.SparseTensor(fd, cd, device="cuda"))
        return decode_heads(out.F)


for _ in range(3):
    step()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
t0 = time.perf_counter()
for a, b in ev:
    a.record(); step(); b.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / len(ev)
ms = sorted(a.elapsed_time(b) for a, b in ev)
print("%s %s: MinkUNet34C forward+decode  median %.3f ms (min %.3f, wall/iter %.3f ms), N=%d" % (wl, mode, ms[len(ms) // 2], ms[0], wall * 1e3, len(cd)))
# per-level voxel counts
st = ME.SparseTensor(fd, cd, device="cuda")
with torch.no_grad():
    model(st)
print("voxels per tensor stride:", {k: v.n for k, v in sorted(st.coordinate_manager.levels.items())})
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))

# ---- per-convolution breakdown (CUDA events around every conv_table_forward call)
import collections  # noqa: E402
from canonicalvoting_b200.sparse import functional as Fn  # noqa: E402
rec = []
orig = Fn.conv_table_forward


def timed(x, w, table, bias=None, mode=None):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = orig(x, w, table, bias, mode)
    b.record()
    rec.append(((table.shape[0], w.shape[1], w.shape[2], table.shape[1]), a, b))
    return out


Fn.conv_table_forward = timed
step()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, a, b in rec:
    agg.setdefault(key, []).append(a.elapsed_time(b))
print("%8s %5s %5s %4s %5s %10s %10s" % ("n_out", "cin", "cout", "k3", "calls", "ms/call", "ms total"))
tot = 0.0
for key, v in agg.items():
    tot += sum(v)
    print("%8d %5d %5d %4d %5d %10.3f %10.3f" % (key + (len(v), sum(v) / len(v), sum(v))))
print("total conv ms (incl. weight transposes):", tot)
