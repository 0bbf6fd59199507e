#!/usr/bin/env python
"""tools/host_profile.py -- host-side (enqueue) cost of one scene through the engine, split by phase (GPU box)."""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from canonicalvoting_b200 import _lib  # noqa: E402
from canonicalvoting_b200 import hv_cuda as H  # noqa: E402
from canonicalvoting_b200.engine import MinkUNetEngine  # noqa: E402
from canonicalvoting_b200.sparse.coords import _stream  # noqa: E402

dev = torch.device("cuda", 0)
sc = bench.scene_for("C2", 0)
model = bench.make_model().to(dev)
eng = MinkUNetEngine(model, 9, True, pipeline=True)
c_h, f_h = bench.scene_tensors(sc)
c_d, f_d = c_h.to(dev), f_h.to(dev)
res, R = sc["res"], sc["num_rots"]
pts = (c_d[:, 1:].float() * res).contiguous()
corner, _, dims = H.grid_dims(pts, res)
L = _lib.load()
acc = {}


def tick(name, t0):
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
    return t1


for it in range(40):
    if it == 10:
        torch.cuda.synchronize()
        acc.clear()
        w0 = time.perf_counter()
    t = time.perf_counter()
    main = torch.cuda.current_stream()
    with torch.cuda.stream(eng._side):
        cm = eng.build_maps(c_d)
        ready = eng._side.record_event()
    main.wait_event(ready)
    t = tick("build_maps (side stream, 4 syncs)", t)
    arr, out, keep = eng.build(c_d, f_d, cm)
    t = tick("build program (python)", t)
    _lib.check(L.cvb200_sc_run_program(arr, len(arr), _stream()), "run")
    t = tick("run_program (C++ launches)", t)
    eng._ring.append((keep, main.record_event()))
    while len(eng._ring) > 3:
        old, done = eng._ring.popleft()
        done.synchronize()
        del old
    t = tick("ring wait", t)
    xyz, scale, cls, prob = eng.decode(out)
    t = tick("decode", t)
    p = (c_d[:, 1:].float() * res).contiguous()
    o = H.forward_host(p, xyz, scale, prob, res, R, corner, dims)
    t = tick("vote", t)
torch.cuda.synchronize()
w1 = time.perf_counter()
n = 30
for k, v in acc.items():
    print("%-36s %7.3f ms / scene" % (k, 1e3 * v / n))
print("%-36s %7.3f ms / scene (wall incl. GPU drain)" % ("total", 1e3 * (w1 - w0) / n))
