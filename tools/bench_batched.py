#!/usr/bin/env python
"""tools/bench_batched.py [C2] -- round-2 measurement (written without a GPU at the end of round 1): scenes/s of the hot path
when B scenes go through the U-Net as ONE batched sparse tensor (batch index = coordinate column 0, as train_joint.py:82
collates them) followed by one vote per scene, for B = 1, 2, 4, 8.  The 44 launches of the small levels cost about the same
for B times the rows, so the expectation is ~0.75 + 0.7 / B ms per scene.  Inputs resident,
one stream, CUDA events, L2 not flushed (working set of B scenes >> L2 for B >= 2)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from canonicalvoting_b200 import hv_cuda as H  # noqa: E402
from canonicalvoting_b200.engine import MinkUNetEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda", 0)
model = bench.make_model().to(dev)
eng = MinkUNetEngine(model, bench.NCLASSES, True)
out = {"workload": wl}
for B in (1, 2, 4, 8):
    scenes = [bench.scene_for(wl, seed=s) for s in range(B)]
    res, R = scenes[0]["res"], scenes[0]["num_rots"]
    coords, feats, spans, geo = [], [], [], []
    row = 0
    for b, sc in enumerate(scenes):
        c, f = bench.scene_tensors(sc)
        c[:, 0] = b
        coords.append(c)
        feats.append(f)
        spans.append((row, row + len(c)))
        row += len(c)
        p = (c[:, 1:].float() * res).contiguous().to(dev)
        cr, _, dm = H.grid_dims(p, res)
        geo.append((cr, dm))
    coords_d, feats_d = torch.cat(coords).contiguous().to(dev), torch.cat(feats).contiguous().to(dev)

    def step():
        xyz, scale, cls, prob, points = eng.predict(coords_d, feats_d, res=res)
        grids = []
        for (r0, r1), (cr, dm) in zip(spans, geo):
            grids.append(H.forward_host(points[r0:r1], xyz[r0:r1], scale[r0:r1], prob[r0:r1], res, R, cr, dm))
        return grids

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = max(4, 24 // B)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    out["B%d" % B] = {"ms_per_forward": ms, "ms_per_scene": ms / B, "scenes_per_sec": 1e3 * B / ms}
    print("B=%d: %.3f ms per forward, %.3f ms per scene, %.0f scenes/s" % (B, ms, ms / B, 1e3 * B / ms), flush=True)
print(json.dumps(out))
