#!/usr/bin/env python
"""tools/profile_scene.py [C2|C5] [n_scenes] -- one scene after another through the engine's stream path (map builder, convolution
program, decode + scan_points, vote), plain launches, no graph: the target of the ncu launch-list / --set full captures (GPU box)."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from canonicalvoting_b200 import hv_cuda as H  # noqa: E402
from canonicalvoting_b200.engine import MinkUNetEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
sc = bench.scene_for(wl, 0)
eng = MinkUNetEngine(bench.make_model().to(dev), 9, True)
c_h, f_h = bench.scene_tensors(sc)
c_d, f_d = c_h.to(dev), f_h.to(dev)
res, R = sc["res"], sc["num_rots"]
pts = (c_d[:, 1:].float() * res).contiguous()
corner, _, dims = H.grid_dims(pts, res)
for it in range(reps):
    xyz, scale, cls, prob, points = eng.predict(c_d, f_d, res=res)
    H.forward_host(points, xyz, scale, prob, res, R, corner, dims)
    torch.cuda.synchronize()
print("done")
