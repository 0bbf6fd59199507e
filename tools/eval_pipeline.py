#!/usr/bin/env python
"""tools/eval_pipeline.py [C2|C5] -- the eval_joint.py scene pipeline stage by stage on one synthetic scene (GPU box):
U-Net engine + decode (random-init weights: timing only) -> Hough voting on the scene's synthetic predictions ->
candidate loop with the LCC back-projection check (eval_joint.py:195-263) -> per-class OBB NMS (:265-280).
CUDA events per stage, median of 10; prints one JSON line (BASELINE configs[4] is the C5 case)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
import hough_voting  # noqa: E402
import hv_cuda  # noqa: E402
from canonicalvoting_b200.engine import MinkUNetEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda", 0)
sc = bench.scene_for(wl, 0)
R, res = sc["num_rots"], sc["res"]
model = bench.make_model().to(dev)
eng = MinkUNetEngine(model, 9, True)
c_h, f_h = bench.scene_tensors(sc)
c_d, f_d = c_h.to(dev), f_h.to(dev)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
pts, xyz, scale, obj, cls = d(sc["points"]), d(sc["xyz"]), d(sc["scale"]), d(sc["obj"]), d(sc["class_pred"])
res_t = torch.tensor(res, dtype=torch.float32, device=dev)
rots_t = torch.tensor(R, dtype=torch.int32, device=dev)
thresh = 60.0 * R / 120          # eval_joint.py:18 assumes num_rots = 120


def stage_times(n=10):
    t = {k: [] for k in ("unet_decode", "vote", "back_project", "nms")}
    out = None
    for it in range(n + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        eng.predict(c_d, f_d)
        ev[1].record()
        go, gr, gs = hv_cuda.forward(pts, xyz, scale, obj, res_t, rots_t)
        ev[2].record()
        boxes, scores, classes = hough_voting.back_project(go, gr, gs, pts, xyz, obj, cls, res, thresh_high=thresh)
        ev[3].record()
        keep = hough_voting.nms_per_class(boxes, scores, classes, 9, 0.3)
        ev[4].record()
        torch.cuda.synchronize()
        if it >= 2:
            for i, k in enumerate(t):
                t[k].append(ev[i].elapsed_time(ev[i + 1]))
        out = (len(boxes), len(keep))
    return {k: float(np.median(v)) for k, v in t.items()}, out


ms, (n_boxes, n_keep) = stage_times()
print(json.dumps({"workload": wl, "points": len(sc["points"]), "grid": sc["grid"], "num_rots": R, "ms": ms,
                  "ms_total": sum(ms.values()), "candidates_accepted": n_boxes, "boxes_after_nms": n_keep,
                  "what": "MinkUNetEngine.predict -> hv_cuda.forward -> hough_voting.back_project -> hough_voting.nms_per_class"}))
