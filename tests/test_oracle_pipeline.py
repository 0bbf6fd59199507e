"""CPU end-to-end check of the oracle chain on planted objects: vote (oracle/hv_oracle.c) -> candidate loop with the LCC
back-projection check (oracle/candidate_loop.py) -> per-class OBB NMS (oracle/obb_nms.py) -> detection metric
(oracle/detection_metric.py) against the ground-truth boxes of the synthetic scene.  The synthetic per-point predictions are the
planted objects' LCC coordinates + noise, so every box the loop accepts must be one of them (precision 1 at IoU 0.5), in the
yaw convention of eval_joint.py:213-215,299.  This is what ties the four oracles -- each pinned separately -- together."""
import numpy as np
import pytest

from canonicalvoting_b200 import evaluate as E
from canonicalvoting_b200 import synthetic
from oracle import detection_metric as OM
from oracle import candidate_loop as CL
from oracle import hv_oracle as O
from oracle import obb_nms as ON

RES = 0.03


@pytest.mark.parametrize("n,G,R,seed", [(20000, 64, 12, 1), (20000, 64, 24, 2), (30000, 96, 12, 4)])
def test_accepted_boxes_are_the_planted_objects(n, G, R, seed):
    sc = synthetic.make_scene(n, G, R, seed=seed)
    go, gr, gs = O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(RES), R)
    boxes, scores, classes, iters = CL.loop_numpy(go.copy(), gr, gs, sc["points"], sc["xyz"], sc["obj"], sc["class_pred"], RES,
                                                  thresh_high=60.0 * R / 120)
    keep = np.asarray(ON.nms_per_class(boxes, scores, classes, 9, 0.3), dtype=np.int64)
    dets = E.scene_detections(boxes, scores, classes, keep)
    gt = [(E.CATEGORIES[k], E.gt_box(c[0], c[1], c[2], yaw, h[0], h[1], h[2])) for c, h, yaw, k in sc["boxes"]]
    assert len(dets) >= 3 and iters > len(dets)
    for thresh in (0.25, 0.5):
        rec, prec, ap = OM.evaluate({"scene": dets}, {"scene": gt}, thresh)
        n_tp = sum(int(round(rec[c][-1] * sum(1 for g in gt if g[0] == c))) for c in rec if not np.isscalar(rec[c]))
        assert n_tp == len(dets), "an accepted box is not a planted object at IoU %.2f" % thresh
        for c in prec:
            if not np.isscalar(prec[c]):
                assert prec[c][-1] == 1.0
    ret = OM.compute_map({"scene": dets}, {"scene": gt}, 0.5)
    assert ret["mAP"] >= len(dets) / len(gt) * 0.5          # each found object contributes its class's recall
