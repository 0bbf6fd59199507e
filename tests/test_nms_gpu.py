"""GPU parity of the OBB IoU / per-class NMS (csrc/obb_nms.cu) against oracle/obb_nms.py on seeded random boxes."""
import numpy as np
import pytest
import torch

from oracle import obb_nms as O
from tests.test_oracle_nms import _box

pytestmark = pytest.mark.gpu


def _random_boxes(k, seed, spread):
    rng = np.random.default_rng(seed)
    boxes = np.stack([_box(rng.uniform(0, spread), rng.uniform(0, spread), rng.uniform(0.3, 1.5), rng.uniform(0.3, 1.5), rng.uniform(0, 2 * np.pi),
                           y0=rng.uniform(0, 0.3), y1=rng.uniform(0.5, 1.5)) for _ in range(k)])
    scores = rng.uniform(0.3, 1.0, k).astype(np.float32)
    scores[rng.integers(0, k, k // 5)] = np.float32(0.75)             # ties
    classes = rng.integers(0, 9, k)
    return boxes, scores, classes


@pytest.mark.parametrize("k,spread", [(1, 1.0), (40, 3.0), (300, 6.0)])
def test_iou_matrix_and_nms_match_oracle(k, spread):
    from canonicalvoting_b200 import obb
    boxes, scores, classes = _random_boxes(k, 100 + k, spread)
    m = obb.iou_matrix(torch.from_numpy(boxes).cuda(), torch.from_numpy(boxes).cuda()).cpu().numpy()
    want = np.array([[O.get_iou_obb(boxes[i], boxes[j]) for j in range(k)] for i in range(min(k, 40))])
    assert np.abs(m[:min(k, 40)] - want).max() <= 1e-12
    assert (m > 0.3).sum() > k or k == 1                                # the scene is crowded enough to suppress something
    for thr in (0.3, 0.05):
        got = obb.nms_per_class(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(classes).cuda(), 9, thr)
        assert got.cpu().tolist() == O.nms_per_class(boxes, scores, classes, 9, thr)


def test_nms_edge_cases():
    from canonicalvoting_b200 import obb
    empty = obb.nms_per_class(torch.zeros((0, 8, 3)).cuda(), torch.zeros(0).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), 9)
    assert empty.numel() == 0
    boxes, scores, classes = _random_boxes(20, 7, 2.0)
    classes[:5] = 9                                                      # background / out-of-range classes are dropped
    got = obb.nms_per_class(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(classes).cuda(), 9)
    assert got.cpu().tolist() == O.nms_per_class(boxes, scores, classes, 9, 0.3)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        obb.nms_per_class(torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(classes), 9)


@pytest.mark.parametrize("seed", [0, 3])
def test_detection_metric_on_device_ious_matches_oracle(seed):
    """canonicalvoting_b200/evaluate.py with its default IoU source (obb.iou_matrix on the device, one call per scene
    and class) against oracle/detection_metric.py (the reference's per-pair loop, utils/calc_map.py:78-168)."""
    from canonicalvoting_b200 import evaluate as E
    from oracle import detection_metric as OM
    from tests.test_oracle_map import random_eval_case
    pred_all, gt_all = random_eval_case(seed)
    gt_all = {s: [(c, b.astype(np.float32)) for c, b in v] for s, v in gt_all.items()}     # the device takes float32 boxes
    for thresh in (0.25, 0.5):
        want = OM.compute_map(pred_all, gt_all, thresh)
        got = E.compute_map(pred_all, gt_all, thresh)
        assert list(got) == list(want)
        for k in want:
            np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), np.asarray(want[k], dtype=np.float64), rtol=0, atol=1e-12, err_msg=k)
