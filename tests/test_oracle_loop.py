"""CPU tests of the candidate-loop oracle (oracle/candidate_loop.py): the explicit float32 restatement
(loop_numpy) against the script's own torch op sequence (loop_torch, eval_joint.py:201-263), plus
constructed cases.  The reference has no test for this loop; loop_torch is the pin."""
import numpy as np
import pytest
import torch

from oracle import candidate_loop as CL
from oracle import hv_oracle as O
from tests.helpers import small_scene

RES = 0.03


def _grids(sc, R):
    return O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(RES), R)


def _run_both(sc, R, **kw):
    go, gr, gs = _grids(sc, R)
    g1 = go.copy()
    b1, s1, c1, it1, tr = CL.loop_numpy(g1, gr, gs, sc["points"], sc["xyz"], sc["obj"], sc["class_pred"], RES,
                                        return_trace=True, **kw)
    g2 = torch.from_numpy(go.copy())
    t = torch.from_numpy
    b2, s2, c2, it2 = CL.loop_torch(g2, t(gr), t(gs), t(sc["points"]), t(sc["xyz"]), t(sc["obj"]), t(sc["class_pred"]),
                                    RES, **kw)
    return (b1, s1, c1, it1, g1, tr), (b2, s2, c2, it2, g2.numpy())


@pytest.mark.parametrize("n,G,R,seed", [(5000, 32, 4, 0), (20000, 64, 12, 1), (12000, 48, 8, 5)])
def test_numpy_restatement_matches_script_ops(n, G, R, seed):
    sc = small_scene(n, G, R, seed)
    a, b = _run_both(sc, R, thresh_high=60.0 * R / 120)
    assert a[3] == b[3] and a[3] > 3, "iteration counts differ"
    assert len(a[0]) == len(b[0])
    np.testing.assert_allclose(a[0], b[0], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[2], b[2])
    np.testing.assert_array_equal(a[4], b[4])       # the zeroed grid is identical voxel for voxel


def test_eval_separate_variant_and_thresholds():
    sc = small_scene(8000, 40, 6, 3)
    a, b = _run_both(sc, 6, thresh_high=3.0, elim_hi_inclusive=False, thresh_low=5, valid_ratio=0.1)
    assert a[3] == b[3]
    np.testing.assert_array_equal(a[4], b[4])
    np.testing.assert_array_equal(a[2], b[2])


def test_recovers_a_planted_box():
    """One object, noise-free predictions: the loop must return its box (centre, yaw mod pi/2-symmetry, class)."""
    sc = small_scene(6000, 48, 24, 7, n_objects=1, snap_yaw=True)
    go, gr, gs = _grids(sc, 24)
    boxes, scores, classes, iters = CL.loop_numpy(go, gr, gs, sc["points"], sc["xyz"], sc["obj"], sc["class_pred"], RES,
                                                  thresh_high=8.0)
    assert len(boxes) >= 1
    centre, half, yaw, cls = sc["boxes"][0]
    k = int(np.argmin(np.linalg.norm(boxes.mean(1) - centre, axis=1)))
    assert np.linalg.norm(boxes[k].mean(0) - centre) < 2.5 * RES
    assert classes[k] == cls
    assert 0.6 <= scores[k] <= 1.0


def test_empty_grid_terminates_immediately():
    go = np.zeros((8, 8, 8), np.float32)
    out = CL.loop_numpy(go, np.zeros((8, 8, 8, 2), np.float32), np.ones((8, 8, 8, 3), np.float32),
                        np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32), np.ones(4, np.float32),
                        np.zeros(4, np.int64), RES)
    assert out[0].shape == (0, 8, 3) and out[3] == 0
