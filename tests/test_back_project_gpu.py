"""GPU parity tests of `hough_voting.back_project` (csrc/bp_loop.cu through the C ABI) against the CPU oracle
of the reference's candidate loop (oracle/candidate_loop.py; eval_joint.py:195-263).
Bar: the per-iteration integer decisions (peak voxel, points inside, confident points, accept/reject),
class ids and the zeroed grid are bit-exact; boxes / scores within 1e-6 absolute (same float32 op order)."""
import numpy as np
import pytest
import torch

from oracle import candidate_loop as CL
from oracle import hv_oracle as O
from tests.helpers import small_scene

pytestmark = pytest.mark.gpu
RES = 0.03


def _case(n, G, R, seed, **kw):
    import hough_voting
    sc = small_scene(n, G, R, seed)
    go, gr, gs = O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(RES), R)
    ref_grid = go.copy()
    wb, ws, wc, wit, wtr = CL.loop_numpy(ref_grid, gr, gs, sc["points"], sc["xyz"], sc["obj"], sc["class_pred"], RES,
                                         return_trace=True, **kw)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dgo = d(go)
    boxes, scores, classes, trace, iters = hough_voting.back_project(
        dgo, d(gr), d(gs), d(sc["points"]), d(sc["xyz"]), d(sc["obj"]), d(sc["class_pred"]), RES, return_trace=True, **kw)
    assert iters == wit, "iteration count differs: %d vs oracle %d" % (iters, wit)
    np.testing.assert_array_equal(trace.cpu().numpy(), np.asarray(wtr, np.int32).reshape(-1, 4))
    np.testing.assert_array_equal(dgo.cpu().numpy(), ref_grid)
    np.testing.assert_array_equal(classes.cpu().numpy(), wc)
    np.testing.assert_allclose(boxes.cpu().numpy(), wb, rtol=0, atol=1e-6)
    np.testing.assert_allclose(scores.cpu().numpy(), ws, rtol=0, atol=0)
    return len(wb), wit


@pytest.mark.parametrize("n,G,R,seed", [(5000, 32, 4, 0), (20000, 64, 12, 1), (777, 20, 7, 2), (50000, 128, 12, 0)],
                         ids=["C1", "mid", "ragged", "C2"])
def test_matches_oracle(n, G, R, seed):
    k, it = _case(n, G, R, seed, thresh_high=60.0 * R / 120)
    assert it > 0


def test_eval_separate_variant():
    _case(8000, 40, 6, 3, thresh_high=3.0, elim_hi_inclusive=0, thresh_low=5, valid_ratio=0.1)


def test_C5_full_size_after_cuda_vote():
    """BASELINE configs[4]: CUDA vote (200k points, 256^3, R=24) followed by the loop; the oracle runs on the
    grids the CUDA vote produced, so this checks the loop alone at full size."""
    import hough_voting
    import hv_cuda
    sc = small_scene(200_000, 256, 24, 0)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    p, x, s, o = d(sc["points"]), d(sc["xyz"]), d(sc["scale"]), d(sc["obj"])
    go, gr, gs = hv_cuda.forward(p, x, s, o, torch.tensor(RES).cuda(), torch.tensor(24, dtype=torch.int32).cuda())
    ref_grid = go.cpu().numpy().copy()
    wb, ws, wc, wit, wtr = CL.loop_numpy(ref_grid, gr.cpu().numpy(), gs.cpu().numpy(), sc["points"], sc["xyz"], sc["obj"],
                                         sc["class_pred"], RES, return_trace=True, thresh_high=12.0)
    boxes, scores, classes, trace, iters = hough_voting.back_project(go, gr, gs, p, x, o, d(sc["class_pred"]), RES,
                                                                     return_trace=True, thresh_high=12.0)
    assert iters == wit and iters > 10
    np.testing.assert_array_equal(trace.cpu().numpy(), np.asarray(wtr, np.int32).reshape(-1, 4))
    np.testing.assert_array_equal(go.cpu().numpy(), ref_grid)
    np.testing.assert_array_equal(classes.cpu().numpy(), wc)
    np.testing.assert_allclose(boxes.cpu().numpy(), wb, rtol=0, atol=1e-6)
    # size-independent properties: scores are probabilities of points inside; the loop left no peak above the threshold
    assert float(go.max()) < 12.0
    assert ((scores >= 0) & (scores <= 1)).all()


def test_errors_and_empty():
    import hough_voting
    z = lambda *s: torch.zeros(*s, device="cuda")
    out = hough_voting.back_project(z(8, 8, 8), z(8, 8, 8, 2), z(8, 8, 8, 3) + 1, z(4, 3), z(4, 3), z(4) + 1,
                                    torch.zeros(4, dtype=torch.int64, device="cuda"), RES)
    assert out[0].shape == (0, 8, 3) and out[1].numel() == 0
    with pytest.raises(RuntimeError, match="grid_obj must be a CUDA tensor"):
        hough_voting.back_project(torch.zeros(8, 8, 8), z(8, 8, 8, 2), z(8, 8, 8, 3), z(4, 3), z(4, 3), z(4),
                                  torch.zeros(4, dtype=torch.int64, device="cuda"), RES)
    with pytest.raises(RuntimeError, match="thresh_high must be > 0"):
        hough_voting.back_project(z(8, 8, 8), z(8, 8, 8, 2), z(8, 8, 8, 3), z(4, 3), z(4, 3), z(4),
                                  torch.zeros(4, dtype=torch.int64, device="cuda"), RES, thresh_high=0.0)
