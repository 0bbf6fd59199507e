"""GPU parity of the vote-map proposal sampler (csrc/hv_proposals.cu, canonicalvoting_b200/proposals.py) against
oracle/proposals.py: the y-projection bit-exact, one forward() with injected draws bit-exact (locations, scales, number
of trials), and the module end to end with torch.multinomial (shape / membership properties)."""
import numpy as np
import pytest
import torch

from oracle import proposals as P
from tests.test_oracle_proposals import RES, vote_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(4, 5, 6), (33, 7, 70), (128, 128, 128), (1, 1, 1)])
def test_project_y_matches_numpy(shape):
    from canonicalvoting_b200 import proposals
    rng = np.random.default_rng(sum(shape))
    g = rng.integers(0, 6, shape).astype(np.float32)          # many ties along y: the first maximum must win
    vmax, arg = proposals.project_y(torch.from_numpy(g).cuda())
    wmax, warg = P.project_y_numpy(g)
    np.testing.assert_array_equal(vmax.cpu().numpy(), wmax)
    np.testing.assert_array_equal(arg.cpu().numpy(), warg)
    assert torch.equal(vmax, torch.from_numpy(g).cuda().max(1)[0])


@pytest.mark.parametrize("n,G,R,seed,num_proposal,n_draw", [(5000, 32, 4, 0, 64, 96), (12000, 48, 8, 5, 150, 96), (20000, 64, 12, 1, 1500, 2250)])
def test_trials_match_oracle_with_injected_draws(n, G, R, seed, num_proposal, n_draw):
    from canonicalvoting_b200 import proposals
    hv_map, hv_scale, corner0, seeds, draws = vote_case(n, G, R, seed, n_draw=n_draw, trials=40)
    wc, ws, used, dmins = P.sample_numpy(hv_map, hv_scale, RES, corner0, seeds, [d.numpy() for d in draws], num_proposal)
    _, arg = proposals.project_y(torch.from_numpy(hv_map).cuda())
    state = proposals._Trial(num_proposal, n_draw, torch.device("cuda"))
    gs = torch.from_numpy(hv_scale).cuda()
    trials, cnt = 0, 0
    while cnt < num_proposal:
        proposals.append_proposals(state, draws[trials].cuda(), arg, gs, RES, corner0, torch.from_numpy(seeds).cuda())
        trials += 1
        cnt = int(state.count.item())
    assert trials == used
    np.testing.assert_array_equal(state.loc.cpu().numpy(), wc)
    np.testing.assert_array_equal(state.scale.cpu().numpy(), ws)


def test_module_end_to_end():
    """HoughVotingModule.forward (sunrgbd/brnetcanon.py:115-162) on a voted scene: vote through hv_cuda.forward(...,
    corners), draw with torch.multinomial; every returned location is a (x, argmax_y, z) cell centre of the vote map
    carrying that cell's voted scale, and lies within 0.3 m of a seed whenever any draw of its trial did."""
    from canonicalvoting_b200 import proposals
    from tests.helpers import small_scene
    sc = small_scene(5000, 32, 4, 0)
    t = lambda a: torch.from_numpy(a).cuda()
    pts = t(sc["points"])
    corners = torch.stack([pts.min(0)[0], pts.max(0)[0]])
    seeds = pts[::20].contiguous()
    mod = proposals.HoughVotingModule(res=RES, num_rots=4, num_proposal=128)
    torch.manual_seed(0)
    cand, probs, scales = mod(pts, t(sc["xyz"]), t(sc["scale"]), t(sc["obj"]), corners, seeds)
    assert cand.shape == (128, 3) and scales.shape == (128, 3) and probs.shape == (128,) and float(probs.abs().sum()) == 0.0
    # the vote op accumulates with float reductions in arbitrary order, so a second vote of the same scene matches the
    # module's own to rounding only: membership is checked with the north_star tolerance (1e-4 relative)
    import hv_cuda
    hv_map, _, hv_scale = hv_cuda.forward(pts, t(sc["xyz"]), t(sc["scale"]), t(sc["obj"]), mod.res, mod.num_rots, corners)
    cell = torch.round((cand - corners[0]) / RES).long()
    assert bool(((cell >= 0) & (cell < torch.tensor(hv_map.shape, device="cuda"))).all())
    at_cell = hv_map[cell[:, 0], cell[:, 1], cell[:, 2]]
    col_max = hv_map.max(1)[0][cell[:, 0], cell[:, 2]]
    assert bool((at_cell >= col_max * (1 - 1e-4) - 1e-6).all())
    torch.testing.assert_close(scales, hv_scale[cell[:, 0], cell[:, 1], cell[:, 2]], rtol=1e-4, atol=1e-6)
    d = torch.cdist(cand, seeds).min(-1)[0]
    assert float(d.max()) < 0.3 + 1e-5
