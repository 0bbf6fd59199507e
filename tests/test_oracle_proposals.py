"""CPU tests of the proposal-sampler oracle (oracle/proposals.py): the explicit float32 restatement (sample_numpy)
against the script's own torch op sequence (sample_torch, sunrgbd/brnetcanon.py:118-161) with the random draws injected,
plus constructed cases.  The reference has no test for this module; sample_torch is the pin."""
import numpy as np
import pytest
import torch

from oracle import hv_oracle as O
from oracle import proposals as P
from tests.helpers import small_scene

RES = 0.03


def vote_case(n, G, R, seed, n_seeds=256, n_draw=96, trials=12):
    """(hv_map, hv_scale, corner0, seeds, draws): a voted scene, seeds = a random subset of its points (jittered),
    draws from the reference's distribution with a seeded CPU generator."""
    sc = small_scene(n, G, R, seed)
    hv_map, _, hv_scale = O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(RES), R)
    rng = np.random.default_rng(seed + 17)
    seeds = (sc["points"][rng.choice(n, n_seeds, replace=False)] + rng.normal(0, 0.05, (n_seeds, 3))).astype(np.float32)
    corner0 = sc["points"].min(0)
    dist, _ = P.distribution(torch.from_numpy(hv_map))
    g = torch.Generator().manual_seed(seed)
    draws = [torch.multinomial(dist, n_draw, replacement=True, generator=g) for _ in range(trials)]
    return hv_map, hv_scale, corner0, seeds, draws


@pytest.mark.parametrize("n,G,R,seed,num_proposal", [(5000, 32, 4, 0, 64), (12000, 48, 8, 5, 150), (3000, 24, 6, 2, 40)])
def test_numpy_restatement_matches_script_ops(n, G, R, seed, num_proposal):
    hv_map, hv_scale, corner0, seeds, draws = vote_case(n, G, R, seed)
    t = torch.from_numpy
    c1, s1, used1 = P.sample_torch(t(hv_map), t(hv_scale), RES, t(corner0), t(seeds), draws, num_proposal)
    c2, s2, used2, dmins = P.sample_numpy(hv_map, hv_scale, RES, corner0, seeds, [d.numpy() for d in draws], num_proposal)
    # torch.cdist may take the matmul route (|a|^2 + |b|^2 - 2ab): a draw whose nearest seed is within rounding of the
    # radius could flip; the seeded cases stay clear of it
    assert min(np.abs(d - np.float32(0.3)).min() for d in dmins[:used2]) > 1e-4
    assert used1 == used2 and used1 >= 1
    assert c1.shape == (num_proposal, 3)
    np.testing.assert_array_equal(c1.numpy(), c2)
    np.testing.assert_array_equal(s1.numpy(), s2)


def test_keep_all_when_no_draw_is_near_a_seed_and_first_argmax():
    hv_map = np.zeros((4, 5, 6), np.float32)
    hv_map[1, 3, 2] = 2.0
    hv_map[1, 4, 2] = 2.0                     # tie along y: the first maximum wins
    hv_map[2, 0, 5] = 1.0
    hv_scale = np.arange(4 * 5 * 6 * 3, dtype=np.float32).reshape(4, 5, 6, 3)
    vmax, arg = P.project_y_numpy(hv_map)
    assert arg[1, 2] == 3 and vmax[1, 2] == 2.0 and arg[0, 0] == 0
    far = np.full((3, 3), 100.0, np.float32)
    draws = [np.array([1 * 6 + 2, 2 * 6 + 5, 0], np.int64)]
    c, s, used, _ = P.sample_numpy(hv_map, hv_scale, 0.5, np.array([1.0, 2.0, 3.0], np.float32), far, draws, 2)
    assert used == 1                          # nothing within 0.3 -> every draw is kept, truncated to num_proposal
    np.testing.assert_array_equal(c, np.array([[1.5, 3.5, 4.0], [2.0, 2.0, 5.5]], np.float32))
    np.testing.assert_array_equal(s, np.stack([hv_scale[1, 3, 2], hv_scale[2, 0, 5]]))
    t = torch.from_numpy
    c2, s2, _ = P.sample_torch(t(hv_map), t(hv_scale), 0.5, torch.tensor([1.0, 2.0, 3.0]), t(far), [t(d) for d in draws], 2)
    np.testing.assert_array_equal(c2.numpy(), c)
    np.testing.assert_array_equal(s2.numpy(), s)


def test_degenerate_map_gives_uniform_distribution():
    dist, _ = P.distribution(torch.zeros(3, 4, 5))
    assert torch.equal(dist, torch.full((15,), 1e-7).sqrt())   # flat already: the sum 15 * sqrt(1e-7) is above the 1e-7 floor
    dist, _ = P.distribution(torch.full((3, 4, 5), float("nan")))
    assert torch.equal(dist, torch.ones(15))
