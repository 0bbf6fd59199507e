"""GPU parity tests of the vote op, through the reference-facing boundary (`hv_cuda.forward /
backward`, which call the C ABI of libcvb200.so) against
  (1) the CPU oracle (oracle/hv_oracle.c) fed the device's own cos/sin table,
  (2) the UNMODIFIED reference kernel built for sm_100a (oracle/_ref/hv_cuda_ref.so), when present,
  (3) size-independent properties at BASELINE.json's full sizes.
Bar: integer voxel indices bit-exact; float maps within 1e-4 relative (north_star)."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_grid_close, small_scene

pytestmark = pytest.mark.gpu

RES = 0.03


def _dev(sc):
    t = lambda k: torch.from_numpy(sc[k]).cuda()
    return t("points"), t("xyz"), t("scale"), t("obj")


def _scalars(num_rots):
    return torch.tensor(RES, dtype=torch.float32).cuda(), torch.tensor(num_rots, dtype=torch.int32).cuda()


def _ref():
    from oracle import build_ref
    return build_ref.load_ref()


CASES = [
    ("C1", dict(n=5000, grid=32, num_rots=4, seed=0)),                       # BASELINE configs[0]
    ("C1-uniform", dict(n=5000, grid=32, num_rots=4, seed=1, uniform=True)),
    ("ragged", dict(n=777, grid=20, num_rots=7, seed=2)),
    ("R120", dict(n=2000, grid=24, num_rots=120, seed=3)),                  # reference default R
    ("C2", dict(n=50000, grid=128, num_rots=12, seed=0)),                    # BASELINE configs[1] vote half
    ("odd-grid", dict(n=6000, grid=37, num_rots=9, seed=5)),                 # grid edge not a multiple of the 8-voxel tile
    ("C5", dict(n=200000, grid=256, num_rots=24, seed=0)),                   # BASELINE configs[4] vote half
]


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
def test_forward_matches_oracle(name, kw):
    import hv_cuda
    from canonicalvoting_b200 import hv_cuda as H
    from oracle import hv_oracle as O
    sc = small_scene(kw["n"], kw["grid"], kw["num_rots"], kw["seed"], uniform=kw.get("uniform", False))
    R = kw["num_rots"]
    p, x, s, o = _dev(sc)
    res_t, rots_t = _scalars(R)
    go, gr, gs = hv_cuda.forward(p, x, s, o, res_t, rots_t)
    corner, dims = O.grid_dims(sc["points"], np.float32(RES))
    assert tuple(go.shape) == tuple(dims) and tuple(gr.shape) == tuple(dims) + (2,) and tuple(gs.shape) == tuple(dims) + (3,)
    # bit-exact integer part
    ct, st = H.theta_table(R)
    theta = (ct.cpu().numpy(), st.cpu().numpy())
    idx = H.vote_indices(p, x, s, RES, R, corner, dims).cpu().numpy()
    ogo, ogr, ogs, oidx = O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(RES), R,
                                    theta=theta, return_votes=True)
    assert np.array_equal(idx, oidx), "vote voxel indices differ from the oracle"
    assert ((go.cpu().numpy() != 0) == (ogo != 0)).all(), "support of grid_obj differs"
    assert_grid_close(go.cpu().numpy(), ogo, what="grid_obj")
    assert_grid_close(gr.cpu().numpy(), ogr, what="grid_rot", atol_frac=2e-6)
    assert_grid_close(gs.cpu().numpy(), ogs, what="grid_scale")


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
def test_forward_backward_match_reference_build(name, kw):
    ref = _ref()
    if ref is None:
        pytest.skip("oracle/_ref/hv_cuda_ref.so not built")
    import hv_cuda
    sc = small_scene(kw["n"], kw["grid"], kw["num_rots"], kw["seed"], uniform=kw.get("uniform", False))
    p, x, s, o = _dev(sc)
    res_t, rots_t = _scalars(kw["num_rots"])
    want = ref.forward(p, x, s, o, res_t, rots_t)
    got = hv_cuda.forward(p, x, s, o, res_t, rots_t)
    torch.cuda.synchronize()
    for g, w, nm in zip(got, want, ("grid_obj", "grid_rot", "grid_scale")):
        assert g.shape == w.shape and g.dtype == w.dtype
        assert_grid_close(g.cpu().numpy(), w.cpu().numpy(), what=nm, atol_frac=2e-6)
    assert torch.equal(got[0] != 0, want[0] != 0), "support of grid_obj differs from the reference kernel"
    grad = torch.randn_like(want[0])
    wb = ref.backward(grad, p, x, s, o, res_t, rots_t)
    gb = hv_cuda.backward(grad, p, x, s, o, res_t, rots_t)
    for g, w, nm in zip(gb, wb, ("d_xyz", "d_scale", "d_obj")):
        assert g.shape == w.shape
        assert_grid_close(g.cpu().numpy(), w.cpu().numpy(), rtol=1e-4, atol_frac=1e-5, what=nm)


def test_backward_matches_oracle():
    import hv_cuda
    from canonicalvoting_b200 import hv_cuda as H
    from oracle import hv_oracle as O
    sc = small_scene(4000, 32, 8, seed=11)
    p, x, s, o = _dev(sc)
    res_t, rots_t = _scalars(8)
    grad = torch.randn(32, 32, 32, device="cuda")
    d_xyz, d_scale, d_obj = hv_cuda.backward(grad, p, x, s, o, res_t, rots_t)
    ct, st = H.theta_table(8)
    w = O.backward(grad.cpu().numpy(), sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(RES), 8,
                   theta=(ct.cpu().numpy(), st.cpu().numpy()))
    assert_grid_close(d_xyz.cpu().numpy(), w[0], rtol=1e-4, atol_frac=1e-5, what="d_xyz")
    assert_grid_close(d_scale.cpu().numpy(), w[1], rtol=1e-4, atol_frac=1e-5, what="d_scale")
    assert_grid_close(d_obj.cpu().numpy(), w[2], rtol=1e-4, atol_frac=1e-5, what="d_obj")


def test_autograd_glue_like_the_reference_scripts():
    # HoughVoting / HVFunction as in train_joint.py:22-56, including backward through grid_obj
    from hough_voting import HoughVoting
    sc = small_scene(1500, 16, 6, seed=4)
    p, x, s, o = _dev(sc)
    x.requires_grad_(True); s.requires_grad_(True); o.requires_grad_(True)
    hv = HoughVoting(RES, 6)
    grid_obj, grid_rot, grid_scale = hv(p, x, s, o)
    (grid_obj * torch.linspace(0, 1, grid_obj.numel(), device="cuda").view_as(grid_obj)).sum().backward()
    assert x.grad.shape == x.shape and s.grad.shape == s.shape and o.grad.shape == o.shape
    assert torch.isfinite(x.grad).all() and o.grad.abs().sum() > 0


def test_full_size_properties_C5():
    """BASELINE configs[4] vote half (200k points, 256^3, R=24): size-independent properties."""
    import hv_cuda
    from canonicalvoting_b200 import hv_cuda as H
    sc = small_scene(200_000, 256, 24, seed=0)
    p, x, s, o = _dev(sc)
    res_t, rots_t = _scalars(24)
    go, gr, gs = hv_cuda.forward(p, x, s, o, res_t, rots_t)
    assert tuple(go.shape) == (256, 256, 256)
    corner, _, dims = H.grid_dims(p, RES)
    assert dims == (256, 256, 256) and corner == (0.0, 0.0, 0.0)
    idx = H.vote_indices(p, x, s, RES, 24, corner, dims)
    kept = (idx[:, :, 0] >= 0)
    # checksum: the 8 trilinear weights of a kept vote sum to 1
    want = (kept.sum(1).double() * o.double()).sum().item()
    assert abs(go.double().sum().item() - want) <= 1e-4 * want
    # support: every touched voxel lies in the 2x2x2 neighbourhood of some kept vote
    touched = torch.zeros(256 ** 3, dtype=torch.bool, device="cuda")
    f = idx[kept].long()
    for a in (0, 1):
        for b in (0, 1):
            for d in (0, 1):
                touched[((f[:, 0] + a) * 256 + f[:, 1] + b) * 256 + f[:, 2] + d] = True
    assert not (go.flatten() != 0)[~touched].any()
    # linearity in objectness (exact scaling by a power of two, up to atomic ordering)
    go2, gr2, gs2 = hv_cuda.forward(p, x, s, o * 2, res_t, rots_t)
    assert_grid_close(go2.cpu().numpy(), 2 * go.cpu().numpy(), what="linearity grid_obj")
    # normalised channels: grid_scale is a convex combination of per-point scales
    nz = go > 1e-3
    assert gs[nz].min() >= s.min() * (1 - 1e-4) and gs[nz].max() <= s.max() * (1 + 1e-4)
    assert (gr[nz].norm(dim=-1) <= 1 + 1e-4).all()
    # idempotence: the write-out re-zeroes what the scatter touched, so the cached workspace is all-zero
    # again and a second / third call reproduces the first
    for _ in range(2):
        go3, gr3, gs3 = hv_cuda.forward(p, x, s, o, res_t, rots_t)
        assert_grid_close(go3.cpu().numpy(), go.cpu().numpy(), what="idempotence obj")
        assert_grid_close(gs3.cpu().numpy(), gs.cpu().numpy(), what="idempotence scale")
    for w in H._work_cache.values():
        assert not w.any(), "workspace not re-zeroed"


def test_edge_cases():
    import hv_cuda
    res_t, rots_t = _scalars(5)
    # single point: dims (1,1,1), every vote fails g < dim-1 = 0 -> all-zero grids
    one = torch.tensor([[0.3, 0.6, 0.9]], device="cuda")
    go, gr, gs = hv_cuda.forward(one, torch.zeros(1, 3, device="cuda"), torch.ones(1, 3, device="cuda"),
                                 torch.ones(1, device="cuda"), res_t, rots_t)
    assert tuple(go.shape) == (1, 1, 1) and go.item() == 0 and not gr.any() and not gs.any()
    # all votes out of bounds (huge offsets)
    sc = small_scene(300, 10, 5, seed=9)
    p, x, s, o = _dev(sc)
    go, gr, gs = hv_cuda.forward(p, x + 100.0, s, o, res_t, rots_t)
    assert not go.any() and not gr.any() and not gs.any()
    # duplicated points double the map
    g1 = hv_cuda.forward(p, x, s, o, res_t, rots_t)[0]
    g2 = hv_cuda.forward(torch.cat([p, p]), torch.cat([x, x]), torch.cat([s, s]), torch.cat([o, o]), res_t, rots_t)[0]
    assert_grid_close(g2.cpu().numpy(), 2 * g1.cpu().numpy(), what="duplicates")
    # reference error behaviour (hv_cuda.cpp:26-28)
    with pytest.raises(RuntimeError, match="xyz_labels must be contiguous"):
        hv_cuda.forward(p, torch.zeros(3, 300, device="cuda").t(), s, o, res_t, rots_t)
    with pytest.raises(RuntimeError, match="obj_labels must be a CUDA tensor"):
        hv_cuda.forward(p, x, s, o.cpu(), res_t, rots_t)
    # corners override (sunrgbd/brnetcanon.py:99 call form) == default geometry when corners = min/max
    corners = torch.stack([p.min(0)[0], p.max(0)[0]])
    g3 = hv_cuda.forward(p, x, s, o, res_t, rots_t, corners)[0]
    assert_grid_close(g3.cpu().numpy(), g1.cpu().numpy(), what="corners override")


def test_non_default_stream_and_float64():
    import hv_cuda
    sc = small_scene(2000, 24, 6, seed=6)
    p, x, s, o = _dev(sc)
    res_t, rots_t = _scalars(6)
    base = hv_cuda.forward(p, x, s, o, res_t, rots_t)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        other = hv_cuda.forward(p, x, s, o, res_t, rots_t)
    st.synchronize()
    for a, b in zip(base, other):
        assert_grid_close(b.cpu().numpy(), a.cpu().numpy(), what="stream")
    d = hv_cuda.forward(p.double(), x.double(), s.double(), o.double(), res_t, rots_t)
    assert d[0].dtype == torch.float64
    assert_grid_close(d[0].cpu().numpy(), base[0].cpu().numpy(), what="float64 io")


def test_scalar_tensors_are_not_confused_by_address_reuse():
    """res / num_rots arrive as 0-dim device tensors; their host copies are cached per tensor OBJECT."""
    import hv_cuda
    sc = small_scene(800, 16, 4, seed=5)
    p, x, s, o = _dev(sc)
    res_t = torch.tensor(RES, dtype=torch.float32).cuda()
    sums = []
    for R in (4, 8, 4, 12):
        rots_t = torch.tensor(R, dtype=torch.int32).cuda()      # very likely the same address every time
        go, _, _ = hv_cuda.forward(p, x, s, o, res_t, rots_t)
        sums.append(float(go.sum()))
        del rots_t
    assert abs(sums[0] - sums[2]) <= 1e-3 * sums[0]
    assert sums[1] > 1.5 * sums[0] and sums[3] > 2.2 * sums[0]
    rots_t = torch.tensor(4, dtype=torch.int32).cuda()
    a = float(hv_cuda.forward(p, x, s, o, res_t, rots_t)[0].sum())
    rots_t.fill_(8)                                             # in-place update bumps the version counter
    b = float(hv_cuda.forward(p, x, s, o, res_t, rots_t)[0].sum())
    assert b > 1.5 * a
