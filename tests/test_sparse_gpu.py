"""GPU parity tests of the sparse-voxel stack (csrc/sparse_coords.cu, csrc/sparse_conv*.cu through the C ABI and
the MinkowskiEngine-compatible modules) against oracle/sparse_oracle.py (which tests/test_oracle_sparse.py pins
to dense conv3d).  fp32 CUDA-core path: 1e-5 relative; coordinate maps: exact; gradients: against torch
autograd through the oracle."""
import numpy as np
import pytest
import torch

from oracle import sparse_oracle as SO

pytestmark = pytest.mark.gpu


def _scene(n=3000, G=24, batch=2, cin=3, seed=0, negative=False):
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(batch):
        lin = torch.randperm(G ** 3, generator=g)[:n]
        rows.append(torch.stack([torch.full_like(lin, b), lin // (G * G), (lin // G) % G, lin % G], 1))
    coords = torch.cat(rows).int()
    if negative:
        coords[:, 1:] -= G // 2
    feats = torch.randn(coords.shape[0], cin, generator=g)
    return coords, feats


def _close(a, b, rtol=1e-5, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = max(1.0, float(b.abs().max()))
    err = float((a - b).abs().max())
    assert a.shape == b.shape and err <= rtol * scale, "%s: max abs err %.3e (scale %.3e)" % (what, err, scale)


@pytest.mark.parametrize("negative", [False, True])
def test_coordinate_maps_match_oracle(negative):
    import MinkowskiEngine as ME
    coords, feats = _scene(negative=negative)
    st = ME.SparseTensor(feats, coords, device="cuda")
    cm = st.coordinate_manager
    d = cm.down(1)
    coarse, parent, koff = SO.coarse_coords(coords, 2)
    assert torch.equal(cm.levels[2].coords.cpu(), coarse)          # numbered by first child: deterministic
    assert torch.equal(d["parent"].cpu().long(), parent) and torch.equal(d["koff"].cpu().long(), koff)
    ch = d["children"].cpu()
    for i in range(0, len(coords), 97):
        assert ch[parent[i], koff[i]] == i
    assert int((ch >= 0).sum()) == len(coords)
    nbr = cm.kernel_map(1, 3).cpu()
    index = {tuple(c): i for i, c in enumerate(coords.tolist())}
    for o in range(0, len(coords), 211):
        c = coords[o].tolist()
        for k in range(27):
            ix, iy, iz = k % 3 - 1, (k // 3) % 3 - 1, k // 9 - 1
            assert nbr[o, k] == index.get((c[0], c[1] + ix, c[2] + iy, c[3] + iz), -1)
    # second level: stride-4 coordinates from the stride-2 level, kernel map with step 2
    cm.down(2)
    c4, _, _ = SO.coarse_coords(coarse, 4)
    assert torch.equal(cm.levels[4].coords.cpu(), c4)
    nbr2 = cm.kernel_map(2, 3).cpu()
    index2 = {tuple(c): i for i, c in enumerate(coarse.tolist())}
    for o in range(0, len(coarse), 53):
        c = coarse[o].tolist()
        for k in (0, 5, 13, 22, 26):
            ix, iy, iz = k % 3 - 1, (k // 3) % 3 - 1, k // 9 - 1
            assert nbr2[o, k] == index2.get((c[0], c[1] + 2 * ix, c[2] + 2 * iy, c[3] + 2 * iz), -1)


@pytest.mark.parametrize("K,cin,cout", [(3, 3, 32), (5, 3, 32), (3, 32, 64), (3, 96, 96), (3, 40, 24)])
def test_conv_same_forward_backward(K, cin, cout):
    import MinkowskiEngine as ME
    coords, feats = _scene(n=2000, G=20, cin=cin, seed=K + cin)
    conv = ME.MinkowskiConvolution(cin, cout, kernel_size=K, bias=True, dimension=3).cuda()
    x = feats.cuda().requires_grad_(True)
    y = conv(ME.SparseTensor(x, coords, device="cuda")).F
    xo = feats.clone().double().requires_grad_(True)
    wo = conv.kernel.detach().cpu().double().requires_grad_(True)
    bo = conv.bias.detach().cpu().double().requires_grad_(True)
    yo = SO.conv_same(coords, xo, wo, K, 1, bo)
    _close(y, yo, what="forward")
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
    y.backward(gy.cuda())
    yo.backward(gy.double())
    _close(x.grad, xo.grad, what="dX")
    _close(conv.kernel.grad, wo.grad, rtol=2e-5, what="dW")
    _close(conv.bias.grad, bo.grad, rtol=2e-5, what="dbias")


def test_conv_down_up_forward_backward():
    import MinkowskiEngine as ME
    coords, feats = _scene(n=2500, G=22, cin=32, seed=9, negative=True)
    down = ME.MinkowskiConvolution(32, 48, kernel_size=2, stride=2, dimension=3).cuda()
    up = ME.MinkowskiConvolutionTranspose(48, 24, kernel_size=2, stride=2, dimension=3).cuda()
    x = feats.cuda().requires_grad_(True)
    s0 = ME.SparseTensor(x, coords, device="cuda")
    s1 = down(s0)
    s2 = up(s1)
    assert s1.tensor_stride == 2 and s2.tensor_stride == 1 and s2.F.shape == (len(coords), 24)
    xo = feats.clone().double().requires_grad_(True)
    wd = down.kernel.detach().cpu().double().requires_grad_(True)
    wu = up.kernel.detach().cpu().double().requires_grad_(True)
    coarse, f1 = SO.conv_down(coords, xo, wd, 1)
    f2 = SO.conv_up(coords, f1, wu, 2)
    assert torch.equal(s1.C.cpu(), coarse)
    _close(s1.F, f1, what="down forward")
    _close(s2.F, f2, what="up forward")
    gy = torch.randn(f2.shape, generator=torch.Generator().manual_seed(2))
    s2.F.backward(gy.cuda())
    f2.backward(gy.double())
    _close(x.grad, xo.grad, what="dX")
    _close(down.kernel.grad, wd.grad, rtol=2e-5, what="dW down")
    _close(up.kernel.grad, wu.grad, rtol=2e-5, what="dW up")


def test_minkunet_forward_matches_oracle_net_and_backward_runs():
    from canonicalvoting_b200 import sparse as ME
    from canonicalvoting_b200.minkunet import MinkUNet14A
    torch.manual_seed(0)
    coords, feats = _scene(n=1500, G=40, batch=2, cin=3, seed=3)
    model = MinkUNet14A(3, 20).cuda().eval()
    with torch.no_grad():
        for m in model.modules():                       # non-trivial BN statistics
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
        y = model(ME.SparseTensor(feats, coords, device="cuda")).F
    yo = SO.OracleNet(model).forward(coords, feats)
    assert y.shape == (len(coords), 20)
    _close(y, yo, rtol=2e-4, what="MinkUNet14A eval forward")
    model.train()
    out = model(ME.SparseTensor(feats, coords, device="cuda")).F
    out.square().mean().backward()
    for n_, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n_


def test_reference_row_order_and_errors():
    import MinkowskiEngine as ME
    coords, feats = _scene(n=500, G=12, batch=1)
    st = ME.SparseTensor(feats, coords, device="cuda")
    assert torch.equal(st.C.cpu(), coords) and torch.equal(st.F.cpu(), feats)      # input order preserved
    up = ME.MinkowskiConvolutionTranspose(3, 4, kernel_size=2, stride=2, dimension=3).cuda()
    with pytest.raises(RuntimeError, match="never created"):
        up(ME.SparseTensor(feats, coordinate_manager=st.coordinate_manager, tensor_stride=2))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ME.SparseTensor(feats, coords, device="cpu")


@pytest.mark.parametrize("cin,cout", [(32, 32), (32, 64), (64, 64), (96, 96), (128, 96), (128, 128), (256, 256), (384, 256), (64, 16)])
def test_tensor_core_conv_matches_fp32_path(cin, cout):
    """tcgen05 kind::tf32 implicit GEMM vs the exact-fp32 CUDA-core kernel on the same tables: TF32 rounds the
    inputs to 10 mantissa bits, accumulation is fp32 -> 3e-3 of the output scale."""
    from canonicalvoting_b200 import sparse as ME
    from canonicalvoting_b200.sparse.functional import conv_table_forward
    coords, _ = _scene(n=3000, G=20, batch=2, seed=cin + cout)
    n = len(coords)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, cin, generator=g).cuda()
    st = ME.SparseTensor(x, coords, device="cuda")
    cm = st.coordinate_manager
    d = cm.down(1)
    nc = cm.levels[2].n
    for name, table, n_in in (("same3", cm.kernel_map(1, 3), n), ("down", d["children"], n), ("up", d["up_table"], nc)):
        k3 = table.shape[1]
        xin = torch.randn(n_in, cin, generator=g).cuda()
        w = (torch.randn(k3, cin, cout, generator=g) / (cin * 4) ** 0.5).cuda()
        b = torch.randn(1, cout, generator=g).cuda() if name == "same3" else None
        ref = conv_table_forward(xin, w, table, b, mode="fp32")
        got = conv_table_forward(xin, w, table, b, mode="tf32")
        torch.cuda.synchronize()
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        assert err <= 3e-3 * scale, "%s cin=%d cout=%d: max err %.3e vs scale %.3e" % (name, cin, cout, err, scale)
        assert err > 0 or scale == 0          # it really is a different arithmetic path


def test_tensor_core_tail_tile_and_empty_rows():
    from canonicalvoting_b200.sparse.functional import conv_table_forward
    g = torch.Generator().manual_seed(3)
    n_in, n_out, cin, cout = 500, 130, 64, 32            # 130 rows: a full tile + a 2-row tail tile
    table = torch.randint(-1, n_in, (n_out, 27), generator=g, dtype=torch.int64).int()
    table[5] = -1                                          # a row without any neighbour -> exact zeros
    table[:, 20:] = -1                                     # offsets nobody uses are skipped
    x = torch.randn(n_in, cin, generator=g).cuda()
    w = torch.randn(27, cin, cout, generator=g).cuda() * 0.1
    ref = conv_table_forward(x, w, table.cuda(), None, mode="fp32")
    got = conv_table_forward(x, w, table.cuda(), None, mode="tf32")
    assert float((got - ref).abs().max()) <= 3e-3 * float(ref.abs().max())
    assert float(got[5].abs().max()) == 0.0


def test_engine_matches_module_path_and_oracle():
    """MinkUNetEngine (fused program, BN folded, TF32 tensor cores) vs the module-by-module fp32 path and the CPU oracle."""
    from canonicalvoting_b200 import sparse as ME
    from canonicalvoting_b200.engine import MinkUNetEngine
    from canonicalvoting_b200.minkunet import MinkUNet34C, decode_heads
    torch.manual_seed(1)
    coords, feats = _scene(n=4000, G=48, batch=2, cin=3, seed=11, negative=True)
    model = MinkUNet34C(3, 64).cuda().eval()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.05); m.running_var.uniform_(0.8, 1.2); m.weight.uniform_(0.8, 1.2); m.bias.normal_(0, 0.05)
        ME.set_forward_mode("fp32")
        ref = model(ME.SparseTensor(feats, coords, device="cuda")).F
    eng = MinkUNetEngine(model)
    got = eng(coords.cuda(), feats.cuda())
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    assert got.shape == ref.shape == (len(coords), 64)
    assert err <= 6e-3 * scale, "engine vs fp32 module path: max err %.3e, scale %.3e" % (err, scale)
    # head decode kernel == the torch ops of eval_joint.py:173-190 on the same features
    xyz, sc, cls, prob = eng.decode(got)
    wxyz, wsc, wcls, wprob = decode_heads(got)
    assert torch.equal(cls, wcls)
    torch.testing.assert_close(xyz, wxyz, rtol=0, atol=0)
    torch.testing.assert_close(sc, wsc, rtol=1e-6, atol=0)
    torch.testing.assert_close(prob, wprob, rtol=1e-5, atol=1e-7)
    # decode + scan_points in one kernel: points == coords[:, 1:] * res exactly (eval_joint.py:193)
    out5 = eng.decode(got, coords.cuda(), 0.03)
    assert torch.equal(out5[0], xyz) and torch.equal(out5[3], prob)
    assert torch.equal(out5[4], coords.cuda()[:, 1:].float() * 0.03)
    # stem variants: 4-channel gather inside the convolution kernel (default) vs im2col + product
    eng2 = MinkUNetEngine(model)
    eng2.stem_gather4 = False
    eng2.refresh()
    alt = eng2(coords.cuda(), feats.cuda())
    torch.cuda.synchronize()
    assert float((alt - got).abs().max()) <= 2e-3 * scale
    # second call on another scene re-uses the packed weights
    coords2, feats2 = _scene(n=1000, G=30, batch=1, cin=3, seed=12)
    out2 = eng(coords2.cuda(), feats2.cuda())
    assert out2.shape == (len(coords2), 64) and torch.isfinite(out2).all()


def test_tensor_core_conv_variants_match_oracle():
    """The persistent tcgen05 kernel in every launch variant (whole tiles only / tiles split into k-pieces reduced in-kernel,
    with and without programmatic dependent launch, host-side or DEVICE-side row count) against the CPU oracle of the
    primitive (oracle/sparse_oracle.py conv_table, float64): TF32 operands (10 mantissa bits) with fp32 accumulation -> 3e-3
    of the output scale; the variants among themselves differ only by the association of the fp32 sums (5e-5)."""
    from canonicalvoting_b200 import _lib
    from canonicalvoting_b200.sparse.coords import _ptr, _stream
    from canonicalvoting_b200.sparse.functional import conv_table_forward
    L = _lib.load()
    g = torch.Generator().manual_seed(3)
    try:
        for (n_in, n_out, cin, cout, k3) in [(500, 130, 64, 32, 27), (3000, 3000, 96, 96, 27), (900, 200, 256, 256, 27),
                                             (4000, 1000, 128, 96, 8), (700, 700, 96, 64, 1), (2000, 300, 384, 256, 1),
                                             (30000, 148 * 128 + 5000, 32, 32, 27), (9000, 50000, 64, 96, 8)]:
            table = torch.randint(-1, n_in, (n_out, k3), generator=g, dtype=torch.int64).int()
            table[torch.rand(n_out, k3, generator=g) < 0.5] = -1
            x = torch.randn(n_in, cin, generator=g)
            w = torch.randn(k3, cin, cout, generator=g) * 0.1
            b = torch.randn(1, cout, generator=g)
            want = SO.conv_table(x, w, table, b)
            scale = float(want.abs().max())
            xd, wd, bd, td = x.cuda(), w.cuda(), b.cuda(), table.cuda()
            first = None
            for split, pdl in ((0, 0), (1, 0), (1, 1)):
                L.cvb200_sc_set_conv_options(split, pdl)
                for rep in range(3):     # repeated launches: the split scratch must clean itself
                    got = conv_table_forward(xd, wd, td, bd, mode="tf32")
                    err = float((got.cpu().double() - want).abs().max())
                    assert err <= 3e-3 * scale, "split %d pdl %d rep %d case %s: err %.3e scale %.3e" % (
                        split, pdl, rep, (n_in, n_out, cin, cout, k3), err, scale)
                    first = got if first is None else first
                    assert float((got - first).abs().max()) <= 5e-5 * scale
            # device-side row count: buffers sized for an upper bound, the real n_out read (and planned for) by the kernel
            L.cvb200_sc_set_conv_options(1, 1)
            for real in (n_out, max(n_out // 3, 1), 1):
                ub = n_out + 777
                tab_ub = torch.full((ub, k3), -1, dtype=torch.int32)
                tab_ub[:n_out] = table
                out = torch.full((ub, cout), float("nan"), device="cuda")
                cnt = torch.tensor([real], dtype=torch.int32, device="cuda")
                op = _lib.ScOp()
                op.kind, op.cin, op.cout, op.k3, op.ldi, op.ldo, op.ldr, op.relu = 0, cin, cout, k3, cin, cout, 0, 0
                op.n_out, op.n_in = ub, n_in
                wt = wd.transpose(1, 2).contiguous()
                tab_d = tab_ub.cuda()
                op.in_, op.w, op.bias, op.residual, op.table, op.out = xd.data_ptr(), wt.data_ptr(), bd.data_ptr(), None, tab_d.data_ptr(), out.data_ptr()
                op.n_out_dev = cnt.data_ptr()
                _lib.check(L.cvb200_sc_run_program((_lib.ScOp * 1)(op), 1, _stream()), "run_program")
                torch.cuda.synchronize()
                assert float((out[:real].cpu().double() - want[:real]).abs().max()) <= 3e-3 * scale, (real, n_out, cin, cout, k3)
                assert torch.isnan(out[real:]).all(), "rows beyond the device-side count were written"
    finally:
        L.cvb200_sc_set_conv_options(1, 1)


def test_auto_mode_tensor_cores_without_autograd_fp32_with():
    """Library default ('auto'): the unchanged inference call of the reference scripts -- model(ME.SparseTensor(...)) under
    torch.no_grad(), eval_joint.py:160-171 -- runs its convolutions on tcgen05 and builds its coordinate maps with the fused
    builder (one synchronisation); as soon as autograd records, the exact-fp32 kernels are used."""
    import MinkowskiEngine as ME
    from canonicalvoting_b200.sparse import functional as Fn
    coords, feats = _scene(n=3000, G=24, batch=1, cin=32, seed=5)
    conv = ME.MinkowskiConvolution(32, 64, kernel_size=3, dimension=3).cuda()
    Fn.set_forward_mode("fp32")
    with torch.no_grad():
        exact = conv(ME.SparseTensor(feats, coords, device="cuda")).F
    Fn.set_forward_mode("tf32")
    with torch.no_grad():
        tc = conv(ME.SparseTensor(feats, coords, device="cuda")).F
    Fn.set_forward_mode("auto")
    with torch.no_grad():
        st = ME.SparseTensor(feats, coords, device="cuda")
        auto_nograd = conv(st).F
    assert 16 in st.coordinate_manager.levels, "the fused builder did not run"        # all levels exist after the first request
    auto_grad = conv(ME.SparseTensor(feats.cuda().requires_grad_(True), coords, device="cuda")).F
    scale = float(exact.abs().max())
    assert float((auto_nograd - tc).abs().max()) <= 5e-5 * scale          # same kernel (split tiles: fp32 sums in arrival order)
    assert float((tc - exact).abs().max()) > 5e-5 * scale                 # and a different arithmetic from the exact path
    assert torch.equal(auto_grad.detach(), exact)
    auto_grad.sum().backward()
    assert conv.kernel.grad is not None


def test_persistent_conv_without_split_is_bit_reproducible():
    from canonicalvoting_b200 import _lib
    from canonicalvoting_b200.sparse.functional import conv_table_forward
    L = _lib.load()
    g = torch.Generator().manual_seed(5)
    table = torch.randint(-1, 900, (700, 27), generator=g, dtype=torch.int64).int().cuda()
    x = torch.randn(900, 128, generator=g).cuda()
    w = torch.randn(27, 128, 128, generator=g).cuda() * 0.1
    try:
        L.cvb200_sc_set_conv_options(0, 1)
        a = conv_table_forward(x, w, table, None, mode="tf32")
        b = conv_table_forward(x, w, table, None, mode="tf32")
        assert torch.equal(a, b)
    finally:
        L.cvb200_sc_set_conv_options(1, 1)


@pytest.mark.parametrize("negative", [False, True])
def test_fused_map_builder_equals_stepwise_manager(negative):
    """cvb200_sc_build_maps (one enqueue, one sync, upper-bound allocations, device-side counts) produces exactly the
    tables of the step-by-step coordinate manager: same coarse numbering, same offset order."""
    from canonicalvoting_b200.sparse.coords import CoordinateManager
    coords, _ = _scene(n=6000, G=40, batch=2, seed=21, negative=negative)
    coords = coords.cuda()
    ref = CoordinateManager(coords)
    for ts in (1, 2, 4, 8):
        ref.down(ts)
    got = CoordinateManager.build_unet(coords, 5, 4)
    assert sorted(got.levels) == sorted(ref.levels) == [1, 2, 4, 8, 16]
    for ts in (1, 2, 4, 8, 16):
        assert got.levels[ts].n == ref.levels[ts].n
        assert torch.equal(got.levels[ts].coords, ref.levels[ts].coords)
        assert torch.equal(got.kernel_map(ts, 3), ref.kernel_map(ts, 3))
        assert torch.equal(got._nbr[("ident", ts)].view(-1), torch.arange(ref.levels[ts].n, dtype=torch.int32, device="cuda"))
    assert torch.equal(got.kernel_map(1, 5), ref.kernel_map(1, 5))
    for ts in (1, 2, 4, 8):
        for key in ("children", "up_table", "parent", "koff"):
            assert torch.equal(got.down(ts)[key], ref.down(ts)[key]), (ts, key)
    # a level built lazily on top of the fused tables (the module path's behaviour) still works: 3^3 map via the stored hash map
    assert torch.equal(got.kernel_map(2, 1), ref.kernel_map(2, 1))


@pytest.mark.parametrize("cin,cout,k3", [(32, 32, 27), (96, 96, 27), (128, 96, 27), (64, 128, 8), (256, 256, 27), (384, 256, 1), (192, 128, 27)])
def test_tensor_core_wgrad_matches_fp32_kernel(cin, cout, k3):
    """tcgen05 weight gradient (MN-major tf32 operands, rows as the contraction dimension) vs the fp32 CUDA-core kernel."""
    from canonicalvoting_b200.sparse.functional import conv_wgrad
    g = torch.Generator().manual_seed(cin + cout + k3)
    n_in, n_out = 5000, 4100                      # 4100 rows: a ragged last 32-row stage
    table = torch.randint(-1, n_in, (n_out, k3), generator=g, dtype=torch.int64).int()
    table[torch.rand(n_out, k3, generator=g) < 0.5] = -1
    x = torch.randn(n_in, cin, generator=g).cuda()
    dout = torch.randn(n_out, cout, generator=g).cuda()
    ref = conv_wgrad(x, dout, table.cuda(), mode="fp32")
    got = conv_wgrad(x, dout, table.cuda(), mode="tf32")
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    assert got.shape == ref.shape == (k3, cin, cout)
    assert err <= 3e-3 * scale, "cin=%d cout=%d k3=%d: max err %.3e vs scale %.3e" % (cin, cout, k3, err, scale)
    assert err > 0


def test_small_width_conv_im2col_path_forward_and_wgrad():
    """3-channel 5^3 stem in tf32 mode: im2col + tensor-core product / tensor-core weight gradient == the fp32 kernels."""
    from canonicalvoting_b200 import sparse as ME
    coords, feats = _scene(n=3000, G=24, batch=2, cin=3, seed=5)
    torch.manual_seed(2)
    conv = ME.MinkowskiConvolution(3, 32, kernel_size=5, dimension=3).cuda()
    out = {}
    for mode in ("fp32", "tf32"):
        ME.set_forward_mode(mode)
        try:
            conv.zero_grad()
            y = conv(ME.SparseTensor(feats, coords, device="cuda")).F
            (y * torch.linspace(-1, 1, y.numel(), device="cuda").view_as(y)).sum().backward()
            out[mode] = (y.detach().clone(), conv.kernel.grad.detach().clone())
        finally:
            ME.set_forward_mode("fp32")
    for a, b, what in zip(out["tf32"], out["fp32"], ("forward", "dW")):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 3e-3 * float(b.abs().max()), what


def test_edge_cases_tiny_inputs():
    """One-row convolutions, fewer rows than a 32-row stage in the weight gradient, a 10-voxel scene through the fused
    map builder and the engine (levels shrink to a single voxel)."""
    from canonicalvoting_b200.engine import MinkUNetEngine
    from canonicalvoting_b200.minkunet import MinkUNet34C
    from canonicalvoting_b200.sparse.coords import CoordinateManager
    from canonicalvoting_b200.sparse.functional import conv_table_forward, conv_wgrad
    g = torch.Generator().manual_seed(9)
    x = torch.randn(7, 64, generator=g).cuda()
    w = torch.randn(27, 64, 32, generator=g).cuda() * 0.1
    table = torch.randint(-1, 7, (1, 27), generator=g, dtype=torch.int64).int().cuda()
    ref = conv_table_forward(x, w, table, None, mode="fp32")
    got = conv_table_forward(x, w, table, None, mode="tf32")
    assert float((got - ref).abs().max()) <= 3e-3 * max(float(ref.abs().max()), 1e-6)
    table = torch.randint(-1, 7, (5, 27), generator=g, dtype=torch.int64).int().cuda()
    dout = torch.randn(5, 32, generator=g).cuda()
    rw, gw = conv_wgrad(x, dout, table, mode="fp32"), conv_wgrad(x, dout, table, mode="tf32")
    assert float((gw - rw).abs().max()) <= 3e-3 * max(float(rw.abs().max()), 1e-6)
    empty = conv_wgrad(x, dout[:0], table[:0], mode="tf32")
    assert empty.shape == (27, 64, 32) and float(empty.abs().max()) == 0.0
    coords = torch.tensor([[0, i, (3 * i) % 7, (5 * i) % 11] for i in range(10)], dtype=torch.int32).cuda()
    ref_cm = CoordinateManager(coords)
    for ts in (1, 2, 4, 8):
        ref_cm.down(ts)
    cm = CoordinateManager.build_unet(coords, 5, 4)
    assert [cm.levels[ts].n for ts in (1, 2, 4, 8, 16)] == [ref_cm.levels[ts].n for ts in (1, 2, 4, 8, 16)]
    torch.manual_seed(0)
    model = MinkUNet34C(3, 64).cuda().eval()
    out = MinkUNetEngine(model)(coords, torch.rand(10, 3).cuda())
    assert out.shape == (10, 64) and torch.isfinite(out).all()


@pytest.mark.parametrize("qsize", [0.03, None])
def test_sparse_quantize_on_device_equals_host(qsize):
    """ME.utils.sparse_quantize (utils/dataloader.py:197): the device path (hash map, first point per voxel) returns exactly
    what the numpy path returns: voxel coordinates in first-occurrence order, representative indices, inverse map."""
    import MinkowskiEngine as ME
    g = torch.Generator().manual_seed(4)
    pts = (torch.rand(20000, 3, generator=g) * (3.0 if qsize else 40.0) - (1.0 if qsize else 10.0)).float()   # negatives included
    feats = torch.rand(20000, 3, generator=g)
    hc, hf, hi, hv = ME.utils.sparse_quantize(pts, features=feats, quantization_size=qsize, return_index=True, return_inverse=True)
    dc, df, di, dv = ME.utils.sparse_quantize(pts.cuda(), features=feats.cuda(), quantization_size=qsize, return_index=True, return_inverse=True)
    assert 0 < len(hc) < 20000
    assert torch.equal(dc.cpu(), hc) and torch.equal(di.cpu(), hi) and torch.equal(dv.cpu(), hv)
    assert torch.equal(df.cpu(), hf)
    only_index = ME.utils.sparse_quantize(pts.cuda(), quantization_size=qsize, return_index=True)[1]
    assert torch.equal(only_index.cpu(), hi)
    bc = ME.utils.batched_coordinates([dc, dc[:10]])
    assert bc.is_cuda and bc.shape == (len(dc) + 10, 4) and int(bc[-1, 0]) == 1


def test_per_category_engines_share_one_coordinate_map():
    """eval_separate.py:136-186 runs 9 per-category MinkUNet34C(3, 8) models on the same scene: the engines share one
    prefetch() handle (maps built once) and decode_separate gives the script's xyz / scale / objectness."""
    from canonicalvoting_b200.engine import MinkUNetEngine
    from canonicalvoting_b200.minkunet import MinkUNet34C
    coords, feats = _scene(n=3000, G=40, batch=1, cin=3, seed=31)
    engines = []
    for seed in (0, 1):
        torch.manual_seed(seed)
        engines.append(MinkUNetEngine(MinkUNet34C(3, 8).cuda().eval(), pipeline=True))
    handle = engines[0].prefetch(coords.cuda(), feats.cuda())
    outs = [e(None, None, handle) for e in engines]
    torch.cuda.synchronize()
    alone = engines[1](coords.cuda(), feats.cuda())
    assert outs[0].shape == outs[1].shape == (len(coords), 8)
    # same maps, same program; split tiles are summed with float atomics, so two runs differ by a few 1e-4 after 63 layers
    assert float((outs[1] - alone).abs().max()) <= 2e-3 * float(alone.abs().max())
    assert float((outs[0] - outs[1]).abs().max()) > 0.05 * float(alone.abs().max())      # different weights
    xyz, scale, prob = engines[0].decode_separate(outs[0])
    torch.testing.assert_close(prob, torch.softmax(outs[0][:, 6:8], -1)[:, 1])
    torch.testing.assert_close(scale, torch.exp(outs[0][:, 3:6]))
    assert xyz.shape == (len(coords), 3) and bool(((prob >= 0) & (prob <= 1)).all())


@pytest.mark.parametrize("variant", ["MinkUNet14D", "MinkUNet18B", "MinkUNet34B"])
def test_engine_runs_the_other_family_members(variant):
    """The fused engine is built from the module tree, not from a MinkUNet34C layer list: wide (384-channel, three channel
    splits), shallow and narrow-decoder (32-channel) variants of utils/minkunet.py:208-245 against the fp32 module path."""
    import canonicalvoting_b200.minkunet as M
    from canonicalvoting_b200 import sparse as ME
    from canonicalvoting_b200.engine import MinkUNetEngine
    torch.manual_seed(3)
    coords, feats = _scene(n=2500, G=40, batch=1, cin=3, seed=21)
    model = getattr(M, variant)(3, 64).cuda().eval()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.05); m.running_var.uniform_(0.8, 1.2)
        ME.set_forward_mode("fp32")
        ref = model(ME.SparseTensor(feats, coords, device="cuda")).F
    got = MinkUNetEngine(model)(coords.cuda(), feats.cuda())
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert got.shape == ref.shape and float((got - ref).abs().max()) <= 6e-3 * scale


def test_scene_graph_replays_match_the_step_by_step_engine():
    """SceneGraph: map builder + convolution program (device-side row counts, planned in the kernels) + decode (+ vote) captured
    ONCE into a CUDA graph and replayed for different scenes of the same voxel count, against the engine's ordinary path (host
    knows every level size) on the same scenes."""
    from canonicalvoting_b200 import hv_cuda as H
    from canonicalvoting_b200 import synthetic
    from canonicalvoting_b200.engine import MinkUNetEngine
    from canonicalvoting_b200.minkunet import MinkUNet14A
    torch.manual_seed(3)
    model = MinkUNet14A(3, 64).cuda().eval()
    eng = MinkUNetEngine(model)
    n, G, R = 6000, 48, 6
    scenes = [synthetic.make_scene(n, G, R, seed=s) for s in (1, 2, 3)]
    vote = dict(res=0.03, num_rots=R, corner=(0.0, 0.0, 0.0), dims=(G, G, G))
    lane = eng.graph_lane(n, vote=vote, fuse_decode=False)
    fused = eng.graph_lane(n, vote=vote)                 # head decode in the epilogue of `final` (the default for the joint head)
    assert fused.fuse_decode and fused.launches == lane.launches - 1
    for sc in scenes + scenes[:1]:                      # the lanes are reused; the last replay repeats the first scene
        coords = torch.cat([torch.zeros(n, 1, dtype=torch.int32), torch.from_numpy(sc["coords"])], 1).cuda()
        feats = (torch.from_numpy(sc["feats"]) * 2 - 1).cuda()
        out = lane.run(coords, feats)
        torch.cuda.synchronize()
        want_f = eng(coords, feats)
        xyz, scale, cls, prob, points = eng.decode(want_f, coords, 0.03)
        cm = eng.build_maps(coords)
        assert lane.level_counts() == [cm.levels[ts].n for ts in (1, 2, 4, 8, 16)]
        s_ = float(want_f.abs().max())
        # same kernels and tables; the fp32 sums of split tiles are associated in arrival order, which 42 layers turn into a few
        # 1e-4 of the output scale (measured 3e-4; TF32 itself: 3e-3) -- a wrong row count or plan would show as O(1)
        assert float((out["feats"] - want_f).abs().max()) <= 2e-3 * s_
        assert torch.equal(out["points"], points)
        # the decode picks xyz / scale of the arg-max class: rows at a near-tie of two logits may pick another slot -> count them
        def rows_off(a, b, tol):
            return int(((a - b).abs().reshape(n, -1).max(1).values > tol).sum())
        assert rows_off(out["xyz"], xyz, 1e-3 * s_) <= 0.02 * n and rows_off(out["prob"], prob, 1e-3) <= 0.02 * n
        assert int((out["class_pred"] != cls).sum()) <= 0.02 * n
        go, gr, gs = H.forward_host(out["points"], out["xyz"], out["scale"], out["prob"], 0.03, R, vote["corner"], vote["dims"])
        torch.testing.assert_close(out["grids"][0], go, rtol=1e-4, atol=1e-5 * float(go.max()))
        # fused decode: same arithmetic on the accumulator rows as the separate kernel on the stored rows
        fo = fused.run(coords, feats)
        torch.cuda.synchronize()
        assert fo["feats"] is None and torch.equal(fo["points"], points)
        assert int((fo["class_pred"] != out["class_pred"]).sum()) <= 0.02 * n
        assert rows_off(fo["xyz"], out["xyz"], 1e-3 * s_) <= 0.02 * n
        assert rows_off(fo["scale"], out["scale"], 1e-3 * float(scale.max())) <= 0.02 * n
        assert rows_off(fo["prob"], out["prob"], 1e-3) <= 0.02 * n
        fgo, _, _ = H.forward_host(fo["points"], fo["xyz"], fo["scale"], fo["prob"], 0.03, R, vote["corner"], vote["dims"])
        torch.testing.assert_close(fo["grids"][0], fgo, rtol=1e-4, atol=1e-5 * float(fgo.max()))
    with pytest.raises(RuntimeError, match="built for 6000 voxels"):
        lane.run(coords[:100], feats[:100])


def test_out_of_range_coordinates_are_rejected():
    """16 bits per coordinate field in the hash keys (csrc/sparse_hash.cuh): out-of-range voxels raise instead of aliasing."""
    import MinkowskiEngine as ME
    feats = torch.zeros(3, 3)
    ok = torch.tensor([[0, -30000, 0, 30000], [1, 5, 5, 5], [0, 1, 2, 3]], dtype=torch.int32)
    ME.SparseTensor(feats, ok, device="cuda")
    for bad in ([[0, 40000, 0, 0]], [[0, 0, -32768, 0]], [[70000, 0, 0, 0]], [[-1, 0, 0, 0]]):
        c = torch.cat([ok[:2], torch.tensor(bad, dtype=torch.int32)])
        with pytest.raises(ValueError, match="out of range"):
            ME.SparseTensor(feats, c, device="cuda")


def test_module_path_defers_and_fuses_in_inference():
    """The unchanged reference call sequence (conv -> norm -> relu, residual blocks, ME.cat, 1x1x1 down-samples, transposed
    convolutions) under torch.no_grad(): convolutions are deferred and launched as the engine's fused op (BatchNorm folded,
    bias + residual + ReLU in the epilogue).  Same numbers as MinkUNetEngine on the same model; with autograd on nothing is deferred."""
    from canonicalvoting_b200 import sparse as ME
    from canonicalvoting_b200.engine import MinkUNetEngine
    from canonicalvoting_b200.minkunet import MinkUNet34C
    from canonicalvoting_b200.sparse import functional as Fn
    torch.manual_seed(4)
    coords, feats = _scene(n=5000, G=40, batch=2, cin=3, seed=21)
    model = MinkUNet34C(3, 64).cuda().eval()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.05); m.running_var.uniform_(0.8, 1.2); m.weight.uniform_(0.8, 1.2); m.bias.normal_(0, 0.05)
    eng = MinkUNetEngine(model)
    want = eng(coords.cuda(), feats.cuda())
    Fn.set_forward_mode("auto")
    with torch.no_grad():
        st = model(ME.SparseTensor(feats, coords, device="cuda"))
        assert st._pending is not None                     # `final` itself is still deferred until somebody reads .F
        got = st.F
    scale = float(want.abs().max())
    assert got.shape == want.shape and float((got - want).abs().max()) <= 2e-3 * scale     # same kernels; split tiles sum in arrival order (measured 3e-4)
    Fn.set_forward_mode("fp32")
    with torch.no_grad():
        exact = model(ME.SparseTensor(feats, coords, device="cuda")).F
    assert float((got - exact).abs().max()) <= 6e-3 * scale                                 # TF32 vs exact fp32
    # training mode / autograd: the step-by-step path, BatchNorm with batch statistics
    Fn.set_forward_mode("auto")
    model.train()
    out = model(ME.SparseTensor(feats, coords, device="cuda"))
    assert out._pending is None and out.F.requires_grad
