"""CPU tests that pin the sparse-convolution oracle (oracle/sparse_oracle.py) against DENSE convolutions
(torch.nn.functional.conv3d / conv_transpose3d on the densified scene) for every (kernel, stride) variant of
MinkUNet34C, plus the host-side ME.utils helpers and the model's structure.  ME itself is unavailable
offline (parity unpinned); dense equivalence is what pins the semantics."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sparse_oracle as SO


def _scene(n=60, G=8, batch=2, cin=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(batch):
        lin = torch.randperm(G ** 3, generator=g)[:n]
        rows.append(torch.stack([torch.full_like(lin, b), lin // (G * G), (lin // G) % G, lin % G], 1))
    coords = torch.cat(rows).int()
    feats = torch.randn(coords.shape[0], cin, generator=g, dtype=torch.float64)
    return coords, feats, G, batch


def _dense(coords, feats, G, batch):
    d = torch.zeros(batch, feats.shape[1], G, G, G, dtype=feats.dtype)
    c = coords.long()
    d[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = feats
    return d


def _w_dense(kernel, K):
    # kernel [K^3, Cin, Cout], k = ix + K*(iy + K*iz)  ->  conv3d weight [Cout, Cin, kx, ky, kz] (spatial dims x,y,z)
    cin, cout = kernel.shape[1], kernel.shape[2]
    return kernel.reshape(K, K, K, cin, cout).permute(4, 3, 2, 1, 0).contiguous()   # [iz,iy,ix,..] -> [.., ix,iy,iz]


@pytest.mark.parametrize("K", [3, 5])
def test_same_conv_equals_dense_conv3d(K):
    coords, feats, G, batch = _scene()
    g = torch.Generator().manual_seed(1)
    kernel = torch.randn(K ** 3, 3, 4, generator=g, dtype=torch.float64)
    out = SO.conv_same(coords, feats, kernel, K)
    dense = F.conv3d(_dense(coords, feats, G, batch), _w_dense(kernel, K), padding=K // 2)
    c = coords.long()
    torch.testing.assert_close(out, dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]])


def test_same_conv_at_tensor_stride_2_is_a_dilated_dense_conv():
    coords, feats, G, batch = _scene(n=40, G=8)
    coords = coords.clone(); coords[:, 1:] *= 2            # a stride-2 coordinate set
    kernel = torch.randn(27, 3, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    out = SO.conv_same(coords, feats, kernel, 3, tensor_stride=2)
    dense = F.conv3d(_dense(coords, feats, 2 * G, batch), _w_dense(kernel, 3), padding=2, dilation=2)
    c = coords.long()
    torch.testing.assert_close(out, dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]])


def test_down_conv_equals_dense_stride2_conv():
    coords, feats, G, batch = _scene(n=80, G=8)
    kernel = torch.randn(8, 3, 6, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    coarse, out = SO.conv_down(coords, feats, kernel, tensor_stride=1)
    dense = F.conv3d(_dense(coords, feats, G, batch), _w_dense(kernel, 2), stride=2)
    c = coarse.long()
    assert (c[:, 1:] % 2 == 0).all()
    torch.testing.assert_close(out, dense[c[:, 0], :, c[:, 1] // 2, c[:, 2] // 2, c[:, 3] // 2])
    # every occupied coarse cell is present exactly once
    occ = (F.max_pool3d(_dense(coords, torch.ones(len(coords), 1, dtype=torch.float64), G, batch), 2) > 0).sum()
    assert len(coarse) == int(occ) == len({tuple(r) for r in coarse.tolist()})


def test_up_conv_equals_dense_conv_transpose():
    coords, feats, G, batch = _scene(n=80, G=8)
    coarse, parent, koff = SO.coarse_coords(coords, 2)
    cf = torch.randn(len(coarse), 4, dtype=torch.float64, generator=torch.Generator().manual_seed(4))
    kernel = torch.randn(8, 4, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    out = SO.conv_up(coords, cf, kernel, tensor_stride=2)
    cc = coarse.clone(); cc[:, 1:] //= 2
    dense_in = _dense(cc, cf, G // 2, batch)
    # conv_transpose3d weight [Cin, Cout, kx, ky, kz]
    wt = kernel.reshape(2, 2, 2, 4, 3).permute(3, 4, 2, 1, 0).contiguous()
    dense = F.conv_transpose3d(dense_in, wt, stride=2)
    c = coords.long()
    torch.testing.assert_close(out, dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]])


def test_negative_coordinates_floor():
    coords = torch.tensor([[0, -1, -2, -3], [0, -4, 0, 1], [0, 3, 2, 1]], dtype=torch.int32)
    coarse, parent, koff = SO.coarse_coords(coords, 2)
    assert coarse.tolist() == [[0, -2, -2, -4], [0, -4, 0, 0], [0, 2, 2, 0]]
    assert koff.tolist() == [1 + 2 * (0 + 2 * 1), 0 + 2 * (0 + 2 * 1), 1 + 2 * (0 + 2 * 1)]


def test_me_utils_helpers():
    import MinkowskiEngine as ME
    bc = ME.utils.batched_coordinates([np.array([[0.2, 1.7, -0.5]]), torch.tensor([[3, 4, 5], [6, 7, 8]])])
    assert bc.dtype == torch.int32 and bc.tolist() == [[0, 0, 1, -1], [1, 3, 4, 5], [1, 6, 7, 8]]
    pts = np.array([[0.01, 0.02, 0.0], [0.05, 0.0, 0.0], [0.02, 0.01, 0.02], [0.31, 0.0, 0.0]])
    idx = ME.utils.sparse_quantize(pts, quantization_size=0.03, return_index=True)[1]
    assert sorted(idx.tolist()) == [0, 1, 3]       # rows 0 and 2 share voxel (0,0,0)
    k = torch.empty(27, 16, 32)
    ME.utils.kaiming_normal_(k, mode="fan_out", nonlinearity="relu")
    assert abs(float(k.std()) - (2.0 / (32 * 27)) ** 0.5) < 0.01


def test_minkunet34c_structure_matches_the_reference_model():
    from canonicalvoting_b200.minkunet import MinkUNet34C
    m = MinkUNet34C(3, 64)
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == 37_860_320 and len(list(m.parameters())) == 188
    assert tuple(sd["conv0p1s1.kernel"].shape) == (125, 3, 32)
    assert tuple(sd["block5.0.conv1.kernel"].shape) == (27, 384, 256)
    assert tuple(sd["final.kernel"].shape) == (96, 64) and tuple(sd["final.bias"].shape) == (1, 64)
    assert "block1.0.norm1.bn.running_mean" in sd and "block2.0.downsample.0.kernel" in sd
    n_conv = sum(1 for k in sd if k.endswith("kernel"))
    n_bn = sum(1 for k in sd if k.endswith("bn.weight"))
    assert (n_conv, n_bn) == (63, 62)


def test_model_family_matches_reference_state_dicts():
    """canonicalvoting_b200/minkunet.py (a table) builds what the reference's utils/minkunet.py builds: state-dict keys,
    order and shapes of every variant, recorded from the unmodified reference file by tools/make_model_golden.py."""
    import json
    import os
    import canonicalvoting_b200.minkunet as M
    with open(os.path.join(os.path.dirname(__file__), "golden", "minkunet_state_dicts.json")) as f:
        golden = json.load(f)
    assert len(golden) == 11
    for name, want in golden.items():
        cls, head = name.split("/")
        m = getattr(M, cls)(3, 64 if head == "joint" else 8)
        got = [[k, list(t.shape)] for k, t in m.state_dict().items()]
        assert got == want, name


def test_bottleneck_block_structure():
    """MinkowskiEngine.modules.resnet_block.Bottleneck (utils/minkunet.py:30): 1^3 -> 3^3 -> 1^3 with 4x expansion."""
    from MinkowskiEngine.modules.resnet_block import BasicBlock, Bottleneck
    b = Bottleneck(64, 32, dimension=3)
    assert Bottleneck.expansion == 4 and BasicBlock.expansion == 1
    sd = b.state_dict()
    assert tuple(sd["conv1.kernel"].shape) == (64, 32) and tuple(sd["conv2.kernel"].shape) == (27, 32, 32)
    assert tuple(sd["conv3.kernel"].shape) == (32, 128) and "norm3.bn.running_var" in sd


def test_decompose_splits_rows_by_batch_in_order():
    """SparseTensor.decomposed_coordinates_and_features (sunrgbd/brnetcanon.py:227): per-scene lists, row order kept."""
    from canonicalvoting_b200.sparse.modules import decompose
    g = torch.Generator().manual_seed(0)
    coords = torch.randint(0, 50, (40, 4), generator=g).int()
    coords[:, 0] = torch.randint(0, 3, (40,), generator=g).int()
    feats = torch.randn(40, 5, generator=g)
    cs, fs = decompose(coords, feats)
    assert len(cs) == len(fs) == 3
    for b in range(3):
        m = coords[:, 0] == b
        assert torch.equal(cs[b], coords[m][:, 1:]) and torch.equal(fs[b], feats[m])
    assert decompose(coords[:0], feats[:0]) == ([], [])


def test_fast_cpu_port_matches_dict_oracle():
    """The vectorised CPU port timed by bench.py (cpu_baseline of the U-Net half) == the dict-based oracle."""
    from canonicalvoting_b200.minkunet import MinkUNet14A
    torch.manual_seed(0)
    coords, feats, G, batch = _scene(n=400, G=16, batch=2, seed=5)
    coords = coords.clone(); coords[:, 1:] -= 5
    feats = feats.float()
    model = MinkUNet14A(3, 12).eval()
    a = SO.OracleNet(model).forward(coords, feats)
    b = SO.FastCpuNet(model).forward(coords, feats)
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)


def test_grad_net_is_the_fast_net_with_autograd_leaves():
    """GradCpuNet (the reference of the tf32 training-step test) = FastCpuNet in float64 with the parameters as autograd leaves:
    same forward, and its gradients agree with central finite differences of the loss."""
    from canonicalvoting_b200.minkunet import MinkUNet14A
    torch.manual_seed(2)
    g = torch.Generator().manual_seed(4)
    lin = torch.randperm(48 ** 3, generator=g)[:1500]
    coords = torch.stack([torch.zeros_like(lin), lin // (48 * 48), (lin // 48) % 48, lin % 48], 1).int()
    feats = torch.randn(1500, 3, generator=g)
    model = MinkUNet14A(3, 8).train()
    net = SO.GradCpuNet(model)
    out = net.forward(coords, feats.double())
    with torch.no_grad():
        f32 = SO.FastCpuNet(model).forward(coords, feats)
    assert float((out.detach() - f32.double()).abs().max()) <= 1e-4 * float(f32.abs().max())
    target = torch.randn(out.shape, generator=g).double()
    loss = ((out - target) ** 2).mean()
    loss.backward()
    grads = net.grads()
    assert set(grads) == {n for n, _ in model.named_parameters()} and all(v is not None for v in grads.values())
    for name, idx in (("final.kernel", (3, 5)), ("block4.0.conv1.kernel", (13, 7, 9)), ("bn0.bn.weight", (4,))):
        p = dict(model.named_parameters())[name]
        eps, vals = 2e-4, []
        for sgn in (1, -1):
            with torch.no_grad():
                p[idx] += sgn * eps
            o = SO.GradCpuNet(model).forward(coords, feats.double())
            vals.append(float(((o - target) ** 2).mean()))
            with torch.no_grad():
                p[idx] -= sgn * eps
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - float(grads[name][idx])) <= 2e-2 * max(abs(fd), 1e-3), (name, fd, float(grads[name][idx]))


def test_kernel_offset_permutation_hook():
    """sparse.utils.permute_kernel_offsets: a checkpoint whose kernels number the offsets with z fastest computes, after the hook,
    the same convolution with this package's x-fastest maps (checked by swapping the x and z axes of the scene under the oracle);
    the hook is an involution and leaves 1x1x1 kernels and every other entry alone."""
    from canonicalvoting_b200.sparse.utils import permute_kernel_offsets
    g = torch.Generator().manual_seed(9)
    lin = torch.randperm(12 ** 3, generator=g)[:300]
    coords = torch.stack([torch.zeros_like(lin), lin // 144, (lin // 12) % 12, lin % 12], 1).int()
    feats = torch.randn(300, 4, generator=g).double()
    for K in (3, 5):
        w_ckpt = torch.randn(K ** 3, 4, 6, generator=g).double()
        state = {"conv.kernel": w_ckpt, "down.kernel": torch.randn(4, 6).double(), "bn.bn.weight": torch.ones(6)}
        fixed = permute_kernel_offsets(state, "zyx")
        swapped = coords[:, [0, 3, 2, 1]].contiguous()                 # z-fastest numbering on (x, y, z) == x-fastest numbering on (z, y, x)
        want = SO.conv_same(swapped, feats, w_ckpt, K)
        got = SO.conv_same(coords, feats, fixed["conv.kernel"], K)
        assert float((got - want).abs().max()) <= 1e-12
        assert fixed["down.kernel"] is state["down.kernel"] and fixed["bn.bn.weight"] is state["bn.bn.weight"]
        assert torch.equal(permute_kernel_offsets(fixed, "zyx")["conv.kernel"], w_ckpt)
        mir = permute_kernel_offsets(state, "mirror")["conv.kernel"]
        assert torch.equal(mir, w_ckpt.flip(0)) and torch.equal(permute_kernel_offsets({"conv.kernel": mir}, "mirror")["conv.kernel"], w_ckpt)
        ident = permute_kernel_offsets(state, lambda k: torch.arange(k ** 3))["conv.kernel"]
        assert torch.equal(ident, w_ckpt)
