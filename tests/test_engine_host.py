"""CPU tests of the host-side packing logic the inference paths share (canonicalvoting_b200/engine.py pack_conv / _fold): what the
tensor-core kernel is handed must describe the same convolution as the module's parameters.  The kernel's own contraction is
emulated here in float64 from the PACKED operands and compared with the oracle on the unpacked ones (no GPU involved)."""
import torch

from canonicalvoting_b200 import sparse as ME
from canonicalvoting_b200.engine import pack_conv
from oracle import sparse_oracle as SO


def _scene(n=400, G=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    lin = torch.randperm(G ** 3, generator=g)[:n]
    coords = torch.stack([torch.zeros_like(lin), lin // (G * G), (lin // G) % G, lin % G], 1).int()
    return coords, g


def _kernel_map(coords, K):
    """[n, K^3] neighbour table in the package's offset order (x fastest), -1 = missing."""
    index = {tuple(c): i for i, c in enumerate(coords.tolist())}
    h = K // 2
    nbr = torch.full((len(coords), K ** 3), -1, dtype=torch.long)
    for o, c in enumerate(coords.tolist()):
        for k in range(K ** 3):
            j = index.get((c[0], c[1] + k % K - h, c[2] + (k // K) % K - h, c[3] + k // (K * K) - h))
            if j is not None:
                nbr[o, k] = j
    return nbr


def _bn(c, g):
    bn = ME.MinkowskiBatchNorm(c).eval()
    with torch.no_grad():
        bn.bn.weight.copy_(torch.rand(c, generator=g) + 0.5); bn.bn.bias.copy_(torch.randn(c, generator=g) * 0.1)
        bn.bn.running_mean.copy_(torch.randn(c, generator=g) * 0.1); bn.bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    return bn


def test_packed_tensor_core_operands_describe_the_folded_convolution():
    coords, g = _scene()
    nbr = _kernel_map(coords, 3)
    conv = ME.MinkowskiConvolution(32, 40, kernel_size=3, bias=True, dimension=3)        # 40 channels: padded to 48 for the kernel
    bn = _bn(40, g)
    x = torch.randn(len(coords), 32, generator=g).double()
    w, b, kind, g4 = pack_conv(conv, bn)
    assert kind == 0 and g4 == 0 and tuple(w.shape) == (27, 48, 32) and tuple(b.shape) == (48,)
    # the kernel's contraction from the packed operands: out[o, co] = sum_k x[nbr[o,k]] . w[k, co, :] + b[co]
    out = torch.zeros(len(coords), 48, dtype=torch.float64)
    for k in range(27):
        rows = torch.nonzero(nbr[:, k] >= 0)[:, 0]
        out[rows] += x[nbr[rows, k]] @ w[k].double().t()
    out += b.double()
    want = SO.conv_same(coords, x, conv.kernel.detach().double(), 3, 1, conv.bias.detach().double())
    want = torch.nn.functional.batch_norm(want, bn.bn.running_mean.double(), bn.bn.running_var.double(), bn.bn.weight.detach().double(),
                                          bn.bn.bias.detach().double(), False, 0.0, bn.bn.eps)
    assert float((out[:, :40] - want).abs().max()) <= 1e-5 * float(want.abs().max())
    assert float(out[:, 40:].abs().max()) == 0.0                                          # the padding channels stay zero


def test_packed_stem_gathers_eight_neighbours_of_four_channels_per_k_block():
    """The 3-channel 5^3 stem (utils/minkunet.py:53) as CVB200_OP_CONV_TC_GATHER4: input padded to 4 channels, contraction axis =
    (neighbour, channel) pairs padded to a multiple of 32 -- w[co][4 k + c]."""
    coords, g = _scene(seed=1)
    nbr = _kernel_map(coords, 5)
    conv = ME.MinkowskiConvolution(3, 32, kernel_size=5, dimension=3)
    bn = _bn(32, g)
    x = torch.randn(len(coords), 3, generator=g).double()
    w, b, kind, g4 = pack_conv(conv, bn)
    assert kind == 3 and g4 == 125 and tuple(w.shape) == (1, 32, 512)                     # 125 neighbours -> 16 k-blocks of 8 x 4
    x4 = torch.nn.functional.pad(x, (0, 1))
    out = torch.zeros(len(coords), 32, dtype=torch.float64)
    for k in range(125):
        rows = torch.nonzero(nbr[:, k] >= 0)[:, 0]
        out[rows] += x4[nbr[rows, k]] @ w[0, :, 4 * k:4 * k + 4].double().t()
    assert float(w[0, :, 500:].abs().max()) == 0.0                                        # padding of the contraction axis
    out += b.double()
    want = SO.conv_same(coords, x, conv.kernel.detach().double(), 5, 1, None)
    want = torch.nn.functional.batch_norm(want, bn.bn.running_mean.double(), bn.bn.running_var.double(), bn.bn.weight.detach().double(),
                                          bn.bn.bias.detach().double(), False, 0.0, bn.bn.eps)
    assert float((out - want).abs().max()) <= 1e-5 * float(want.abs().max())


def test_other_input_widths_are_left_to_the_caller():
    conv = ME.MinkowskiConvolution(40, 32, kernel_size=3, dimension=3)
    assert pack_conv(conv, None) is None
