"""tools/initcheck_probe.py -- run under `compute-sanitizer --tool initcheck`: a biased sparse convolution forward +
backward with the address ranges of every tensor printed, so that a reported address can be attributed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import MinkowskiEngine as ME  # noqa: E402
from tests.test_sparse_gpu import _scene  # noqa: E402


def rng(name, t):
    print("%-8s %x .. %x  shape %s" % (name, t.data_ptr(), t.data_ptr() + t.numel() * t.element_size(), tuple(t.shape)), flush=True)


for K, cin, cout in [(3, 32, 64), (3, 96, 96)]:
    print("case", K, cin, cout, flush=True)
    coords, feats = _scene(n=2000, G=20, cin=cin, seed=K + cin)
    conv = ME.MinkowskiConvolution(cin, cout, kernel_size=K, bias=True, dimension=3).cuda()
    x = feats.cuda().requires_grad_(True)
    y = conv(ME.SparseTensor(x, coords, device="cuda")).F
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
    g = gy.cuda()
    torch.cuda.synchronize()
    for nm, t in (("x", x), ("y", y), ("g", g), ("kernel", conv.kernel), ("bias", conv.bias)):
        rng(nm, t)
    print("direct sum", flush=True)
    s = g.sum(0, keepdim=True)
    torch.cuda.synchronize()
    print("backward", flush=True)
    y.backward(g)
    torch.cuda.synchronize()
    rng("x.grad", x.grad)
    rng("k.grad", conv.kernel.grad)
    rng("b.grad", conv.bias.grad)
