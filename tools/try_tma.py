"""Bring-up check of the TMA-fed convolution against the cp.async one (run on the GPU box, under `timeout`)."""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from canonicalvoting_b200 import _lib
from canonicalvoting_b200.sparse.functional import conv_table_forward
L = _lib.load()
g = torch.Generator().manual_seed(3)
ok = True
for (n_in, n_out, cin, cout, k3) in [(500, 130, 64, 32, 27), (3000, 3000, 96, 96, 27), (900, 200, 256, 256, 27), (4000, 1000, 128, 96, 8), (700, 700, 96, 64, 1)]:
    table = torch.randint(-1, n_in, (n_out, k3), generator=g, dtype=torch.int64).int()
    table[torch.rand(n_out, k3, generator=g) < 0.5] = -1
    x = torch.randn(n_in, cin, generator=g).cuda()
    w = torch.randn(k3, cin, cout, generator=g).cuda() * 0.1
    L.cvb200_sc_set_conv_impl(0)
    ref = conv_table_forward(x, w, table.cuda(), None, mode="tf32")
    torch.cuda.synchronize()
    L.cvb200_sc_set_conv_impl(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
    got = conv_table_forward(x, w, table.cuda(), None, mode="tf32")
    torch.cuda.synchronize()
    err = float((got - ref).abs().max())
    print((n_in, n_out, cin, cout, k3), "max |tma - cp.async| =", err, "scale", float(ref.abs().max()), flush=True)
    ok &= err <= 1e-5 * float(ref.abs().max())
print("TMA OK" if ok else "TMA MISMATCH")
