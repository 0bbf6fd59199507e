#!/usr/bin/env python
"""tools/try_bf16_conv.py -- FIRST GPU CHECK of the experimental bf16 tensor-core convolution (csrc/sparse_conv_bf16.cu,
written at the end of round 1 without a GPU).  For a list of layer shapes: output vs an fp64 reference on the same
bf16-rounded operands (bias / residual / ReLU / fp32 output / split tiles / tail tiles), then the time of the 96 -> 96 and
128 -> 96 3^3 layers on a C2-sized level against the TF32 kernel.  Run under `timeout`: a protocol bug traps after ~2 s
(bounded mbarrier waits), it does not hang.

    python tools/try_bf16_conv.py            # parity, then timing
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from canonicalvoting_b200.sparse import bf16 as B  # noqa: E402
from canonicalvoting_b200.sparse.functional import conv_table_forward  # noqa: E402


def random_table(n_out, n_in, k3, density, g):
    t = torch.randint(0, n_in, (n_out, k3), generator=g)
    t[torch.rand(n_out, k3, generator=g) > density] = -1
    return t.int()


def reference(x, kernel, table, bias, residual, relu):
    out = torch.zeros(table.shape[0], kernel.shape[2], dtype=torch.float64, device=x.device)
    xd, kd = x.double(), kernel.double()
    for k in range(table.shape[1]):
        ok = table[:, k] >= 0
        out[ok] += xd[table[ok, k].long()] @ kd[k]
    if bias is not None:
        out += bias.double()
    if residual is not None:
        out += residual.double()
    return out.clamp_min(0) if relu else out


def check(n_out, n_in, cin, cout, k3, bias, residual, relu, out_f32, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_in, cin, generator=g).to(torch.bfloat16).cuda()
    kernel = (torch.randn(k3, cin, cout, generator=g) * (2.0 / (cin * min(k3, 9))) ** 0.5).to(torch.bfloat16).cuda()
    table = random_table(n_out, n_in, k3, 0.35, g).cuda()
    b = torch.randn(cout, generator=g).cuda() if bias else None
    r = torch.randn(n_out, cout, generator=g).to(torch.bfloat16).cuda() if residual else None
    got = B.conv_table_forward_bf16(x, B.pack_weights(kernel), table, cin, b, r, relu, out_f32).double()
    torch.cuda.synchronize()
    want = reference(x, kernel, table, b, r, relu)
    scale = float(want.abs().max())
    tol = (1e-5 if out_f32 else 1.0 / 128) * scale           # fp32 accumulation; a bf16 output is rounded to 8 bits of mantissa
    err = float((got - want).abs().max())
    print("%-60s max err %.3e (scale %.3e) %s" % ("n_out=%d n_in=%d %d->%d k3=%d bias=%d res=%d relu=%d f32=%d" % (
        n_out, n_in, cin, cout, k3, bias, residual, relu, out_f32), err, scale, "ok" if err <= tol else "FAIL"), flush=True)
    return err <= tol


def timing(n, cin, cout, k3=27, iters=20):
    g = torch.Generator().manual_seed(1)
    table = random_table(n, n, k3, 0.30, g).cuda()
    x32 = torch.randn(n, cin, generator=g).cuda()
    kernel = (torch.randn(k3, cin, cout, generator=g) * 0.05).cuda()
    x16, wp = x32.to(torch.bfloat16), B.pack_weights(kernel)

    def ev(fn):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e3
    t32 = ev(lambda: conv_table_forward(x32, kernel, table, None, mode="tf32"))
    t16 = ev(lambda: B.conv_table_forward_bf16(x16, wp, table, cin))
    print("n=%d %d->%d k3=%d: tf32 %.1f us, bf16 %.1f us (includes the python wrapper and the output allocation of both)" % (n, cin, cout, k3, t32, t16))


if __name__ == "__main__":
    ok = True
    cases = [  # n_out, n_in, cin, cout, k3, bias, residual, relu, out_f32
        (128, 128, 64, 32, 1, 0, 0, 0, 1), (128, 128, 64, 32, 1, 0, 0, 0, 0), (300, 300, 64, 64, 27, 0, 0, 0, 1),
        (300, 300, 32, 32, 27, 1, 1, 1, 0), (1000, 1000, 96, 96, 27, 1, 1, 1, 0), (1000, 1000, 128, 96, 27, 0, 1, 1, 0),
        (2000, 700, 32, 64, 8, 0, 0, 1, 0), (700, 2000, 256, 128, 8, 0, 0, 1, 0), (196, 196, 256, 256, 27, 0, 1, 1, 0),
        (940, 940, 384, 256, 27, 0, 0, 1, 0), (5000, 5000, 96, 64, 1, 1, 0, 0, 1), (5000, 5000, 96, 16, 1, 1, 0, 0, 1),
        (20000, 20000, 96, 96, 27, 0, 1, 1, 0), (50000, 50000, 128, 96, 27, 0, 1, 1, 0), (1, 5, 32, 16, 27, 0, 0, 0, 1),
    ]
    for i, c in enumerate(cases):
        ok &= check(*c, seed=i)
    print("PARITY", "OK" if ok else "FAILED", flush=True)
    if ok:
        timing(50000, 96, 96)
        timing(50000, 128, 96)
        timing(17000, 128, 96)
