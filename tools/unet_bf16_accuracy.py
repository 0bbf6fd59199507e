#!/usr/bin/env python
"""tools/unet_bf16_accuracy.py -- round-2 tool (not yet run on a GPU): MinkUNet34C forward with every convolution but the stem
on the experimental bf16 kernel (csrc/sparse_conv_bf16.cu), layer by layer from python, against the exact-fp32 module path and
the TF32 engine on the same scene.  Answers the open point of DESIGN.md section 6 item 1: what bf16 activations cost in
accuracy (TF32 engine today: 2e-3 of the output scale).  Run tools/try_bf16_conv.py first.

    python tools/unet_bf16_accuracy.py [C2]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from canonicalvoting_b200 import sparse as ME  # noqa: E402
from canonicalvoting_b200.engine import _DECODER, _ENCODER, MinkUNetEngine, _fold  # noqa: E402
from canonicalvoting_b200.sparse import bf16 as B  # noqa: E402
from canonicalvoting_b200.sparse.coords import CoordinateManager  # noqa: E402
from canonicalvoting_b200.sparse.functional import conv_table_forward  # noqa: E402


def forward_bf16(model, coords, feats):
    """BatchNorm folded, ReLU / residual in the epilogue, activations bf16 between layers; wiring = utils/minkunet.py:122-180."""
    dev = feats.device
    cm = CoordinateManager(coords)

    def ident(n):
        return torch.arange(n, dtype=torch.int32, device=dev).view(-1, 1).contiguous()

    def conv(x, conv_mod, bn_mod, table, relu, residual=None, out_f32=False):
        w, b = _fold(conv_mod, bn_mod)
        return B.conv_table_forward_bf16(x, B.pack_weights(w), table, w.shape[1], b, residual, relu, out_f32)

    def blocks(stage, x, nbr):
        for blk in stage:
            t = conv(x, blk.conv1, blk.norm1, nbr, True)
            r = x if blk.downsample is None else conv(x, blk.downsample[0], blk.downsample[1], ident(x.shape[0]), False)
            x = conv(t, blk.conv2, blk.norm2, nbr, True, residual=r)
        return x

    m = model
    w0, b0 = _fold(m.conv0p1s1, m.bn0)                                  # 3-channel 5^3 stem: fp32 CUDA-core kernel
    x = torch.relu(conv_table_forward(feats, w0, cm.kernel_map(1, m.conv0p1s1.kernel_size), b0.view(1, -1), mode="fp32")).to(torch.bfloat16)
    skips, ts = [x], 1
    for cname, bname, block in _ENCODER:
        d = cm.down(ts)
        x = conv(x, getattr(m, cname), getattr(m, bname), d["children"], True)
        ts *= 2
        x = blocks(getattr(m, block), x, cm.kernel_map(ts, 3))
        skips.append(x)
    skips.pop()
    for cname, bname, block in _DECODER:
        ts //= 2
        x = conv(x, getattr(m, cname), getattr(m, bname), cm._down[ts]["up_table"], True)
        x = torch.cat([x, skips.pop()], 1).contiguous()
        x = blocks(getattr(m, block), x, cm.kernel_map(ts, 3))
    return conv(x, m.final, None, ident(x.shape[0]), False, out_f32=True)


if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    dev = torch.device("cuda", 0)
    sc = bench.scene_for(wl, 0)
    model = bench.make_model().to(dev)
    c_h, f_h = bench.scene_tensors(sc)
    coords, feats = c_h.to(dev), f_h.to(dev)
    with torch.no_grad():
        ME.set_forward_mode("fp32")
        ref = model(ME.SparseTensor(feats, coords, device=dev)).F
        tf32 = MinkUNetEngine(model)(coords, feats)
        bf = forward_bf16(model, coords, feats)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    print("output scale %.3e | tf32 engine: max err %.3e (%.2e of scale) | bf16 layers: max err %.3e (%.2e of scale), rms %.2e of scale" % (
        scale, float((tf32 - ref).abs().max()), float((tf32 - ref).abs().max()) / scale, float((bf - ref).abs().max()),
        float((bf - ref).abs().max()) / scale, float((bf - ref).pow(2).mean().sqrt()) / scale))
    # what it does to the decoded heads (eval_joint.py:173-190): class decisions and objectness
    from canonicalvoting_b200.minkunet import decode_heads
    a, b = decode_heads(ref), decode_heads(bf)
    print("class_pred agreement %.4f, max |prob| diff %.3e, max |xyz| diff %.3e" % (
        float((a[2] == b[2]).float().mean()), float((a[3] - b[3]).abs().max()), float((a[0] - b[0]).abs().max())))
