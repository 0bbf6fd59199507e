"""EXPERIMENTAL host side of the bf16 tensor-core convolution (csrc/sparse_conv_bf16.cu) -- a round-2 work item that has
not run on a GPU yet and is not used by the engine or the modules.  `tools/try_bf16_conv.py` is its first check.

    wp = pack_weights(kernel)                                # [K3, cin, cout] float -> [cout, K3*cin] bf16, contraction axis flattened
    y = conv_table_forward_bf16(x_bf16, wp, table, cin, bias=None, residual=None, relu=False, out_f32=False)
"""
import torch

from .. import _lib
from .coords import _ptr, _stream

KB = 64          # bf16 elements per k-block (one 128-byte row of the shared-memory tiles)


def pack_weights(kernel):
    """w[co, k * cin + c] = kernel[k, c, co] as bf16 (what the weight TMA of the kernel reads, box = 64 columns x nc rows)."""
    k3, cin, cout = kernel.shape
    return kernel.permute(2, 0, 1).reshape(cout, k3 * cin).to(torch.bfloat16).contiguous()


def chunk_source(j, c, cin):
    """(kernel offset, first channel) of the 16-byte chunk c of k-block j on the flattened (offset, channel) axis."""
    flat = KB * j + 8 * c
    return flat // cin, flat % cin


def conv_table_forward_bf16(x, wp, table, cin, bias=None, residual=None, relu=False, out_f32=False):
    """x bf16 [n_in, cin], wp = pack_weights(kernel), table int32 [n_out, K3] -> [n_out, cout] bf16 (float32 if out_f32)."""
    L = _lib.load()
    if not (x.is_cuda and x.dtype == torch.bfloat16 and wp.dtype == torch.bfloat16):
        raise RuntimeError("conv_table_forward_bf16: bf16 CUDA tensors expected (there is no CPU path)")
    n_out, k3 = table.shape
    cout = wp.shape[0]
    assert x.shape[1] == cin and wp.shape[1] == k3 * cin
    x = x.contiguous()
    out = torch.empty((n_out, cout), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    b = bias.float().contiguous().view(-1) if bias is not None else None
    r = residual.contiguous() if residual is not None else None
    with torch.cuda.device(x.device):
        rc = L.cvb200_sc_conv_forward_bf16(_ptr(x), x.shape[0], cin, cin, _ptr(wp), cout, _ptr(table), n_out, k3,
                                           _ptr(b) if b is not None else None, _ptr(r) if r is not None else None,
                                           cout if r is not None else 0, 1 if relu else 0, _ptr(out), cout, 1 if out_f32 else 0, _stream())
        _lib.check(rc, "cvb200_sc_conv_forward_bf16")
    return out
