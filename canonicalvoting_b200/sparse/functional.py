"""Autograd glue of the sparse convolution (host side of csrc/sparse_conv*.cu).

One primitive, driven by a neighbour table:   out[o] = sum_k in[table[o,k]] @ W[k]  (+ bias)
  kind 'same'  stride-1 K^3 convolution (K odd): table = kernel map [n, K^3], in/out share the level
  kind 'down'  stride-2 2^3 convolution:        table = children [n_coarse, 8]
  kind 'up'    transposed stride-2 2^3 conv:    table = parent table [n_fine, 8] (-1 except koff -> parent)
Input gradients are the same primitive on the transposed map:
  'same': the map is its own transpose with mirrored offsets  -> weights W[K^3-1-k]^T
  'down': transpose = parent table,  weights W[k]^T;   'up': transpose = children table, weights W[k]^T
Weight gradients: dW[k] = sum_o in[table[o,k]]^T (x) dout[o]   (cvb200_sc_conv_wgrad, split-K + atomics).
"""
import ctypes

import torch

from .. import _lib
from .coords import _ptr, _stream

# "fp32": CUDA-core implicit GEMM, exact fp32 accumulation (parity mode, always used for gradients)
# "tf32": tcgen05 tensor-core implicit GEMM (kind::tf32, fp32 accumulate in TMEM) for forward / input-gradient
_FORWARD_MODE = "fp32"


def set_forward_mode(mode):
    global _FORWARD_MODE
    if mode not in ("fp32", "tf32"):
        raise ValueError("mode must be 'fp32' or 'tf32'")
    if mode == "tf32" and not hasattr(_lib.load(), "cvb200_sc_conv_forward_tc"):
        raise RuntimeError("libcvb200.so was built without the tensor-core convolution")
    _FORWARD_MODE = mode


def get_forward_mode():
    return _FORWARD_MODE


def conv_table_forward(x, w, table, bias=None, mode=None):
    """x [n_in,cin] f32, w [K3,cin,cout] f32, table [n_out,K3] i32 -> [n_out,cout] f32."""
    L = _lib.load()
    n_out, k3 = table.shape
    cin, cout = w.shape[1], w.shape[2]
    assert x.is_cuda and x.dtype == torch.float32 and x.shape[1] == cin and w.shape[0] == k3
    x, w = x.contiguous(), w.contiguous()
    out = torch.empty((n_out, cout), dtype=torch.float32, device=x.device)
    b = bias.contiguous().view(-1) if bias is not None else None
    mode = mode or _FORWARD_MODE
    with torch.cuda.device(x.device):
        if mode == "tf32" and cin % 32 == 0 and cout % 16 == 0 and 16 <= cout <= 256 and k3 <= 32:
            wt = w.transpose(1, 2).contiguous()     # [k3, cout, cin]: K-major B operand
            rc = L.cvb200_sc_conv_forward_tc(_ptr(x), x.shape[0], cin, _ptr(wt), cout, _ptr(table), n_out, k3,
                                             _ptr(b) if b is not None else None, _ptr(out), _stream())
            _lib.check(rc, "cvb200_sc_conv_forward_tc")
        else:
            rc = L.cvb200_sc_conv_forward(_ptr(x), cin, _ptr(w), cout, _ptr(table), n_out, k3,
                                          _ptr(b) if b is not None else None, _ptr(out), _stream())
            _lib.check(rc, "cvb200_sc_conv_forward")
    return out


def conv_wgrad(a, b, table, table_on_b=False):
    """dW [K3, ca, cb] = sum_r a[ia]^T (x) b[ib] (see include/cvb200.h)."""
    L = _lib.load()
    n_rows, k3 = table.shape
    a, b = a.contiguous(), b.contiguous()
    dw = torch.empty((k3, a.shape[1], b.shape[1]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = L.cvb200_sc_conv_wgrad(_ptr(a), a.shape[1], _ptr(b), b.shape[1], _ptr(table), n_rows, k3,
                                    1 if table_on_b else 0, _ptr(dw), _stream())
        _lib.check(rc, "cvb200_sc_conv_wgrad")
    return dw


class SparseConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, table, table_t, kind):
        ctx.save_for_backward(x, w)
        ctx.table, ctx.table_t, ctx.kind, ctx.has_bias = table, table_t, kind, bias is not None
        return conv_table_forward(x, w, table, bias)

    @staticmethod
    def backward(ctx, gout):
        x, w = ctx.saved_tensors
        gout = gout.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if ctx.kind == "same":
                wt = w.flip(0).transpose(1, 2).contiguous()
            else:
                wt = w.transpose(1, 2).contiguous()
            gx = conv_table_forward(gout, wt, ctx.table_t)
        if ctx.needs_input_grad[1]:
            gw = conv_wgrad(x, gout, ctx.table)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gout.sum(0, keepdim=True)
        return gx, gw, gb, None, None, None


def sparse_conv(x, w, bias, table, table_t, kind):
    return SparseConvFunction.apply(x, w, bias, table, table_t, kind)
