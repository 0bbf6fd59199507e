"""Autograd glue of the sparse convolution (host side of csrc/sparse_conv*.cu).

One primitive, driven by a neighbour table:   out[o] = sum_k in[table[o,k]] @ W[k]  (+ bias)
  kind 'same'  stride-1 K^3 convolution (K odd): table = kernel map [n, K^3], in/out share the level
  kind 'down'  stride-2 2^3 convolution:        table = children [n_coarse, 8]
  kind 'up'    transposed stride-2 2^3 conv:    table = parent table [n_fine, 8] (-1 except koff -> parent)
Input gradients are the same primitive on the transposed map:
  'same': the map is its own transpose with mirrored offsets  -> weights W[K^3-1-k]^T
  'down': transpose = parent table,  weights W[k]^T;   'up': transpose = children table, weights W[k]^T
Weight gradients: dW[k] = sum_o in[table[o,k]]^T (x) dout[o]   (cvb200_sc_conv_wgrad: fp32 CUDA cores;
cvb200_sc_conv_wgrad_tc: tcgen05, the rows as the contraction dimension).
A convolution with a tiny input width (the 3-channel 5^3 stem, utils/minkunet.py:53) runs in tf32 mode as im2col + a
1x1x1 convolution on the [N, K^3*cin -> padded to 32] matrix, forward and weight gradient alike.
"""
import ctypes

import torch

from .. import _lib
from .coords import _ptr, _stream

# "fp32": CUDA-core implicit GEMM, exact fp32 accumulation (parity mode, always used for gradients)
# "tf32": tcgen05 tensor-core implicit GEMM (kind::tf32, fp32 accumulate in TMEM) for forward, input and weight gradient
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


def conv_wgrad(a, b, table, table_on_b=False, mode=None):
    """dW [K3, ca, cb] = sum_r a[ia]^T (x) b[ib] (see include/cvb200.h).  In tf32 mode the contraction over the rows runs
    on the tensor cores (cvb200_sc_conv_wgrad_tc) when the channel counts allow it."""
    L = _lib.load()
    n_rows, k3 = table.shape
    a, b = a.contiguous(), b.contiguous()
    dw = torch.empty((k3, a.shape[1], b.shape[1]), dtype=torch.float32, device=a.device)
    mode = mode or _FORWARD_MODE
    ca, cb = a.shape[1], b.shape[1]
    with torch.cuda.device(a.device):
        if mode == "tf32" and not table_on_b and ca % 32 == 0 and cb % 32 == 0 and cb <= 256:
            rc = L.cvb200_sc_conv_wgrad_tc(_ptr(a), ca, _ptr(b), cb, _ptr(table), n_rows, k3, _ptr(dw), _stream())
            _lib.check(rc, "cvb200_sc_conv_wgrad_tc")
            return dw
        rc = L.cvb200_sc_conv_wgrad(_ptr(a), a.shape[1], _ptr(b), b.shape[1], _ptr(table), n_rows, k3,
                                    1 if table_on_b else 0, _ptr(dw), _stream())
        _lib.check(rc, "cvb200_sc_conv_wgrad")
    return dw


def im2col(x, table, width):
    """col [n_out, width]: col[o, k*cin + c] = x[table[o,k], c] (0 for a missing neighbour and in the padding columns)."""
    L = _lib.load()
    n_out, k3 = table.shape
    cin = x.shape[1]
    assert width % cin == 0 and width >= k3 * cin and cin <= 8
    x = x.contiguous()
    col = torch.empty((n_out, width), dtype=torch.float32, device=x.device)
    o = _lib.ScOp()
    o.kind, o.cin, o.cout, o.k3, o.ldi, o.ldo, o.n_out, o.n_in = 2, cin, width, k3, cin, width, n_out, x.shape[0]
    o.in_, o.out, o.table = x.data_ptr(), col.data_ptr(), table.data_ptr()
    with torch.cuda.device(x.device):
        _lib.check(L.cvb200_sc_run_program((_lib.ScOp * 1)(o), 1, _stream()), "cvb200_sc_run_program(im2col)")
    return col


def _im2col_width(k3, cin):
    w = (k3 * cin + 31) // 32 * 32
    while w % cin:
        w += 32
    return w


class SparseConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, table, table_t, kind):
        ctx.table, ctx.table_t, ctx.kind, ctx.has_bias = table, table_t, kind, bias is not None
        cin, cout = w.shape[1], w.shape[2]
        ctx.im2col = None
        if _FORWARD_MODE == "tf32" and cin % 32 != 0 and cin <= 8 and cout % 32 == 0 and cout <= 256 and not ctx.needs_input_grad[0]:
            # small input width: im2col + one tensor-core product (forward) / one tensor-core weight gradient (backward)
            k3 = w.shape[0]
            width = _im2col_width(k3, cin)
            col = im2col(x, table, width)
            wp = torch.zeros((1, width, cout), dtype=w.dtype, device=w.device)
            wp[0, :k3 * cin] = w.reshape(k3 * cin, cout)
            ident = torch.arange(col.shape[0], dtype=torch.int32, device=x.device).view(-1, 1)
            ctx.im2col = (col, ident, k3, cin)
            ctx.save_for_backward(x, w)
            return conv_table_forward(col, wp, ident, bias)
        ctx.save_for_backward(x, w)
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
        if ctx.needs_input_grad[1] and ctx.im2col is not None:
            col, ident, k3, cin = ctx.im2col
            gw = conv_wgrad(col, gout, ident)[0, :k3 * cin].reshape(k3, cin, gout.shape[1])
        elif ctx.needs_input_grad[1]:
            gw = conv_wgrad(x, gout, ctx.table)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gout.sum(0, keepdim=True)
        return gx, gw, gb, None, None, None


def sparse_conv(x, w, bias, table, table_t, kind):
    return SparseConvFunction.apply(x, w, bias, table, table_t, kind)
