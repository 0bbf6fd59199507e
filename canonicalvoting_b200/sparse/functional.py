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

# "fp32": CUDA-core implicit GEMM, exact fp32 accumulation (parity mode)
# "tf32": tcgen05 tensor-core implicit GEMM (kind::tf32, fp32 accumulate in TMEM) for forward, input and weight gradient
# "auto" (default): tf32 where autograd is off -- the unchanged inference call of the reference scripts, `with torch.no_grad():
#         model(ME.SparseTensor(...))` (eval_joint.py:160-171), lands on the tensor cores -- and fp32 where a graph is recorded
#         (training keeps exact fp32 gradients unless the caller asks for tf32, as tools/bench_train.py and bench.py do)
_FORWARD_MODE = "auto"


def set_forward_mode(mode):
    global _FORWARD_MODE
    if mode not in ("fp32", "tf32", "auto"):
        raise ValueError("mode must be 'fp32', 'tf32' or 'auto'")
    _FORWARD_MODE = mode


def get_forward_mode():
    return _FORWARD_MODE


def resolve_mode(mode=None):
    """The arithmetic a convolution issued NOW uses: an explicit `mode`, else the global setting, 'auto' decided by autograd."""
    mode = mode or _FORWARD_MODE
    if mode == "auto":
        return "fp32" if torch.is_grad_enabled() else "tf32"
    return mode


def _tc_ok(cin, cout, k3):
    return cin % 32 == 0 and cout % 16 == 0 and 16 <= cout <= 512 and k3 <= 32


def packed_weight(w):
    """[K3, cin, cout] -> [K3, cout, cin] contiguous, the K-major B operand of the tensor-core kernel.  The copy is remembered on
    the tensor object together with its version counter: an inference loop packs every kernel once, a training loop once per
    optimizer step (the in-place parameter update bumps the version) instead of once per call."""
    tag = getattr(w, "_cvb200_packed", None)
    if tag is not None and tag[0] == w._version and tag[1].device == w.device:
        return tag[1]
    wt = w.detach().transpose(1, 2).contiguous()
    try:
        w._cvb200_packed = (w._version, wt)
    except Exception:
        pass
    return wt


def mirrored_table(table):
    """A stride-1 kernel map with the offset axis reversed: the transposed map of a centred odd kernel (offset k <-> K^3-1-k),
    which lets the input gradient use the weights as they are.  Remembered on the table object (one copy per scene and map)."""
    m = getattr(table, "_cvb200_mirror", None)
    if m is None:
        m = table.flip(1).contiguous()
        try:
            table._cvb200_mirror = m
        except Exception:
            pass
    return m


def conv_table_forward(x, w, table, bias=None, mode=None, wt=None):
    """x [n_in,cin] f32, w [K3,cin,cout] f32, table [n_out,K3] i32 -> [n_out,cout] f32.  `wt` (tf32 mode): the weight already as
    [K3, cout, cin] (then `w` may be None)."""
    L = _lib.load()
    n_out, k3 = table.shape
    cin, cout = (wt.shape[2], wt.shape[1]) if wt is not None else (w.shape[1], w.shape[2])
    assert x.is_cuda and x.dtype == torch.float32 and x.shape[1] == cin and (wt if wt is not None else w).shape[0] == k3
    x = x.contiguous()
    out = torch.empty((n_out, cout), dtype=torch.float32, device=x.device)
    b = bias.contiguous().view(-1) if bias is not None else None
    mode = resolve_mode(mode)
    with torch.cuda.device(x.device):
        if mode == "tf32" and _tc_ok(cin, cout, k3):
            if wt is None:
                wt = packed_weight(w)               # [k3, cout, cin]: K-major B operand
            wt = wt.contiguous()
            rc = L.cvb200_sc_conv_forward_tc(_ptr(x), x.shape[0], cin, _ptr(wt), cout, _ptr(table), n_out, k3,
                                             _ptr(b) if b is not None else None, _ptr(out), _stream())
            _lib.check(rc, "cvb200_sc_conv_forward_tc")
        else:
            w = w.contiguous() if w is not None else wt.transpose(1, 2).contiguous()
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
    mode = resolve_mode(mode)
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
    def forward(ctx, x, w, bias, table, table_t, kind, mode):
        ctx.table, ctx.table_t, ctx.kind, ctx.has_bias = table, table_t, kind, bias is not None
        cin, cout = w.shape[1], w.shape[2]
        ctx.im2col = None
        ctx.mode = mode                        # decided by the caller: inside a Function autograd is always off
        if mode == "tf32" and cin % 32 != 0 and cin <= 8 and cout % 32 == 0 and cout <= 256 and not ctx.needs_input_grad[0]:
            # small input width: im2col + one tensor-core product (forward) / one tensor-core weight gradient (backward)
            k3 = w.shape[0]
            width = _im2col_width(k3, cin)
            col = im2col(x, table, width)
            wp = torch.zeros((1, width, cout), dtype=w.dtype, device=w.device)
            wp[0, :k3 * cin] = w.reshape(k3 * cin, cout)
            ident = torch.arange(col.shape[0], dtype=torch.int32, device=x.device).view(-1, 1)
            ctx.im2col = (col, ident, k3, cin)
            ctx.save_for_backward(x, w)
            return conv_table_forward(col, wp, ident, bias, mode=mode)
        ctx.save_for_backward(x, w)
        return conv_table_forward(x, w, table, bias, mode=mode)

    @staticmethod
    def backward(ctx, gout):
        x, w = ctx.saved_tensors
        gout = gout.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # gx[i] = sum_k gout[table_t[i,k]] @ W[k]^T: the B operand [K3, cin, cout] of that product is the kernel as stored; a
            # centred odd kernel's transposed map is the map itself with the offsets mirrored
            cin, cout, k3 = w.shape[1], w.shape[2], w.shape[0]
            if ctx.mode == "tf32" and _tc_ok(cout, cin, k3):
                tab = mirrored_table(ctx.table_t) if ctx.kind == "same" else ctx.table_t
                gx = conv_table_forward(gout, None, tab, mode="tf32", wt=w.detach())
            else:
                wt = (w.flip(0) if ctx.kind == "same" else w).transpose(1, 2).contiguous()
                gx = conv_table_forward(gout, wt, ctx.table_t, mode=ctx.mode)
        if ctx.needs_input_grad[1] and ctx.im2col is not None:
            col, ident, k3, cin = ctx.im2col
            gw = conv_wgrad(col, gout, ident, mode=ctx.mode)[0, :k3 * cin].reshape(k3, cin, gout.shape[1])
        elif ctx.needs_input_grad[1]:
            gw = conv_wgrad(x, gout, ctx.table, mode=ctx.mode)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gout.sum(0, keepdim=True)
        return gx, gw, gb, None, None, None, None


def sparse_conv(x, w, bias, table, table_t, kind):
    return SparseConvFunction.apply(x, w, bias, table, table_t, kind, resolve_mode())
