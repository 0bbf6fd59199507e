"""SparseTensor and the layer classes of the MinkowskiEngine subset the reference uses:
MinkowskiConvolution / MinkowskiConvolutionTranspose / MinkowskiBatchNorm / MinkowskiReLU / cat / BasicBlock.
Class names, constructor arguments and attribute names (`kernel`, `bias`, `bn`, `conv1`, `norm1`, ...)
follow ME so that utils/minkunet.py and utils/resnet.py of the reference run unchanged and state-dict keys
match (`conv0p1s1.kernel [125,3,32]`, `block1.0.norm1.bn.weight`, `final.bias [1,64]`).  [ME-recall] marks
behaviour restated from recollection of ME 0.5.x (the package is not available offline)."""
import math

import torch
import torch.nn as nn

from .. import _lib
from .coords import CoordinateManager, _stream
from .functional import resolve_mode, sparse_conv


class _Pending:
    """A convolution that has not been launched yet (inference only: autograd off, tensor-core mode).  The modules that follow
    it in the reference's call sequence -- MinkowskiBatchNorm in eval mode, MinkowskiReLU, the residual `out + residual` of
    BasicBlock (utils/resnet.py / ME's resnet_block.py) -- are folded into it instead of running as separate torch kernels;
    the fused kernel (BatchNorm folded into the weights, bias + residual + ReLU in the epilogue: the engine's op) is
    launched when somebody needs the features (`.F`, the next convolution, `ME.cat`)."""
    __slots__ = ("conv", "x", "table", "bn", "residual", "relu")

    def __init__(self, conv, x, table):
        self.conv, self.x, self.table, self.bn, self.residual, self.relu = conv, x, table, None, None, False

    def run(self):
        conv = self.conv
        w, b, kind, g4 = conv._packed_for_inference(self.bn)
        x = self.x
        n_out = self.table.shape[0]
        cout_p = w.shape[1]                                     # padded to a multiple of 16
        if kind == 3 and x.shape[1] != 4:                      # the 3-channel stem: one 16-byte vector per voxel
            x = torch.nn.functional.pad(x, (0, 4 - x.shape[1]))
        if not x.is_contiguous():
            x = x.contiguous()
        out = torch.empty((n_out, cout_p), dtype=torch.float32, device=x.device)
        res = self.residual
        if res is not None and not res.is_contiguous():
            res = res.contiguous()
        # the op structure is kept per (module, folded BatchNorm): only what changes from call to call is rewritten
        tag = conv.__dict__.get("_cvb200_op")
        if tag is None or tag[0] is not w:
            o = _lib.ScOp()
            o.kind, o.cin, o.cout = kind, w.shape[2], cout_p
            o.w, o.bias, o.ldo = w.data_ptr(), (b.data_ptr() if b is not None else None), cout_p
            tag = (w, o, (_lib.ScOp * 1)(), _lib.load().cvb200_sc_run_program)
            conv.__dict__["_cvb200_op"] = tag
        _, o, arr, run_program = tag
        o.k3 = g4 if kind == 3 else self.table.shape[1]
        o.ldi, o.ldr, o.relu = x.shape[1], (res.shape[1] if res is not None else 0), 1 if self.relu else 0
        o.n_out, o.n_in = n_out, x.shape[0]
        o.in_, o.table, o.out = x.data_ptr(), self.table.data_ptr(), out.data_ptr()
        o.residual = res.data_ptr() if res is not None else None
        arr[0] = o
        dev = x.device.index
        if torch.cuda.current_device() != dev:
            with torch.cuda.device(dev):
                rc = run_program(arr, 1, _stream())
        else:
            rc = run_program(arr, 1, _stream())
        if rc:
            _lib.check(rc, "cvb200_sc_run_program")
        return out[:, :conv.out_channels] if cout_p != conv.out_channels else out


class SparseTensor:
    """Features [N,C] + integer coordinates [N,4] (batch,x,y,z) sharing a CoordinateManager.
    `ME.SparseTensor(feats, coords, device='cuda')` (train_joint.py:250): CPU inputs are moved to the device;
    row order is preserved (callers index `.F` with labels in input order, train_joint.py:256)."""

    def __init__(self, features, coordinates=None, device=None, coordinate_manager=None, tensor_stride=1, **_ignored):
        if coordinate_manager is None:
            if coordinates is None:
                raise ValueError("coordinates or coordinate_manager required")
            dev = torch.device(device) if device is not None else (features.device if features.is_cuda else torch.device("cuda"))
            if dev.type != "cuda":
                raise RuntimeError("canonicalvoting_b200.sparse has no CPU path: pass device='cuda'")
            coordinates = coordinates.to(device=dev, dtype=torch.int32)
            features = features.to(dev)
            coordinate_manager = CoordinateManager(coordinates)
        self._F = features
        self._pending = None
        self.coordinate_manager = coordinate_manager
        self.tensor_stride = tensor_stride if isinstance(tensor_stride, int) else int(tensor_stride[0])

    @property
    def F(self):
        if self._pending is not None:          # a deferred convolution (+ folded BatchNorm / residual / ReLU): launch it now
            self._F = self._pending.run()
            self._pending = None
        return self._F

    def _deferred(self, pending, tensor_stride=None):
        t = SparseTensor(None, coordinate_manager=self.coordinate_manager,
                         tensor_stride=self.tensor_stride if tensor_stride is None else tensor_stride)
        t._pending = pending
        return t

    @property
    def C(self):
        return self.coordinate_manager.levels[self.tensor_stride].coords

    @property
    def device(self):
        return self.coordinate_manager.device

    @property
    def shape(self):
        return self.F.shape

    @property
    def decomposed_coordinates_and_features(self):
        """(list of [Ni,3] coordinates, list of [Ni,C] features), one entry per batch index (sunrgbd/brnetcanon.py:227,318);
        rows keep their order [ME-recall]."""
        return decompose(self.C, self.F)

    @property
    def decomposed_coordinates(self):
        return decompose(self.C, self.F)[0]

    @property
    def decomposed_features(self):
        return decompose(self.C, self.F)[1]

    def _like(self, feats, tensor_stride=None):
        return SparseTensor(feats, coordinate_manager=self.coordinate_manager,
                            tensor_stride=self.tensor_stride if tensor_stride is None else tensor_stride)

    def __add__(self, other):
        p = self._pending
        if p is not None and not p.relu and p.residual is None and isinstance(other, SparseTensor) and p.conv.out_channels % 16 == 0:
            # `out += residual` of a residual block right after conv -> norm: the sum moves into the convolution's epilogue
            q = _Pending(p.conv, p.x, p.table)
            q.bn, q.residual = p.bn, other.F
            return self._deferred(q)
        return self._like(self.F + (other.F if isinstance(other, SparseTensor) else other))

    __iadd__ = __add__

    def __repr__(self):
        return "SparseTensor(N=%d, C=%d, tensor_stride=%d)" % (self.F.shape[0], self.F.shape[1], self.tensor_stride)


def decompose(coords, feats):
    """Split batched rows by the batch column of `coords` [N,4]: ([coords_b [Nb,3]], [feats_b [Nb,C]]) for b = 0..B-1."""
    if coords.shape[0] == 0:
        return [], []
    b = coords[:, 0].long()
    nb = int(b.max()) + 1
    order = torch.argsort(b, stable=True)                 # rows of one scene stay in their input order
    counts = torch.bincount(b, minlength=nb).tolist()
    return list(torch.split(coords[order, 1:], counts)), list(torch.split(feats[order], counts))


def cat(*tensors):
    """ME.cat (utils/minkunet.py:153-177): channel concatenation of tensors on the same coordinate map."""
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tensors[0]
    t0 = tensors[0]
    for t in tensors[1:]:
        if t.coordinate_manager is not t0.coordinate_manager or t.tensor_stride != t0.tensor_stride:
            raise ValueError("cat: tensors must share the coordinate map")
    return t0._like(torch.cat([t.F for t in tensors], 1))


class _ConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, dilation, bias, dimension, transpose):
        super().__init__()
        if dimension != 3:
            raise NotImplementedError("only dimension=3 (the reference uses D=3 everywhere)")
        if dilation != 1:
            raise NotImplementedError("dilation != 1 is never used by MinkUNet (utils/minkunet.py:39)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation, self.dimension = kernel_size, stride, dilation, dimension
        self.is_transpose = transpose
        self.kernel_volume = kernel_size ** 3
        shape = (in_channels, out_channels) if self.kernel_volume == 1 else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        # [ME-recall] MinkowskiConvolutionBase.reset_parameters: uniform(+-1/sqrt(fan)), fan = channels * kernel volume
        n = (self.out_channels if self.is_transpose else self.in_channels) * self.kernel_volume
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def extra_repr(self):
        return "in=%d, out=%d, kernel_size=%d, stride=%d" % (self.in_channels, self.out_channels, self.kernel_size, self.stride)

    # ---- inference: deferred, fused execution
    def _can_defer(self):
        """Autograd off + tensor-core mode + a shape the tcgen05 kernel takes (any cout after padding to 16; cin % 32 == 0, or
        the <= 4-channel stem through the 4-channel gather)."""
        if torch.is_grad_enabled() or resolve_mode() != "tf32" or self.kernel.device.type != "cuda":
            return False
        return self.in_channels % 32 == 0 or (self.in_channels <= 4 and self.kernel_volume > 1 and not self.is_transpose and self.stride == 1)

    def _packed_for_inference(self, bn):
        """(weight, bias, op kind, gather width) with an eval-mode BatchNorm folded in, packed for the tensor-core kernel exactly
        like MinkUNetEngine packs its program; remembered until one of the parameters changes (version counters)."""
        from ..engine import pack_conv
        parts = [self.kernel, self.bias] + ([bn.bn.weight, bn.bn.bias, bn.bn.running_mean, bn.bn.running_var] if bn is not None else [])
        key = tuple((id(t), t._version) for t in parts if t is not None)
        tag = getattr(self, "_cvb200_pack", None)
        if tag is None or tag[0] != key:
            tag = (key, pack_conv(self, bn))
            self._cvb200_pack = tag
        return tag[1]


class MinkowskiConvolution(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, dimension, False)

    def forward(self, x):
        cm, ts = x.coordinate_manager, x.tensor_stride
        defer = self._can_defer()
        if self.kernel_size == 1 and self.stride == 1:
            if defer:
                return x._deferred(_Pending(self, x.F, cm.identity_table(ts)))
            if resolve_mode() == "tf32" and self.in_channels % 32 == 0 and self.out_channels % 32 == 0:
                # training in tf32 mode: the same tensor-core kernels (forward, input and weight gradient) on an identity table
                ident = cm.identity_table(ts)
                return x._like(sparse_conv(x.F, self.kernel.unsqueeze(0), self.bias, ident, ident, "same"))
            f = x.F @ self.kernel                      # plain library GEMM (1x1x1 convolution)
            if self.bias is not None:
                f = f + self.bias
            return x._like(f)
        if self.stride == 1 and self.kernel_size % 2 == 1:
            nbr = cm.kernel_map(ts, self.kernel_size)
            if defer:
                return x._deferred(_Pending(self, x.F, nbr))
            return x._like(sparse_conv(x.F, self.kernel, self.bias, nbr, nbr, "same"))
        if self.stride == 2 and self.kernel_size == 2:
            d = cm.down(ts)
            if defer:
                return x._deferred(_Pending(self, x.F, d["children"]), 2 * ts)
            return x._like(sparse_conv(x.F, self.kernel, self.bias, d["children"], d["up_table"], "down"), 2 * ts)
        raise NotImplementedError("kernel_size=%d stride=%d is not used by the reference" % (self.kernel_size, self.stride))


class MinkowskiConvolutionTranspose(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, dimension, True)

    def forward(self, x):
        cm, ts = x.coordinate_manager, x.tensor_stride
        if not (self.stride == 2 and self.kernel_size == 2):
            raise NotImplementedError("only kernel_size=2, stride=2 (utils/minkunet.py:85-106)")
        fine = ts // 2
        if fine not in cm._down:
            raise RuntimeError("transposed convolution onto tensor stride %d: that coordinate map was never created "
                               "(the decoder re-uses the encoder's maps)" % fine)
        d = cm._down[fine]
        if self._can_defer():
            return x._deferred(_Pending(self, x.F, d["up_table"]), fine)
        return x._like(sparse_conv(x.F, self.kernel, self.bias, d["up_table"], d["children"], "up"), fine)


class MinkowskiBatchNorm(nn.Module):
    """BatchNorm1d over the [N,C] feature matrix of the whole batch [ME-recall]; `.bn` holds the parameters."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x):
        p = x._pending
        if p is not None and p.bn is None and p.residual is None and not p.relu and not self.bn.training and self.bn.track_running_stats:
            q = _Pending(p.conv, p.x, p.table)          # eval mode right after a deferred convolution: fold into its weights
            q.bn = self
            return x._deferred(q)
        return x._like(self.bn(x.F))


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        p = x._pending
        if p is not None and not p.relu:
            q = _Pending(p.conv, p.x, p.table)          # the convolution's epilogue applies it
            q.bn, q.residual, q.relu = p.bn, p.residual, True
            return x._deferred(q)
        return x._like(torch.relu(x.F))


class BasicBlock(nn.Module):
    """MinkowskiEngine.modules.resnet_block.BasicBlock [ME-recall]: conv3-norm-relu-conv3-norm (+downsample) -relu."""
    expansion = 1
    NORM_TYPE = "BN"

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=stride, dilation=dilation, dimension=dimension)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dilation=dilation, dimension=dimension)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        residual = x
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out = out + residual
        return self.relu(out)


class Bottleneck(nn.Module):
    """MinkowskiEngine.modules.resnet_block.Bottleneck [ME-recall]: conv1(1^3)-norm-relu-conv3(3^3, stride)-norm-relu-
    conv1(1^3, 4x planes)-norm (+downsample) -relu.  Imported by utils/minkunet.py:30 / utils/resnet.py:29 for
    MinkUNet50/101 and ResNet50/101, which no reference script instantiates (their PLANES are undefined upstream)."""
    expansion = 4
    NORM_TYPE = "BN"

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=1, dimension=dimension)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=stride, dilation=dilation, dimension=dimension)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv3 = MinkowskiConvolution(planes, planes * self.expansion, kernel_size=1, dimension=dimension)
        self.norm3 = MinkowskiBatchNorm(planes * self.expansion, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        residual = x
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.relu(self.norm2(self.conv2(out)))
        out = self.norm3(self.conv3(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out = out + residual
        return self.relu(out)
