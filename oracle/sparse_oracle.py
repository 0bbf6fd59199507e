"""oracle/sparse_oracle.py -- TEST INFRASTRUCTURE ONLY.

Slow, obviously-correct restatement of the sparse-convolution semantics the reference obtains from the
external MinkowskiEngine package (v0.5.3, README.md:53; call sites utils/minkunet.py:53-114).  ME is not
installable offline and the reference holds no test that pins any ME result, so this oracle is **parity
unpinned** against ME itself; it is pinned instead against DENSE convolutions (torch.nn.functional.conv3d /
conv_transpose3d on a densified scene, tests/test_oracle_sparse.py) for every (kernel, stride) variant the
U-Net uses.  Pure torch on CPU: python dict coordinate maps + index_add_.

Conventions restated from recollection of ME 0.5.x [ME-recall]:
  * kernel [K^3, Cin, Cout]; offset index k = ix + K*(iy + K*iz) (x fastest), centred for odd K,
    {0,1}^3 * tensor_stride for the stride-2 2^3 kernel;
  * stride-2 convolution: output coordinates floor(c / (2 ts)) * 2 ts, new tensor stride 2 ts;
  * transposed stride-2 convolution: output on the cached finer coordinate map, out[fine] = in[parent] @ W[k]
    with k the offset of the fine voxel inside its parent.
"""
import numpy as np
import torch


def _key(c):
    return tuple(int(v) for v in c)


def coarse_coords(coords, new_stride):
    """unique floor(c / new_stride) * new_stride in order of first appearance; returns (coarse [M,4], parent [N], koff [N])."""
    c = coords.clone().long()
    half = new_stride // 2
    par = c.clone()
    par[:, 1:] = torch.div(c[:, 1:], new_stride, rounding_mode="floor") * new_stride
    index, rows, parent = {}, [], []
    for r in par.tolist():
        k = tuple(r)
        if k not in index:
            index[k] = len(rows)
            rows.append(r)
        parent.append(index[k])
    off = (c[:, 1:] - par[:, 1:]) // half
    koff = off[:, 0] + 2 * (off[:, 1] + 2 * off[:, 2])
    return torch.tensor(rows, dtype=torch.int32), torch.tensor(parent), koff


def conv_same(coords, feats, kernel, ksize, tensor_stride=1, bias=None):
    """stride-1 K^3 convolution (K odd) on the coordinate set `coords` [N,4]."""
    index = {_key(c): i for i, c in enumerate(coords.tolist())}
    n, cout = feats.shape[0], kernel.shape[-1]
    out = torch.zeros((n, cout), dtype=feats.dtype)
    h = ksize // 2
    w = kernel.reshape(ksize ** 3, -1, cout)
    for k in range(ksize ** 3):
        ix, iy, iz = k % ksize, (k // ksize) % ksize, k // (ksize * ksize)
        d = ((ix - h) * tensor_stride, (iy - h) * tensor_stride, (iz - h) * tensor_stride)
        src, dst = [], []
        for o, c in enumerate(coords.tolist()):
            j = index.get((c[0], c[1] + d[0], c[2] + d[1], c[3] + d[2]))
            if j is not None:
                src.append(j); dst.append(o)
        if src:
            out.index_add_(0, torch.tensor(dst), feats[torch.tensor(src)] @ w[k])
    return out + (bias if bias is not None else 0)


def conv_down(coords, feats, kernel, tensor_stride=1, bias=None):
    """stride-2 2^3 convolution: returns (coarse coords, features)."""
    coarse, parent, koff = coarse_coords(coords, 2 * tensor_stride)
    out = torch.zeros((coarse.shape[0], kernel.shape[-1]), dtype=feats.dtype)
    for k in range(8):
        m = koff == k
        if m.any():
            out.index_add_(0, parent[m], feats[m] @ kernel[k])
    return coarse, out + (bias if bias is not None else 0)


def conv_up(fine_coords, coarse_feats, kernel, tensor_stride=2, bias=None):
    """transposed stride-2 2^3 convolution from the coarse level (tensor stride `tensor_stride`) onto `fine_coords`."""
    _, parent, koff = coarse_coords(fine_coords, tensor_stride)
    out = torch.zeros((fine_coords.shape[0], kernel.shape[-1]), dtype=coarse_feats.dtype)
    for k in range(8):
        m = koff == k
        if m.any():
            out[m] = coarse_feats[parent[m]] @ kernel[k]
    return out + (bias if bias is not None else 0)


def conv_table(x, w, table, bias=None):
    """The one primitive every convolution variant reduces to (canonicalvoting_b200/sparse/functional.py), restated on CPU in
    float64: out[o] = sum_k x[table[o, k]] @ w[k] (+ bias), a table entry of -1 contributing nothing."""
    x, w, table = x.detach().cpu().double(), w.detach().cpu().double(), table.detach().cpu().long()
    out = torch.zeros((table.shape[0], w.shape[2]), dtype=torch.float64)
    for k in range(table.shape[1]):
        rows = torch.nonzero(table[:, k] >= 0)[:, 0]
        if len(rows):
            out[rows] += x[table[rows, k]] @ w[k]
    return out + (bias.detach().cpu().double().view(1, -1) if bias is not None else 0)


class OracleNet:
    """Forward pass of a canonicalvoting_b200.minkunet model on CPU through the functions above, reading the
    parameters from the model's state dict (BatchNorm in the model's train/eval mode)."""

    def __init__(self, model):
        self.m = model

    def _p(self, t):
        """A parameter of the model as the CPU tensor the restatement computes with."""
        return t.detach().cpu()

    def _bn(self, mod, f):
        bn = mod.bn
        return torch.nn.functional.batch_norm(f, bn.running_mean.cpu().clone(), bn.running_var.cpu().clone(), self._p(bn.weight),
                                              self._p(bn.bias), bn.training, bn.momentum, bn.eps)

    def _conv(self, mod, coords_by_ts, ts, f):
        w = self._p(mod.kernel)
        b = self._p(mod.bias) if mod.bias is not None else None
        if mod.kernel_size == 1:
            return ts, f @ w + (b if b is not None else 0)
        if mod.is_transpose:
            return ts // 2, conv_up(coords_by_ts[ts // 2], f, w, ts, b)
        if mod.stride == 2:
            coarse, out = conv_down(coords_by_ts[ts], f, w, ts, b)
            coords_by_ts[2 * ts] = coarse
            return 2 * ts, out
        return ts, conv_same(coords_by_ts[ts], f, w, mod.kernel_size, ts, b)

    def _block(self, blk, C, ts, f):
        r = f
        _, o = self._conv(blk.conv1, C, ts, f)
        o = torch.relu(self._bn(blk.norm1, o))
        _, o = self._conv(blk.conv2, C, ts, o)
        o = self._bn(blk.norm2, o)
        if blk.downsample is not None:
            _, r = self._conv(blk.downsample[0], C, ts, f)
            r = self._bn(blk.downsample[1], r)
        return torch.relu(o + r)

    def forward(self, coords, feats):
        from canonicalvoting_b200.minkunet import _DECODER, _ENCODER
        m, C = self.m, {1: coords.cpu()}
        ts, f = self._conv(m.conv0p1s1, C, 1, feats.cpu())
        f = torch.relu(self._bn(m.bn0, f))
        skips = [f]
        for conv, bn, block in _ENCODER:
            ts, f = self._conv(getattr(m, conv), C, ts, f)
            f = torch.relu(self._bn(getattr(m, bn), f))
            for blk in getattr(m, block):
                f = self._block(blk, C, ts, f)
            skips.append(f)
        skips.pop()
        for conv, bn, block in _DECODER:
            ts, f = self._conv(getattr(m, conv), C, ts, f)
            f = torch.relu(self._bn(getattr(m, bn), f))
            f = torch.cat([f, skips.pop()], 1)
            for blk in getattr(m, block):
                f = self._block(blk, C, ts, f)
        return self._conv(m.final, C, ts, f)[1]


# ----------------------------------------------------------------------------------------------------------------
# Vectorised CPU port (the timed `cpu_baseline` of the sparse-U-Net half in bench.py): same semantics as the
# functions above, but kernel maps come from sorted packed keys + searchsorted and every offset is one
# gather -> GEMM -> index_add_ -- which is how MinkowskiEngine's own CPU backend evaluates a convolution
# [ME-recall].  tests/test_oracle_sparse.py checks it against the dict-based functions.
def _pack(c):
    c = c.long()
    return ((c[:, 0] & 0xffff) << 48) | (((c[:, 1] + 32768) & 0xffff) << 32) | (((c[:, 2] + 32768) & 0xffff) << 16) | \
        ((c[:, 3] + 32768) & 0xffff)


class FastMaps:
    """Per-scene coordinate levels and kernel maps (cached like ME's coordinate manager)."""

    def __init__(self, coords):
        self.levels = {1: coords.int()}
        self._sorted, self._maps, self._down = {}, {}, {}

    def _lookup(self, ts, query):
        if ts not in self._sorted:
            keys = _pack(self.levels[ts])
            order = torch.argsort(keys)
            self._sorted[ts] = (keys[order], order)
        skeys, order = self._sorted[ts]
        q = _pack(query)
        pos = torch.searchsorted(skeys, q).clamp_(max=len(skeys) - 1)
        hit = skeys[pos] == q
        return order[pos], hit

    def same(self, ts, K):
        key = (ts, K)
        if key not in self._maps:
            c, h, pairs = self.levels[ts], K // 2, []
            for k in range(K ** 3):
                d = torch.tensor([0, (k % K - h) * ts, ((k // K) % K - h) * ts, (k // (K * K) - h) * ts], dtype=torch.int32)
                j, hit = self._lookup(ts, c + d)
                o = torch.nonzero(hit)[:, 0]
                pairs.append((j[o], o))
            self._maps[key] = pairs
        return self._maps[key]

    def down(self, ts):
        if ts not in self._down:
            c = self.levels[ts].long()
            par = c.clone()
            par[:, 1:] = torch.div(c[:, 1:], 2 * ts, rounding_mode="floor") * 2 * ts
            keys = _pack(par)
            uniq, inverse = torch.unique(keys, return_inverse=True)
            first = torch.full((len(uniq),), len(keys), dtype=torch.long).scatter_reduce_(0, inverse, torch.arange(len(keys)), "amin")
            order = torch.argsort(first)                      # number coarse voxels by their first child
            rank = torch.empty_like(order); rank[order] = torch.arange(len(order))
            parent = rank[inverse]
            self.levels[2 * ts] = par[first[order]].int()
            off = (c[:, 1:] - par[:, 1:]) // ts
            self._down[ts] = (parent, off[:, 0] + 2 * (off[:, 1] + 2 * off[:, 2]))
        return self._down[ts]


class FastCpuNet(OracleNet):
    """OracleNet with vectorised maps (float32, torch CPU threads)."""

    def _conv(self, mod, maps, ts, f):
        w = self._p(mod.kernel)
        b = self._p(mod.bias) if mod.bias is not None else None
        cout = w.shape[-1]
        if mod.kernel_size == 1:
            return ts, f @ w + (b if b is not None else 0)
        if mod.is_transpose:
            parent, koff = maps.down(ts // 2)
            out = torch.zeros((len(parent), cout), dtype=f.dtype)
            for k in range(8):
                m = torch.nonzero(koff == k)[:, 0]
                if len(m):
                    out[m] = f[parent[m]] @ w[k]
            return ts // 2, out + (b if b is not None else 0)
        if mod.stride == 2:
            parent, koff = maps.down(ts)
            out = torch.zeros((maps.levels[2 * ts].shape[0], cout), dtype=f.dtype)
            for k in range(8):
                m = torch.nonzero(koff == k)[:, 0]
                if len(m):
                    out.index_add_(0, parent[m], f[m] @ w[k])
            return 2 * ts, out + (b if b is not None else 0)
        out = torch.zeros((f.shape[0], cout), dtype=f.dtype)
        for k, (src, dst) in enumerate(maps.same(ts, mod.kernel_size)):
            if len(src):
                out.index_add_(0, dst, f[src] @ w[k])
        return ts, out + (b if b is not None else 0)

    def forward(self, coords, feats):
        self_maps = FastMaps(coords.cpu())
        return OracleNet.forward_with(self, self_maps, feats)


class GradCpuNet(FastCpuNet):
    """FastCpuNet whose parameters are float64 autograd leaves (copies of the model's): `forward` is then differentiable
    by plain torch autograd, which gives reference parameter gradients for the training step (train_joint.py:284-286)
    without any hand-written backward.  `grads()` maps the model's parameter names to the leaves' gradients."""

    def __init__(self, model):
        super().__init__(model)
        self.leaves = {}

    def _p(self, t):
        leaf = self.leaves.get(id(t))
        if leaf is None:
            leaf = t.detach().cpu().double().requires_grad_(True)
            self.leaves[id(t)] = leaf
        return leaf

    def _bn(self, mod, f):
        bn = mod.bn
        return torch.nn.functional.batch_norm(f, bn.running_mean.cpu().double().clone(), bn.running_var.cpu().double().clone(),
                                              self._p(bn.weight), self._p(bn.bias), bn.training, bn.momentum, bn.eps)

    def grads(self):
        return {name: self.leaves[id(p)].grad for name, p in self.m.named_parameters() if id(p) in self.leaves}


def _forward_with(self, C, feats):
    from canonicalvoting_b200.minkunet import _DECODER, _ENCODER
    m = self.m
    ts, f = self._conv(m.conv0p1s1, C, 1, feats.cpu())
    f = torch.relu(self._bn(m.bn0, f))
    skips = [f]
    for conv, bn, block in _ENCODER:
        ts, f = self._conv(getattr(m, conv), C, ts, f)
        f = torch.relu(self._bn(getattr(m, bn), f))
        for blk in getattr(m, block):
            f = self._block(blk, C, ts, f)
        skips.append(f)
    skips.pop()
    for conv, bn, block in _DECODER:
        ts, f = self._conv(getattr(m, conv), C, ts, f)
        f = torch.relu(self._bn(getattr(m, bn), f))
        f = torch.cat([f, skips.pop()], 1)
        for blk in getattr(m, block):
            f = self._block(blk, C, ts, f)
    return self._conv(m.final, C, ts, f)[1]


OracleNet.forward_with = _forward_with
OracleNet.forward = lambda self, coords, feats: _forward_with(self, {1: coords.cpu()}, feats)
