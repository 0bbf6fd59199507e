"""MinkUNetEngine -- inference executor for the reference's sparse U-Net (eval path: eval_joint.py:151-190).

The module-by-module path (canonicalvoting_b200.sparse / the `MinkowskiEngine` compat package) mirrors the
reference's Python structure and is what training uses.  For inference this engine compiles the SAME
network (any `MinkUNetBase`-shaped model: ours or the reference's own utils/minkunet.py class running on the
compat package) into a flat program of fused convolution ops executed back to back by
`cvb200_sc_run_program` (csrc/sparse_engine.cu):

  * BatchNorm (eval mode) is folded into the convolution weights and a bias  (utils/minkunet.py:123-124 ...);
  * ReLU and the residual add of BasicBlock run in the convolution epilogue;
  * `ME.cat(out, skip)` (utils/minkunet.py:153-177) costs nothing: the encoder writes its skip tensor and the
    transposed convolution writes its output into column slices of one pre-allocated buffer;
  * 1x1x1 convolutions (block down-samples, `final`) use the same tensor-core kernel with an identity table;
  * the 3-channel 5^3 stem runs in the same kernel too (input padded to 4 channels, 8 neighbours gathered per k-block);
  * every coordinate level and kernel map of the scene comes from ONE enqueue + ONE host synchronisation
    (cvb200_sc_build_maps); `prefetch()` moves that (and the upload of host tensors) to a worker thread + side stream so
    that it overlaps the previous scene;
  * the head decode (eval_joint.py:173-190) is one kernel, optionally also emitting scan_points = coords * res (:193).

    engine = MinkUNetEngine(model, pipeline=True)      # after load_state_dict(...), model.eval()
    feats = engine(coords_int32_cuda, feats_cuda)      # == model(ME.SparseTensor(feats, coords)).F up to TF32 rounding
    xyz, scale, class_pred, prob = engine.predict(coords, feats)
    nxt = engine.prefetch(next_coords, next_feats)     # host or device tensors; overlaps the calls above
    xyz, scale, class_pred, prob, points = engine.predict(None, None, nxt, res=0.03)
"""
import collections
import ctypes
import os
import threading

import torch

from . import _lib
from .sparse.coords import CoordinateManager, _ptr, _stream

CONV_OPTIONS = [1, 1]      # (allow_split, launch bits) the library runs with: set_conv_options() keeps it in step

_ENCODER = [("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"), ("conv3p4s2", "bn3", "block3"),
            ("conv4p8s2", "bn4", "block4")]
_DECODER = [("convtr4p16s2", "bntr4", "block5"), ("convtr5p8s2", "bntr5", "block6"), ("convtr6p4s2", "bntr6", "block7"),
            ("convtr7p2s2", "bntr7", "block8")]


def set_conv_options(allow_split=1, launch_bits=1):
    """cvb200_sc_set_conv_options for the whole process (measurement switch; the SceneGraph capture restores these values)."""
    CONV_OPTIONS[0], CONV_OPTIONS[1] = int(allow_split), int(launch_bits)
    _lib.load().cvb200_sc_set_conv_options(int(allow_split), int(launch_bits))


def _fold(conv, bn):
    """(weight [k3, cin, cout] with the BatchNorm scale folded in, bias [cout] or None) as float32 tensors."""
    w = conv.kernel.detach().float()
    if w.dim() == 2:
        w = w.unsqueeze(0)
    b = conv.bias.detach().float().view(-1) if conv.bias is not None else None
    if bn is not None:
        n = bn.bn
        g = n.weight.detach().float() / torch.sqrt(n.running_var.float() + n.eps)
        w = w * g.view(1, 1, -1)
        shift = n.bias.detach().float() - n.running_mean.float() * g
        b = shift if b is None else b * g + shift
    return w, b


def pack_conv(conv, bn, gather4=True):
    """(weight, bias, op kind, gather width) of one convolution (+ folded eval-mode BatchNorm) as cvb200_sc_run_program takes it:
    output channels padded to a multiple of 16; cin % 32 == 0: kind 0, weight [k3, cout, cin]; a <= 4-channel input (the 5^3
    stem): kind 3, the input padded to 4 channels, 8 neighbours x 4 channels per k-block, weight [1, cout, 32 ceil(k3 / 8)];
    anything else: None (the caller falls back)."""
    w, b = _fold(conv, bn)
    cin = w.shape[1]
    if w.shape[2] % 16:      # the tensor-core kernel needs cout % 16 == 0 (the 8-channel heads of the per-category models,
        pad = 16 - w.shape[2] % 16          # eval_separate.py:138): zero output channels, cut off again by the caller
        w = torch.nn.functional.pad(w, (0, pad))
        b = torch.nn.functional.pad(b, (0, pad)) if b is not None else None
    b = b.contiguous() if b is not None else None
    if cin % 32 == 0:
        return w.transpose(1, 2).contiguous(), b, 0, 0
    if cin <= 4 and gather4:
        k3, cout = w.shape[0], w.shape[2]
        kp = 32 * ((k3 + 7) // 8)
        w4 = torch.zeros((kp // 4, 4, cout), dtype=w.dtype, device=w.device)
        w4[:k3, :cin] = w
        return w4.reshape(1, kp, cout).transpose(1, 2).contiguous(), b, 3, k3
    return None


class _Slice:
    """Column slice [c0, c0+c) of a row-major float32 matrix [rows, ld] that lives at byte address `base`."""
    __slots__ = ("base", "rows", "ld", "c0", "c")

    def __init__(self, base, rows, ld, c0, c):
        self.base, self.rows, self.ld, self.c0, self.c = base, rows, ld, c0, c

    @property
    def ptr(self):
        return self.base + 4 * self.c0

    def cols(self, c0, c):
        return _Slice(self.base, self.rows, self.ld, self.c0 + c0, c)


class _Arena:
    """One allocation per forward pass for every activation buffer of the program (sizes are known once the
    coordinate levels are): matrices are carved out at 256-byte aligned offsets."""

    def __init__(self, device):
        self.device, self.items, self.total = device, [], 0

    def matrix(self, rows, cols):
        sl = _Slice(self.total, rows, cols, 0, cols)        # base = offset for now, rebased in commit()
        self.items.append(sl)
        self.total += (rows * cols * 4 + 255) // 256 * 256
        return sl

    def commit(self):
        self.buf = torch.empty(max(self.total, 256), dtype=torch.uint8, device=self.device)
        base = self.buf.data_ptr()
        for sl in self.items:
            sl.base += base
        return self.buf

    def tensor(self, sl):
        off = sl.base - self.buf.data_ptr()
        return self.buf[off:off + sl.rows * sl.ld * 4].view(torch.float32).view(sl.rows, sl.ld)


class MinkUNetEngine:
    def __init__(self, model, nclasses=9, log_scale=True, pipeline=False):
        """pipeline=True builds the coordinate maps of a scene on a side stream: their (few) host synchronisations
        then wait only for the map kernels, not for the previous scene's convolutions still running on the caller's
        stream, so consecutive scenes overlap.  The caller must then hand in inputs that are already complete
        (e.g. produced by `upload`), because the side stream does not wait for the caller's stream."""
        self.nclasses, self.log_scale, self.pipeline = nclasses, log_scale, pipeline
        self.model = model
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("MinkUNetEngine needs the model on a CUDA device (no CPU path)")
        # high priority: the small map kernels must not queue behind the pending CTAs of a programmatically launched convolution
        self._side = torch.cuda.Stream(self.device, priority=-1) if pipeline else None
        self._ring = collections.deque()      # (tensors used by launches in flight, completion event)
        self._pool = None                     # one worker thread for prefetch()
        self._pinned = None                   # count read-back buffer of the fused map builder (one build at a time)
        self._cm_building = None
        self._maps_lock = threading.Lock()
        self.fused_maps = True                # cvb200_sc_build_maps (one sync per scene); False: step-by-step coordinate manager
        self.stem_gather4 = os.environ.get("CVB200_STEM_IM2COL", "0") != "1"   # False: stem as im2col + product (refresh() after changing)
        self.refresh()

    def upload(self, coords_host, feats_host):
        """Host (ideally pinned) -> device copies on the engine's map stream; returns device tensors ready for __call__."""
        if self._side is None:
            return coords_host.to(self.device, non_blocking=True), feats_host.to(self.device, non_blocking=True)
        with torch.cuda.stream(self._side):
            return coords_host.to(self.device, non_blocking=True), feats_host.to(self.device, non_blocking=True)

    def prefetch(self, coords, feats=None):
        """Start building the coordinate maps of a scene (and, for host tensors, its upload) on a worker thread + the
        engine's side stream while the caller keeps launching the previous scene: the maps cost ~1.2 ms of host time
        per 50k-voxel scene (launches + one count read-back per level), more than everything else together.  Returns a
        handle to pass to __call__ / predict as `maps=`.  Needs pipeline=True."""
        if self._side is None:
            raise RuntimeError("MinkUNetEngine.prefetch needs pipeline=True")
        if self._pool is None:
            import concurrent.futures
            self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="cvb200-maps")

        def job():
            with torch.cuda.device(self.device), torch.cuda.stream(self._side):
                c = coords.to(self.device, non_blocking=True) if not coords.is_cuda else coords
                f = feats.to(self.device, non_blocking=True) if (feats is not None and not feats.is_cuda) else feats
                c = c.to(torch.int32).contiguous()
                cm = self.build_maps(c)
                ready = self._side.record_event()
            return c, f, cm, ready
        return self._pool.submit(job)

    def refresh(self):
        """(Re)pack the model's parameters: call again after the weights changed."""
        m = self.model
        self.w, self.im2col, self.gather4 = {}, {}, {}

        def put(name, conv, bn):
            w, b = _fold(conv, bn)
            cin = w.shape[1]
            if w.shape[2] % 16:      # the tensor-core kernel needs cout % 16 == 0 (the 8-channel heads of the per-category models,
                pad = 16 - w.shape[2] % 16          # eval_separate.py:138): zero output channels, cut off again by build()
                w = torch.nn.functional.pad(w, (0, pad))
                b = torch.nn.functional.pad(b, (0, pad)) if b is not None else None
            if cin % 32 == 0:       # tensor-core op: [k3, cout, cin]
                self.w[name] = (w.transpose(1, 2).contiguous(), b.contiguous() if b is not None else None, 0)
            elif cin <= 4 and self.stem_gather4:
                # small input width (the 3-channel 5^3 stem): the input is padded to 4 channels and the convolution kernel
                # gathers 8 neighbours x 4 channels per k-block (CVB200_OP_CONV_TC_GATHER4): w4[co][4 k + c], K = 32 ceil(k3 / 8)
                k3, cout = w.shape[0], w.shape[2]
                kp = 32 * ((k3 + 7) // 8)
                w4 = torch.zeros((kp // 4, 4, cout), dtype=w.dtype, device=w.device)
                w4[:k3, :cin] = w
                self.w[name] = (w4.reshape(1, kp, cout).transpose(1, 2).contiguous(), b.contiguous() if b is not None else None, 3)
                self.gather4[name] = (k3, cin, kp)
            else:
                # small input width, fallback: im2col + one [N, K^3*cin -> pad 32] x [., cout] tensor-core product
                k3, cout = w.shape[0], w.shape[2]
                kp = (k3 * cin + 31) // 32 * 32
                kp = (kp + cin - 1) // cin * cin if kp % cin else kp
                while kp % 32 or kp % cin:
                    kp += 1
                wp = torch.zeros((1, kp, cout), dtype=w.dtype, device=w.device)
                wp[0, :k3 * cin] = w.reshape(k3 * cin, cout)
                self.w[name] = (wp.transpose(1, 2).contiguous(), b.contiguous() if b is not None else None, 0)
                self.im2col[name] = (k3, cin, kp)

        put("conv0p1s1", m.conv0p1s1, m.bn0)
        for conv, bn, block in _ENCODER + _DECODER:
            put(conv, getattr(m, conv), getattr(m, bn))
            for i, blk in enumerate(getattr(m, block)):
                put("%s.%d.conv1" % (block, i), blk.conv1, blk.norm1)
                put("%s.%d.conv2" % (block, i), blk.conv2, blk.norm2)
                if blk.downsample is not None:
                    put("%s.%d.down" % (block, i), blk.downsample[0], blk.downsample[1])
        put("final", m.final, None)
        self.planes = {block: [getattr(m, block)[0].conv1.out_channels, len(getattr(m, block))] for _, _, block in _ENCODER + _DECODER}
        self.init_dim = m.conv0p1s1.out_channels
        self.out_channels = m.final.out_channels

    # ------------------------------------------------------------------ program construction
    def _op(self, ops, name, src, dst, table, relu, residual=None, out_ts=1):
        """Append one fused convolution; slice base addresses are arena offsets until the arena is committed, so the
        pointer fields are filled in by build() afterwards (ops keep references to their slices).  `out_ts` = tensor stride of
        the OUTPUT rows: with a size-agnostic coordinate manager the op reads that level's row count from device memory."""
        w, b, kind = self.w[name]
        k3 = w.shape[0]
        cin, cout = (w.shape[2], w.shape[1]) if kind == 0 else (w.shape[1], w.shape[2])
        assert src.c == cin and dst.c == cout and table.shape[1] == k3, (name, src.c, cin, dst.c, cout, table.shape, k3)
        o = _lib.ScOp()
        o.kind, o.cin, o.cout, o.k3 = kind, cin, cout, k3
        o.ldi, o.ldo, o.ldr, o.relu = src.ld, dst.ld, residual.ld if residual is not None else 0, 1 if relu else 0
        o.n_out = table.shape[0]
        o.n_in = src.rows
        o.w, o.bias, o.table = w.data_ptr(), b.data_ptr() if b is not None else None, table.data_ptr()
        o.n_out_dev = self._count_ptr(out_ts)
        ops.append((o, src, dst, residual))

    def _count_ptr(self, ts):
        cm = self._cm_building
        return cm.count_ptr(ts.bit_length() - 1) if getattr(cm, "static", False) else None

    def _blocks(self, ops, arena, block, cm, ts, x, out_slice=None):
        """BasicBlocks of one stage; the last block writes into `out_slice` (a skip slot) when given."""
        planes, nblk = self.planes[block]
        n = cm.levels[ts].n
        nbr = cm.kernel_map(ts, 3)
        ident = self._identity(cm, ts)
        for i in range(nblk):
            t1 = arena.matrix(n, planes)
            self._op(ops, "%s.%d.conv1" % (block, i), x, t1, nbr, relu=True, out_ts=ts)
            res = x
            if ("%s.%d.down" % (block, i)) in self.w:
                res = arena.matrix(n, planes)
                self._op(ops, "%s.%d.down" % (block, i), x, res, ident, relu=False, out_ts=ts)
            dst = out_slice if (i == nblk - 1 and out_slice is not None) else arena.matrix(n, planes)
            self._op(ops, "%s.%d.conv2" % (block, i), t1, dst, nbr, relu=True, residual=res, out_ts=ts)
            x = dst
        return x

    def _identity(self, cm, ts):
        key = ("ident", ts)
        t = cm._nbr.get(key)
        if t is None:
            t = torch.arange(cm.levels[ts].n, dtype=torch.int32, device=self.device).view(-1, 1).contiguous()
            cm._nbr[key] = t
        return t

    def build_maps(self, coords):
        """Coordinate levels + every neighbour table the network needs (device work + one scalar read per level)."""
        if self.fused_maps:
            # one pinned read-back buffer per engine: builds (caller thread, prefetch worker) take turns
            with self._maps_lock:
                if self._pinned is None:
                    self._pinned = torch.empty(8, dtype=torch.int32).pin_memory()
                return CoordinateManager.build_unet(coords, int(self.model.conv0p1s1.kernel_size), 4, self._pinned)
        cm = CoordinateManager(coords)
        for ts in (1, 2, 4, 8):
            cm.down(ts)
        cm.kernel_map(1, self.model.conv0p1s1.kernel_size)
        for ts in (1, 2, 4, 8, 16):
            cm.kernel_map(ts, 3)
            self._identity(cm, ts)
        return cm

    def can_fuse_decode(self):
        """The head decode can run in the epilogue of `final` (cvb200_decode_args): joint head with 9 classes = 64 channels."""
        return self.nclasses == 9 and self.out_channels == 64 and self.w["final"][0].shape[1] == 64

    def build(self, coords, feats, cm=None, decode=None):
        """Buffers + program for one batch of scenes. Returns (ops array, output tensor [N, Cout], keep-alive list).
        `decode` (_lib.DecodeArgs): the last convolution decodes its rows instead of storing them (the output tensor then
        stays unwritten)."""
        if cm is None:
            cm = self.build_maps(coords)
        self._cm_building = cm
        ts_list = [1, 2, 4, 8, 16]
        n = {ts: cm.levels[ts].n for ts in ts_list}
        P = self.planes
        arena, ops = _Arena(self.device), []
        feats = feats.contiguous()
        # concat buffers of the decoder: [transposed-conv output | encoder skip]
        tr_out = {8: self.w["convtr4p16s2"][0].shape[1], 4: self.w["convtr5p8s2"][0].shape[1],
                  2: self.w["convtr6p4s2"][0].shape[1], 1: self.w["convtr7p2s2"][0].shape[1]}
        skip_c = {8: P["block3"][0], 4: P["block2"][0], 2: P["block1"][0], 1: self.init_dim}
        cat = {ts: arena.matrix(n[ts], tr_out[ts] + skip_c[ts]) for ts in (1, 2, 4, 8)}
        skip = {ts: cat[ts].cols(tr_out[ts], skip_c[ts]) for ts in cat}
        arena.items += list(skip.values())
        # stem: conv0 (5^3, stride 1) + bn0 + relu -> skip slot of level 1
        src = _Slice(feats.data_ptr(), feats.shape[0], feats.shape[1], 0, feats.shape[1])
        stem_table = cm.kernel_map(1, self.model.conv0p1s1.kernel_size)
        if "conv0p1s1" in self.gather4:
            k3, cin, kp = self.gather4["conv0p1s1"]
            # [N, 4]: one 16-byte vector per voxel (a caller that already holds the padded matrix passes it as is)
            feats4 = feats if feats.shape[1] == 4 and cin < 4 else torch.nn.functional.pad(feats, (0, 4 - cin)).contiguous()
            w, b, _ = self.w["conv0p1s1"]
            o = _lib.ScOp()
            o.kind, o.cin, o.cout, o.k3, o.ldi, o.ldo, o.ldr, o.relu = 3, kp, w.shape[1], k3, 4, skip[1].ld, 0, 1
            o.n_out, o.n_in = n[1], feats4.shape[0]
            o.w, o.bias, o.table = w.data_ptr(), b.data_ptr() if b is not None else None, stem_table.data_ptr()
            o.n_out_dev = self._count_ptr(1)
            src4 = _Slice(feats4.data_ptr(), feats4.shape[0], 4, 0, 4)
            ops.append((o, src4, skip[1], None))
            feats = feats4                     # kept alive with the program
        elif "conv0p1s1" in self.im2col:
            k3, cin, kp = self.im2col["conv0p1s1"]
            col = arena.matrix(n[1], kp)
            o = _lib.ScOp()
            o.kind, o.cin, o.cout, o.k3, o.ldi, o.ldo, o.n_out, o.n_in = 2, cin, kp, k3, src.ld, kp, n[1], src.rows
            o.table = stem_table.data_ptr()
            ops.append((o, src, col, None))
            self._op(ops, "conv0p1s1", col, skip[1], self._identity(cm, 1), relu=True)
        else:
            self._op(ops, "conv0p1s1", src, skip[1], stem_table, relu=True)
        x, ts = skip[1], 1
        for (conv, bn, block), skip_ts in zip(_ENCODER, (2, 4, 8, None)):
            d = cm.down(ts)
            y = arena.matrix(n[2 * ts], self.w[conv][0].shape[1])
            self._op(ops, conv, x, y, d["children"], relu=True, out_ts=2 * ts)
            ts *= 2
            x = self._blocks(ops, arena, block, cm, ts, y, skip[skip_ts] if skip_ts is not None else None)
        for (conv, bn, block), fine in zip(_DECODER, (8, 4, 2, 1)):
            d = cm._down[fine]
            up = cat[fine].cols(0, tr_out[fine])
            arena.items.append(up)
            self._op(ops, conv, x, up, d["up_table"], relu=True, out_ts=fine)
            ts = fine
            x = self._blocks(ops, arena, block, cm, ts, cat[fine])
        out = arena.matrix(n[1], self.w["final"][0].shape[1])          # padded to a multiple of 16 channels
        self._op(ops, "final", x, out, self._identity(cm, 1), relu=False, out_ts=1)
        if decode is not None:
            if not self.can_fuse_decode():
                raise RuntimeError("fused head decode needs the joint head: 9 classes, 64 output channels")
            ops[-1][0].decode = ctypes.pointer(decode)
        self._cm_building = None
        buf = arena.commit()
        arr = (_lib.ScOp * len(ops))()
        for i, (o, src_, dst_, res_) in enumerate(ops):
            o.in_, o.out = src_.ptr, dst_.ptr
            o.residual = res_.ptr if res_ is not None else None
            arr[i] = o
        return arr, arena.tensor(out)[:, :self.out_channels], [cm, feats, buf, decode]

    # ------------------------------------------------------------------ execution
    def __call__(self, coords, feats, maps=None):
        """coords int32 [N,4] (batch,x,y,z) and feats float32 [N,Cin] on the device -> features [N, Cout].
        `maps`: handle returned by prefetch() for these coordinates (then coords / feats may be None: the uploaded ones are used)."""
        L = _lib.load()
        if maps is not None:
            c_, f_, cm, ready = maps.result()
            coords = c_
            feats = f_ if f_ is not None else feats
        if not (coords.is_cuda and feats.is_cuda):
            raise RuntimeError("MinkUNetEngine: CUDA tensors expected (there is no CPU path)")
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            coords = coords.to(torch.int32).contiguous()
            if maps is not None:
                main.wait_event(ready)
            elif self._side is not None:
                with torch.cuda.stream(self._side):
                    cm = self.build_maps(coords)
                    ready = self._side.record_event()
                main.wait_event(ready)
            else:
                cm = self.build_maps(coords)
            arr, out, keep = self.build(coords, feats.float(), cm)
            self.execute(arr, keep)
        return out

    def execute(self, arr, keep):
        """Launch a program returned by build() on the current stream (asynchronous); `keep` stays referenced until the
        launches are known to have completed."""
        L = _lib.load()
        rc = L.cvb200_sc_run_program(arr, len(arr), _stream())
        _lib.check(rc, "cvb200_sc_run_program")
        self._ring.append((keep, torch.cuda.current_stream().record_event()))
        while len(self._ring) > 3:
            old_keep, done = self._ring.popleft()
            done.synchronize()
            del old_keep

    def decode(self, feats, coords=None, res=None):
        """Head decode (eval_joint.py:173-190): features [N, 7*C+1] -> (xyz_pred, scale_pred, class_pred int64, prob_pred).
        With `coords` (int32 [N,4]) and `res` the same kernel also returns scan_points = coords[:, 1:] * res (eval_joint.py:193)
        as a fifth tensor: the first argument of hv_cuda.forward."""
        L = _lib.load()
        n = feats.shape[0]
        xyz = torch.empty((n, 3), dtype=torch.float32, device=feats.device)
        scale = torch.empty((n, 3), dtype=torch.float32, device=feats.device)
        cls = torch.empty((n,), dtype=torch.int64, device=feats.device)
        prob = torch.empty((n,), dtype=torch.float32, device=feats.device)
        with torch.cuda.device(feats.device):
            if coords is not None:
                coords = coords.to(torch.int32).contiguous()
                points = torch.empty((n, 3), dtype=torch.float32, device=feats.device)
                rc = L.cvb200_head_decode_points(_ptr(feats), feats.stride(0), n, self.nclasses, 1 if self.log_scale else 0, _ptr(xyz),
                                                 _ptr(scale), _ptr(cls), _ptr(prob), _ptr(coords), ctypes.c_float(float(res)), _ptr(points),
                                                 _stream())
                _lib.check(rc, "cvb200_head_decode_points")
                return xyz, scale, cls, prob, points
            rc = L.cvb200_head_decode(_ptr(feats), feats.stride(0), n, self.nclasses, 1 if self.log_scale else 0, _ptr(xyz),
                                      _ptr(scale), _ptr(cls), _ptr(prob), _stream())
            _lib.check(rc, "cvb200_head_decode")
        return xyz, scale, cls, prob

    def decode_separate(self, feats):
        """Head decode of a per-category model (eval_separate.py:169-182): 8 channels = xyz[3] | scale[3] | objectness logits[2]
        -> (xyz_pred, scale_pred (exp() if log_scale), prob_pred = softmax(logits)[:, 1]).  Several per-category engines can
        share ONE coordinate-map handle of a scene (`maps=` of prefetch()), which is what makes the 9-models-per-scene eval cheap."""
        if feats.shape[1] != 8:
            raise RuntimeError("decode_separate: 8 output channels expected (xyz 3, scale 3, objectness 2)")
        xyz = feats[:, :3].contiguous()
        scale = torch.exp(feats[:, 3:6]) if self.log_scale else feats[:, 3:6].contiguous()
        prob = torch.softmax(feats[:, 6:8], dim=-1)[:, 1].contiguous()
        return xyz, scale.contiguous(), prob

    def graph_lane(self, n_voxels, vote=None, fuse_decode=None):
        """A `SceneGraph` for scenes of exactly `n_voxels` voxels: persistent input / table / activation / output buffers and ONE
        captured CUDA graph holding the coordinate-map builder, the whole convolution program, the head decode (+ scan_points)
        and -- with `vote=dict(res=, num_rots=, corner=, dims=)` -- the vote op.  `fuse_decode` (default: whenever the head
        allows it) runs the decode in the epilogue of the last convolution; out["feats"] is then not produced.  See SceneGraph."""
        return SceneGraph(self, int(n_voxels), vote, self.can_fuse_decode() if fuse_decode is None else bool(fuse_decode))

    def predict(self, coords, feats, maps=None, res=None):
        """Network + head decode.  With `res` the tuple has a fifth element: scan_points = coords[:, 1:] * res."""
        if res is None:
            return self.decode(self(coords, feats, maps))
        if maps is not None:
            coords = maps.result()[0]
        return self.decode(self(coords, feats, maps), coords, res)


class SceneGraph:
    """One scene = one CUDA-graph launch, no host synchronisation, no per-scene host work beyond two input copies.

    The reference's per-scene host loop (eval_joint.py:160-193: build the sparse tensor, ~2000 MinkowskiEngine launches, ten
    torch ops of decode glue) became, in round 1, ~100 launches + one count read-back issued by two Python threads per scene
    -- still ~1.8 ms of host time per 1.5 ms of GPU time, which is what limited 8 ranks on one 32-core host.  Here nothing
    the host enqueues depends on the scene: tables and activations are sized for the upper bound (a coarse level never has
    more voxels than the input), the real level sizes stay in device memory where the map builder writes them and every
    convolution reads its own row count and plans its tiles itself (cvb200_sc_op.n_out_dev).  So the whole sequence is
    captured once per voxel count and replayed.

        lane = engine.graph_lane(50000, vote=dict(res=0.03, num_rots=12, corner=(0, 0, 0), dims=(128, 128, 128)))
        out = lane.run(coords, feats)          # host (pinned) or device tensors; asynchronous on the current stream
        out["xyz"], out["scale"], out["class_pred"], out["prob"], out["points"], out["grids"], out["feats"]

    The returned tensors are the lane's persistent buffers: they are overwritten by the lane's next run (stream-ordered), so
    use one lane per scene in flight.  Memory is what the upper bounds cost: ~2 GB per lane at 50 000 voxels."""

    def __init__(self, engine, n, vote=None, fuse_decode=False):
        from . import hv_cuda as H
        self.engine, self.n, self.vote, self.fuse_decode = engine, n, vote, fuse_decode
        dev = engine.device
        L = _lib.load()
        with torch.cuda.device(dev):
            self.coords = torch.zeros((n, 4), dtype=torch.int32, device=dev)
            self.feats_in = torch.zeros((n, engine.model.conv0p1s1.in_channels), dtype=torch.float32, device=dev)
            pad4 = "conv0p1s1" in engine.gather4
            self.feats4 = torch.zeros((n, 4), dtype=torch.float32, device=dev) if pad4 else self.feats_in
            self.cm = CoordinateManager.static_unet(self.coords, int(engine.model.conv0p1s1.kernel_size), 4)
            f32 = dict(dtype=torch.float32, device=dev)
            self.xyz, self.scale = torch.empty((n, 3), **f32), torch.empty((n, 3), **f32)
            self.cls, self.prob = torch.empty((n,), dtype=torch.int64, device=dev), torch.empty((n,), **f32)
            self.points = torch.empty((n, 3), **f32)
            self.res = float(vote["res"]) if vote else 0.03
            dec = None
            if fuse_decode:
                dec = _lib.DecodeArgs()
                dec.xyz, dec.scale, dec.class_pred, dec.prob = self.xyz.data_ptr(), self.scale.data_ptr(), self.cls.data_ptr(), self.prob.data_ptr()
                dec.coords, dec.points, dec.res = self.coords.data_ptr(), self.points.data_ptr(), self.res
                dec.nclasses, dec.log_scale = engine.nclasses, 1 if engine.log_scale else 0
            arr, out, keep = engine.build(self.coords, self.feats4, self.cm, decode=dec)
            self.arr, self.out_full, self.keep = arr, (None if fuse_decode else out), keep
            self.grids = None
            if vote:
                X, Y, Z = (int(d) for d in vote["dims"])
                self.grids = (torch.empty((X, Y, Z), **f32), torch.empty((X, Y, Z, 2), **f32), torch.empty((X, Y, Z, 3), **f32))
            self.stream = torch.cuda.Stream(dev)
            # The map builder runs on a high-priority branch of the graph: with several scenes in flight its small kernels are
            # then dispatched beside the persistent convolution CTAs of another scene (they need no shared memory and few
            # registers) instead of queueing behind that scene's pending, programmatically launched convolutions.
            self.side = torch.cuda.Stream(dev, priority=-1)

            def body():
                cur = torch.cuda.current_stream()
                if pad4:
                    self.feats4[:, :self.feats_in.shape[1]].copy_(self.feats_in)
                self.side.wait_stream(cur)
                with torch.cuda.stream(self.side):
                    self.cm.enqueue()
                cur.wait_stream(self.side)
                _lib.check(L.cvb200_sc_run_program(arr, len(arr), _stream()), "cvb200_sc_run_program")
                if not fuse_decode:
                    rc = L.cvb200_head_decode_points(_ptr(out), out.stride(0), n, engine.nclasses, 1 if engine.log_scale else 0, _ptr(self.xyz),
                                                     _ptr(self.scale), _ptr(self.cls), _ptr(self.prob), _ptr(self.coords),
                                                     ctypes.c_float(self.res), _ptr(self.points), _stream())
                    _lib.check(rc, "cvb200_head_decode_points")
                if vote:
                    work = H._workspace(L, vote["dims"], dev)
                    rc = L.cvb200_hv_forward(_ptr(self.points), _ptr(self.xyz), _ptr(self.scale), _ptr(self.prob), n, self.res,
                                             int(vote["num_rots"]), _lib.f3(vote["corner"]), _lib.i3(vote["dims"]), _ptr(self.grids[0]),
                                             _ptr(self.grids[1]), _ptr(self.grids[2]), _ptr(work), work.numel(), _stream())
                    _lib.check(rc, "cvb200_hv_forward")

            # warm-up on the capture stream (per-stream scratch allocations, function attributes), then capture
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                body()
            torch.cuda.synchronize(dev)
            self.exec, self.pdl, self.nodes = None, True, 0
            for use_pdl in (1, 0):
                # a driver that cannot capture programmatic dependent launches gets plain stream-ordered launches
                L.cvb200_sc_set_conv_options(CONV_OPTIONS[0], (CONV_OPTIONS[1] & ~1) | use_pdl)
                try:
                    with torch.cuda.stream(self.stream):
                        sp = ctypes.c_void_p(self.stream.cuda_stream)
                        _lib.check(L.cvb200_graph_begin(sp), "cvb200_graph_begin")
                        try:
                            body()
                        except Exception:
                            L.cvb200_graph_abort(sp)
                            raise
                        ex, nn = ctypes.c_void_p(), ctypes.c_int64()
                        prio = 0 if os.environ.get("CVB200_GRAPH_NO_PRIORITY") == "1" else 1      # A/B switch for measurements
                        _lib.check(L.cvb200_graph_end(sp, prio, ctypes.byref(ex), ctypes.byref(nn)), "cvb200_graph_end")
                    self.exec, self.pdl, self.nodes = ex, bool(use_pdl), int(nn.value)
                    break
                except Exception:
                    if not use_pdl:
                        raise
                    torch.cuda.synchronize(dev)
                finally:
                    L.cvb200_sc_set_conv_options(*CONV_OPTIONS)
        self.launches = len(arr) + 31 + (0 if fuse_decode else 1) + (1 if pad4 else 0) + (2 if vote else 0)

    def run(self, coords, feats):
        """Copy one scene's inputs (int32 [n,4] coordinates, float32 [n,C] features; pinned host or device tensors) into the
        lane's buffers and replay the graph on the current stream.  Returns the lane's output tensors (see class docstring)."""
        if coords.shape[0] != self.n:
            raise RuntimeError("SceneGraph built for %d voxels, got %d" % (self.n, coords.shape[0]))
        self.coords.copy_(coords, non_blocking=True)
        self.feats_in.copy_(feats, non_blocking=True)
        _lib.check(_lib.load().cvb200_graph_launch(self.exec, _stream()), "cvb200_graph_launch")
        return {"feats": self.out_full, "xyz": self.xyz, "scale": self.scale, "class_pred": self.cls, "prob": self.prob,
                "points": self.points, "grids": self.grids}

    def __del__(self):
        try:
            if getattr(self, "exec", None):
                _lib.load().cvb200_graph_destroy(self.exec)
        except Exception:
            pass

    def level_counts(self):
        """Voxels per level of the lane's last scene (synchronises; diagnostics / tests)."""
        return [int(v) for v in self.cm.counts.cpu().tolist()]
