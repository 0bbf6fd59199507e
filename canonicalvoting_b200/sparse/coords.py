"""Coordinate manager of the sparse-voxel U-Net (host side of csrc/sparse_coords.cu).

Replaces MinkowskiEngine's CoordinateManager as the reference uses it implicitly: one manager per
`ME.SparseTensor(feats, coords, device='cuda')` (train_joint.py:250, eval_joint.py:169), coordinate
sets per tensor stride, kernel maps cached per (stride, kernel) and re-used by every layer of the
level, and the encoder's coordinate sets re-used by the transposed convolutions of the decoder
(utils/minkunet.py:85-106, which is what makes `ME.cat` with the skip tensors valid).

All tables live on the device; the only host synchronisation is ONE scalar read per stride-2
down-sampling (the number of coarse voxels sizes the next level's tensors).
"""
import ctypes

import torch

from .. import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    """Raw handle of torch's current stream on the current device (the private getter is ~40x cheaper than building a
    torch.cuda.Stream object, and this is called for every launch)."""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


class Level:
    """Coordinate set of one tensor stride: rows (b,x,y,z) int32 + hash table key -> row."""

    def __init__(self, coords, tensor_stride):
        L = _lib.load()
        self.coords = coords                      # int32 [n,4], device, contiguous
        self.n = int(coords.shape[0])
        self.tensor_stride = int(tensor_stride)
        self.capacity = int(L.cvb200_sc_hash_capacity(self.n))
        dev = coords.device
        self.keys = torch.empty(self.capacity, dtype=torch.int64, device=dev)
        self.vals = torch.empty(self.capacity, dtype=torch.int32, device=dev)
        self.table_built = False

    def build_table(self):
        if not self.table_built:
            L = _lib.load()
            rc = L.cvb200_sc_build_table(_ptr(self.coords), self.n, _ptr(self.keys), _ptr(self.vals), self.capacity, _stream())
            _lib.check(rc, "cvb200_sc_build_table")
            self.table_built = True


class CoordinateManager:
    def __init__(self, coords):
        if coords.dtype != torch.int32 or coords.dim() != 2 or coords.shape[1] != 4:
            raise ValueError("coordinates must be int32 [N,4] rows (batch, x, y, z)")
        if not coords.is_cuda:
            raise RuntimeError("canonicalvoting_b200.sparse needs CUDA coordinates (there is no CPU path)")
        self.device = coords.device
        self._validate(coords)
        self.levels = {1: Level(coords.contiguous(), 1)}
        self._nbr = {}      # (tensor_stride, ksize) -> [n, ksize^3] int32
        self._down = {}     # fine tensor_stride -> dict(children, up_table, parent, koff)
        # The first map a layer asks for triggers ONE fused build of everything a U-Net needs (four stride-2 levels, their 3^3
        # maps, children / parent tables) with one host synchronisation, instead of a read-back per level; whatever a network
        # asks beyond that is still built step by step on the same hash tables (identical numbering, tests/test_sparse_gpu.py).
        self._lazy_unet = True

    @staticmethod
    def _validate(coords):
        """The hash keys pack (batch, x, y, z) into 16 bits per field (csrc/sparse_hash.cuh): coordinates outside that range --
        or so close to it that a neighbour probe c +- 2 * 16 (5^3 kernel, tensor stride 16) would leave it -- would silently
        alias other voxels.  One min/max reduction per manager (a size-agnostic SceneGraph validates when its inputs change
        shape, not per replay: its coordinates are the caller's responsibility)."""
        if coords.shape[0] == 0 or getattr(CoordinateManager, "_skip_validation", False):
            return
        lo, hi = torch.aminmax(coords, dim=0)
        lo, hi = lo.tolist(), hi.tolist()
        margin = 64
        if lo[0] < 0 or hi[0] >= 65535 or min(lo[1:]) < -32768 + margin or max(hi[1:]) > 32767 - margin:
            raise ValueError("voxel coordinates out of range: batch index in [0, 65534], x / y / z in [%d, %d] "
                             "(got batch %d..%d, xyz %d..%d)" % (-32768 + margin, 32767 - margin, lo[0], hi[0], min(lo[1:]), max(hi[1:])))

    # ------------------------------------------------------------------ everything a U-Net needs, one enqueue + one sync
    @classmethod
    def build_unet(cls, coords, stem_ksize, n_down=4, pinned_counts=None):
        """Levels 1, 2, ..., 2^n_down with their 3^3 maps, the stem map of level 1, children / parent tables of every
        stride-2 step and the identity tables, built by ONE call of cvb200_sc_build_maps (csrc/sparse_maps.cu) on the
        current stream and ONE synchronisation of it -- instead of a host read-back per level.  The result behaves like a
        manager on which kernel_map() / down() were already called (same tables, same numbering)."""
        cm = cls(coords)
        cm._lazy_unet = False
        cm._fill_unet(stem_ksize, n_down, pinned_counts)
        return cm

    def _fill_unet(self, stem_ksize, n_down=4, pinned_counts=None):
        cm = self
        L = _lib.load()
        n = cm.levels[1].n
        lay = _lib.MapsLayout()
        _lib.check(L.cvb200_sc_maps_layout(n, stem_ksize, n_down, ctypes.byref(lay)), "cvb200_sc_maps_layout")
        dev = cm.device
        with torch.cuda.device(dev):
            ws = torch.empty(lay.total_bytes, dtype=torch.uint8, device=dev)
            if pinned_counts is None:
                pinned_counts = torch.empty(8, dtype=torch.int32).pin_memory()
            stream = torch.cuda.current_stream()
            rc = L.cvb200_sc_build_maps(_ptr(cm.levels[1].coords), n, stem_ksize, n_down, _ptr(ws), ctypes.byref(lay),
                                        ctypes.c_void_p(pinned_counts.data_ptr()), _stream())
            _lib.check(rc, "cvb200_sc_build_maps")
            stream.synchronize()                                   # the one host synchronisation of the scene
        counts = [int(v) for v in pinned_counts[:n_down + 1].tolist()]

        def view(off, rows, cols, dtype=torch.int32):
            nbytes = rows * max(cols, 1) * (8 if dtype == torch.int64 else 4)        # cols == 0: a vector of `rows` elements
            t = ws[off:off + nbytes].view(dtype)
            return t.view(rows, cols) if cols else t

        cm._ws = ws
        arange = view(lay.arange, n, 1)
        for l in range(n_down + 1):
            ts = 1 << l
            if l:
                lv = Level.__new__(Level)
                lv.coords, lv.n, lv.tensor_stride = view(lay.coords[l], counts[l], 4), counts[l], ts
                cm.levels[ts] = lv
            lv = cm.levels[ts]
            lv.capacity = int(lay.capacity)
            lv.keys = view(lay.keys[l], lv.capacity, 0, torch.int64)
            lv.vals = view(lay.vals[l], lv.capacity, 0)
            lv.table_built = True
            cm._nbr[(ts, 3)] = view(lay.nbr3[l], counts[l], 27)
            cm._nbr[("ident", ts)] = arange[:counts[l]]
        if stem_ksize:
            cm._nbr[(1, stem_ksize)] = view(lay.stem_table, n, stem_ksize ** 3)
        for l in range(n_down):
            cm._down[1 << l] = dict(children=view(lay.children[l], counts[l + 1], 8), up_table=view(lay.up_table[l], counts[l], 8),
                                    parent=view(lay.parent[l], counts[l], 0), koff=view(lay.koff[l], counts[l], 0))

    def _maybe_fill_unet(self, tensor_stride, ksize):
        if self._lazy_unet:
            self._lazy_unet = False
            if self.levels[1].n >= 64 and not self._nbr and not self._down:
                stem = ksize if (tensor_stride == 1 and ksize % 2 == 1 and 3 < ksize <= 7) else 0
                self._fill_unet(stem, 4)

    @classmethod
    def static_unet(cls, coords, stem_ksize, n_down=4):
        """Size-agnostic variant of build_unet for CUDA-graph capture: every level and table is a view of `n` rows (the upper
        bound: a coarse level never has more voxels than the input), the real sizes stay in the device array `counts` that the
        kernels read (cvb200_sc_op.n_out_dev) -- no host synchronisation, nothing depends on a size the host would have to
        know.  `coords` is the caller's persistent int32 [n,4] input buffer; `enqueue()` (re)builds everything for its current
        contents on the current stream.  `count_ptr(level)` is the device address of level 2^level's row count."""
        cm = cls(coords)
        L = _lib.load()
        n = cm.levels[1].n
        lay = _lib.MapsLayout()
        _lib.check(L.cvb200_sc_maps_layout(n, stem_ksize, n_down, ctypes.byref(lay)), "cvb200_sc_maps_layout")
        dev = cm.device
        with torch.cuda.device(dev):
            ws = torch.empty(lay.total_bytes, dtype=torch.uint8, device=dev)
        pinned = torch.empty(8, dtype=torch.int32).pin_memory()

        def view(off, rows, cols, dtype=torch.int32):
            nbytes = rows * max(cols, 1) * (8 if dtype == torch.int64 else 4)
            t = ws[off:off + nbytes].view(dtype)
            return t.view(rows, cols) if cols else t

        cm._ws, cm._pinned_counts, cm.static = ws, pinned, True
        arange = view(lay.arange, n, 1)
        for l in range(n_down + 1):
            ts = 1 << l
            if l:
                lv = Level.__new__(Level)
                lv.coords, lv.n, lv.tensor_stride = view(lay.coords[l], n, 4), n, ts
                cm.levels[ts] = lv
            lv = cm.levels[ts]
            lv.capacity = int(lay.capacity)
            lv.keys = view(lay.keys[l], lv.capacity, 0, torch.int64)
            lv.vals = view(lay.vals[l], lv.capacity, 0)
            lv.table_built = True
            cm._nbr[(ts, 3)] = view(lay.nbr3[l], n, 27)
            cm._nbr[("ident", ts)] = arange
        if stem_ksize:
            cm._nbr[(1, stem_ksize)] = view(lay.stem_table, n, stem_ksize ** 3)
        for l in range(n_down):
            cm._down[1 << l] = dict(children=view(lay.children[l], n, 8), up_table=view(lay.up_table[l], n, 8),
                                    parent=view(lay.parent[l], n, 0), koff=view(lay.koff[l], n, 0))
        counts_base = ws.data_ptr() + lay.counts
        cm.count_ptr = lambda level: counts_base + 4 * level
        cm.counts = view(lay.counts, n_down + 1, 0)

        def enqueue():
            rc = L.cvb200_sc_build_maps(_ptr(cm.levels[1].coords), n, stem_ksize, n_down, _ptr(ws), ctypes.byref(lay),
                                        ctypes.c_void_p(pinned.data_ptr()), _stream())
            _lib.check(rc, "cvb200_sc_build_maps")
        cm.enqueue = enqueue
        return cm

    def identity_table(self, tensor_stride):
        """[n, 1] table row -> row: a 1x1x1 convolution as the one primitive."""
        key = ("ident", tensor_stride)
        t = self._nbr.get(key)
        if t is None:
            t = torch.arange(self.levels[tensor_stride].n, dtype=torch.int32, device=self.device).view(-1, 1).contiguous()
            self._nbr[key] = t
        return t

    # ------------------------------------------------------------------ stride-1 kernel maps
    def kernel_map(self, tensor_stride, ksize):
        self._maybe_fill_unet(tensor_stride, ksize)
        key = (tensor_stride, ksize)
        nbr = self._nbr.get(key)
        if nbr is None:
            L = _lib.load()
            lv = self.levels[tensor_stride]
            with torch.cuda.device(self.device):
                lv.build_table()
                nbr = torch.empty((lv.n, ksize ** 3), dtype=torch.int32, device=self.device)
                rc = L.cvb200_sc_kernel_map(_ptr(lv.coords), lv.n, _ptr(lv.keys), _ptr(lv.vals), lv.capacity, ksize,
                                            tensor_stride, _ptr(nbr), _stream())
                _lib.check(rc, "cvb200_sc_kernel_map")
            self._nbr[key] = nbr
        return nbr

    # ------------------------------------------------------------------ stride-2 down / up maps
    def down(self, tensor_stride):
        """Maps between the level `tensor_stride` (fine) and 2*tensor_stride (coarse); creates the coarse level."""
        self._maybe_fill_unet(tensor_stride, 0)
        d = self._down.get(tensor_stride)
        if d is None:
            L = _lib.load()
            fine = self.levels[tensor_stride]
            new_stride = 2 * tensor_stride
            with torch.cuda.device(self.device):
                keys = torch.empty(fine.capacity, dtype=torch.int64, device=self.device)
                vals = torch.empty(fine.capacity, dtype=torch.int32, device=self.device)
                flag = torch.empty(fine.n, dtype=torch.int32, device=self.device)
                rc = L.cvb200_sc_down_flags(_ptr(fine.coords), fine.n, new_stride, _ptr(keys), _ptr(vals), fine.capacity,
                                            _ptr(flag), _stream())
                _lib.check(rc, "cvb200_sc_down_flags")
                incl = torch.cumsum(flag, 0, dtype=torch.int32)
                n_coarse = int(incl[-1].item())                      # the one host sync of this level
                excl = (incl - flag).contiguous()
                coords_c = torch.empty((n_coarse, 4), dtype=torch.int32, device=self.device)
                parent = torch.empty(fine.n, dtype=torch.int32, device=self.device)
                koff = torch.empty(fine.n, dtype=torch.int32, device=self.device)
                children = torch.empty((n_coarse, 8), dtype=torch.int32, device=self.device)
                up_table = torch.empty((fine.n, 8), dtype=torch.int32, device=self.device)
                rc = L.cvb200_sc_down_finish(_ptr(fine.coords), fine.n, new_stride, _ptr(keys), _ptr(vals), fine.capacity,
                                             _ptr(flag), _ptr(excl), n_coarse, _ptr(coords_c), _ptr(parent), _ptr(koff),
                                             _ptr(children), _ptr(up_table), _stream())
                _lib.check(rc, "cvb200_sc_down_finish")
            coarse = Level(coords_c, new_stride)
            if fine.capacity == coarse.capacity:
                # the coarse-key table built above already maps coarse coordinates -> coarse rows
                coarse.keys, coarse.vals, coarse.table_built = keys, vals, True
            self.levels[new_stride] = coarse
            d = dict(children=children, up_table=up_table, parent=parent, koff=koff)
            self._down[tensor_stride] = d
        return d
