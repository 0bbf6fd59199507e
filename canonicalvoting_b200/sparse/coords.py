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
        self.levels = {1: Level(coords.contiguous(), 1)}
        self._nbr = {}      # (tensor_stride, ksize) -> [n, ksize^3] int32
        self._down = {}     # fine tensor_stride -> dict(children, up_table, parent, koff)

    # ------------------------------------------------------------------ stride-1 kernel maps
    def kernel_map(self, tensor_stride, ksize):
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
