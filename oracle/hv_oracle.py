"""oracle/hv_oracle.py -- TEST INFRASTRUCTURE ONLY.

numpy/ctypes front-end of oracle/hv_oracle.c, the CPU restatement of the reference
`hv_cuda.forward/backward` (houghvoting/src/hv_cuda_kernel.cu:12-302).  Signatures
mirror the reference op; arrays are float32 numpy.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhv_oracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int32)


def build():
    src = os.path.join(_HERE, "hv_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libhv_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.cvo_num_threads.restype = ctypes.c_int
        L.cvo_theta_table.argtypes = [ctypes.c_int, _f, _f]
        L.cvo_grid_dims.argtypes = [_f, ctypes.c_int64, ctypes.c_float, _f, _i]
        L.cvo_grid_dims.restype = ctypes.c_int
        L.cvo_vote_forward.argtypes = [_f, _f, _f, _f, ctypes.c_int64, ctypes.c_float, ctypes.c_int, _f, _f,
                                       _f, _i, _f, _f, _f, _i, ctypes.c_int, ctypes.c_int]
        L.cvo_vote_forward.restype = ctypes.c_int64
        L.cvo_average.argtypes = [ctypes.c_int64, _f, _f, _f, ctypes.c_int]
        L.cvo_vote_backward.argtypes = [_f, _f, _f, _f, _f, ctypes.c_int64, ctypes.c_float, ctypes.c_int, _f, _f,
                                        _f, _i, _f, _f, _f, ctypes.c_int]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_i) if a is not None else None


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return lib().cvo_num_threads()


def theta_table(num_rots):
    """glibc cosf/sinf of theta_i = i * (2*3.141592654f / R) (hv_cuda_kernel.cu:35-37)."""
    c = np.empty(num_rots, np.float32)
    s = np.empty(num_rots, np.float32)
    lib().cvo_theta_table(int(num_rots), _fp(c), _fp(s))
    return c, s


def grid_dims(points, res):
    """corner = min(points,0); dims = int((max-min)/res)+1 in float32 (hv_cuda_kernel.cu:129-134)."""
    points = _c(points)
    corner = np.empty(3, np.float32)
    dims = np.empty(3, np.int32)
    rc = lib().cvo_grid_dims(_fp(points), points.shape[0], np.float32(res), _fp(corner), _ip(dims))
    if rc != 0:
        raise ValueError("grid_dims: empty point set")
    return corner, dims


def forward(points, xyz, scale, obj, res, num_rots, theta=None, corner=None, dims=None,
            acc64=True, threads=1, average=True, return_votes=False):
    """Restatement of hv_cuda.forward (hv_cuda_kernel.cu:121-165).

    Returns (grid_obj [X,Y,Z], grid_rot [X,Y,Z,2], grid_scale [X,Y,Z,3]) and, when
    return_votes, the int32 [N,R,3] floor voxel of every vote (-1 = dropped).
    """
    points, xyz, scale, obj = _c(points), _c(xyz), _c(scale), _c(obj)
    n = points.shape[0]
    if corner is None or dims is None:
        corner, dims = grid_dims(points, res)
    corner = _c(corner)
    dims = np.ascontiguousarray(dims, dtype=np.int32)
    X, Y, Z = (int(d) for d in dims)
    g_obj = np.zeros((X, Y, Z), np.float32)
    g_rot = np.zeros((X, Y, Z, 2), np.float32)
    g_scale = np.zeros((X, Y, Z, 3), np.float32)
    ct = st = None
    if theta is not None:
        ct, st = _c(theta[0]), _c(theta[1])
    votes = np.empty((n, num_rots, 3), np.int32) if return_votes else None
    lib().cvo_vote_forward(_fp(points), _fp(xyz), _fp(scale), _fp(obj), n, np.float32(res), int(num_rots),
                           _fp(ct), _fp(st), _fp(corner), _ip(dims), _fp(g_obj), _fp(g_rot), _fp(g_scale),
                           _ip(votes), int(bool(acc64)), int(threads))
    if average:
        lib().cvo_average(X * Y * Z, _fp(g_obj), _fp(g_rot), _fp(g_scale), int(threads))
    if return_votes:
        return g_obj, g_rot, g_scale, votes
    return g_obj, g_rot, g_scale


def backward(grad_grid, points, xyz, scale, obj, res, num_rots, theta=None, corner=None, threads=1):
    """Restatement of hv_cuda.backward (hv_cuda_kernel.cu:265-302): (d_xyz, d_scale, d_obj)."""
    grad_grid, points, xyz, scale, obj = _c(grad_grid), _c(points), _c(xyz), _c(scale), _c(obj)
    n = points.shape[0]
    if corner is None:
        corner, _ = grid_dims(points, res)
    corner = _c(corner)
    dims = np.asarray(grad_grid.shape, dtype=np.int32)
    ct = st = None
    if theta is not None:
        ct, st = _c(theta[0]), _c(theta[1])
    d_xyz = np.empty((n, 3), np.float32)
    d_scale = np.empty((n, 3), np.float32)
    d_obj = np.empty((n,), np.float32)
    lib().cvo_vote_backward(_fp(grad_grid), _fp(points), _fp(xyz), _fp(scale), _fp(obj), n, np.float32(res),
                            int(num_rots), _fp(ct), _fp(st), _fp(corner), _ip(dims),
                            _fp(d_xyz), _fp(d_scale), _fp(d_obj), int(threads))
    return d_xyz, d_scale, d_obj
