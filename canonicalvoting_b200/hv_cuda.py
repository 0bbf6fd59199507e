"""Drop-in for the reference's native module `hv_cuda` (houghvoting/src/hv_cuda.cpp:74-77).

    forward(points, xyz_labels, scale_labels, obj_labels, res, num_rots[, corners]) -> [grid_obj, grid_rot, grid_scale]
    backward(grad_grid, points, xyz_labels, scale_labels, obj_labels, res, num_rots) -> [d_xyz, d_scale, d_obj]

Same positional signatures, return structure and error behaviour as the reference
(RuntimeError "<name> must be a CUDA tensor" / "<name> must be contiguous",
hv_cuda.cpp:26-28).  Differences, all supersets:
  * runs on the tensors' device and torch's CURRENT stream (the reference always uses
    the legacy default stream of device 0, hv_cuda_kernel.cu:143);
  * ONE host synchronisation per forward (grid geometry) instead of twelve
    (hv_cuda_kernel.cu:132-134,151);
  * optional 7th argument `corners` [2,3] overriding min/max of the points, the call
    form sunrgbd/brnetcanon.py:99 uses;
  * float64 inputs are computed in float32 and cast back (the reference's float64
    instantiation also does its geometry in float32, hv_cuda_kernel.cu:29-40).
"""
import ctypes

import torch

from . import _lib

_work_cache = {}       # (device index, stream id) -> zero-filled workspace tensor
_dims_work = {}        # (device index, stream id) -> small scratch for the min/max reduction


def _check_input(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _host_scalar(t):
    """Value of a 0-dim device tensor (res / num_rots).  The value is remembered ON THE TENSOR OBJECT (together
    with its version counter), so the constant tensors HoughVoting holds (train_joint.py:52-53) cost one
    device->host sync ever, not one per call -- and a new tensor that happens to reuse the address is never
    mistaken for an old one."""
    tag = getattr(t, "_cvb200_host_value", None)
    if tag is not None and tag[0] == t._version:
        return tag[1]
    v = t.item()
    try:
        t._cvb200_host_value = (t._version, v)
    except Exception:
        pass
    return v


def _stream_ptr():
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def grid_dims(points, res):
    """(corner[3], maxpt[3], dims[3]) python tuples; the reference's float32 host arithmetic
    (hv_cuda_kernel.cu:129-134).  Synchronises the current stream once."""
    L = _lib.load()
    dev = (points.device.index, torch.cuda.current_stream(points.device).cuda_stream)   # per stream: callers may run concurrently
    work = _dims_work.get(dev)
    if work is None:
        work = torch.empty(L.cvb200_hv_grid_dims_work_bytes(), dtype=torch.uint8, device=points.device)
        _dims_work[dev] = work
    corner = (ctypes.c_float * 3)()
    maxpt = (ctypes.c_float * 3)()
    dims = (ctypes.c_int32 * 3)()
    rc = L.cvb200_hv_grid_dims(_ptr(points), points.shape[0], float(res), _ptr(work), corner, maxpt, dims,
                               _stream_ptr())
    _lib.check(rc, "cvb200_hv_grid_dims")
    return tuple(corner), tuple(maxpt), tuple(dims)


def _workspace(L, dims, device):
    """Zero-filled ONCE per (device, stream); cvb200_hv_forward leaves it all-zero again
    (include/cvb200.h contract), so a larger workspace is reused for smaller grids."""
    need = L.cvb200_hv_forward_work_bytes(_lib.i3(dims))
    key = (device.index, torch._C._cuda_getCurrentRawStream(device.index))
    w = _work_cache.get(key)
    if w is None or w.numel() < need:
        w = torch.zeros(need, dtype=torch.uint8, device=device)
        _work_cache[key] = w
    return w


def forward_host(points, xyz, scale, obj, res, num_rots, corner, dims):
    """Asynchronous core: host-side scalars/geometry already known -> no sync at all."""
    L = _lib.load()
    X, Y, Z = (int(d) for d in dims)
    opts = dict(dtype=torch.float32, device=points.device)
    grid_obj = torch.empty((X, Y, Z), **opts)
    grid_rot = torch.empty((X, Y, Z, 2), **opts)
    grid_scale = torch.empty((X, Y, Z, 3), **opts)
    work = _workspace(L, dims, points.device)
    rc = L.cvb200_hv_forward(_ptr(points), _ptr(xyz), _ptr(scale), _ptr(obj), points.shape[0], float(res),
                             int(num_rots), _lib.f3(corner), _lib.i3(dims), _ptr(grid_obj), _ptr(grid_rot),
                             _ptr(grid_scale), _ptr(work), work.numel(), _stream_ptr())
    if rc != 0:
        _work_cache.clear()  # the all-zero contract may be broken
    _lib.check(rc, "cvb200_hv_forward")
    return grid_obj, grid_rot, grid_scale


def forward(points, xyz_labels, scale_labels, obj_labels, res, num_rots, corners=None):
    """hv_forward (hv_cuda.cpp:30-45) -> hv_cuda_forward (hv_cuda_kernel.cu:121-165)."""
    for t, name in ((points, "points"), (xyz_labels, "xyz_labels"), (scale_labels, "scale_labels"),
                    (obj_labels, "obj_labels"), (res, "res"), (num_rots, "num_rots")):
        _check_input(t, name)
    out_dtype = points.dtype
    with torch.cuda.device(points.device):
        res_h = float(_host_scalar(res))
        rots_h = int(_host_scalar(num_rots))
        p, x, s, o = _f32(points), _f32(xyz_labels), _f32(scale_labels), _f32(obj_labels)
        if corners is not None:
            _check_input(corners, "corners")
            c = corners.detach().float().cpu()
            res32 = torch.tensor(res_h, dtype=torch.float32)
            corner = tuple(float(v) for v in c[0])
            dims = tuple(int(v) + 1 for v in ((c[1] - c[0]) / res32).to(torch.int32))
        else:
            corner, _, dims = grid_dims(p, res_h)
        outs = forward_host(p, x, s, o, res_h, rots_h, corner, dims)
    if out_dtype != torch.float32:
        outs = tuple(t.to(out_dtype) for t in outs)
    return list(outs)


def backward_host(grad_grid, points, xyz, scale, obj, res, num_rots, corner):
    L = _lib.load()
    d_xyz = torch.empty_like(xyz)
    d_scale = torch.empty_like(scale)
    d_obj = torch.empty_like(obj)
    rc = L.cvb200_hv_backward(_ptr(grad_grid), _ptr(points), _ptr(xyz), _ptr(scale), _ptr(obj), points.shape[0],
                              float(res), int(num_rots), _lib.f3(corner), _lib.i3(grad_grid.shape), _ptr(d_xyz),
                              _ptr(d_scale), _ptr(d_obj), _stream_ptr())
    _lib.check(rc, "cvb200_hv_backward")
    return d_xyz, d_scale, d_obj


def backward(grad_grid, points, xyz_labels, scale_labels, obj_labels, res, num_rots):
    """hv_backward (hv_cuda.cpp:47-71) -> hv_cuda_backward (hv_cuda_kernel.cu:265-302)."""
    for t, name in ((grad_grid, "grad_grid"), (points, "points"), (xyz_labels, "xyz_labels"),
                    (scale_labels, "scale_labels"), (obj_labels, "obj_labels"), (res, "res"),
                    (num_rots, "num_rots")):
        _check_input(t, name)
    out_dtype = points.dtype
    with torch.cuda.device(points.device):
        res_h = float(_host_scalar(res))
        rots_h = int(_host_scalar(num_rots))
        p, x, s, o, g = _f32(points), _f32(xyz_labels), _f32(scale_labels), _f32(obj_labels), _f32(grad_grid)
        corner, _, _ = grid_dims(p, res_h)
        outs = backward_host(g, p, x, s, o, res_h, rots_h, corner)
    if out_dtype != torch.float32:
        outs = tuple(t.to(out_dtype) for t in outs)
    return list(outs)


def vote_indices(points, xyz, scale, res, num_rots, corner, dims):
    """int32 [N, num_rots, 3] floor voxel of every vote (-1 = dropped).  Verification aid."""
    L = _lib.load()
    out = torch.empty((points.shape[0], int(num_rots), 3), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        rc = L.cvb200_hv_vote_indices(_ptr(points), _ptr(xyz), _ptr(scale), points.shape[0], float(res),
                                      int(num_rots), _lib.f3(corner), _lib.i3(dims), _ptr(out), _stream_ptr())
    _lib.check(rc, "cvb200_hv_vote_indices")
    return out


def theta_table(num_rots, device="cuda"):
    """(cos, sin) float32 [num_rots] exactly as the device evaluates them."""
    L = _lib.load()
    device = torch.device(device)
    c = torch.empty(int(num_rots), dtype=torch.float32, device=device)
    s = torch.empty_like(c)
    with torch.cuda.device(device):
        rc = L.cvb200_hv_theta_table(int(num_rots), _ptr(c), _ptr(s), _stream_ptr())
    _lib.check(rc, "cvb200_hv_theta_table")
    return c, s
