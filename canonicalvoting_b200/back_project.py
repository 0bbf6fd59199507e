"""back_project -- the reference's candidate loop with the LCC-aware back-projection check
(inline script code, eval_joint.py:195-263 / train_joint.py:355-424) as ONE device-resident loop
(csrc/bp_loop.cu) with no per-iteration host synchronisation.

    boxes, scores, classes = back_project(grid_obj, grid_rot, grid_scale, scan_points, xyz_pred,
                                          prob_pred, class_pred, res, thresh_high=60, ...)

Arguments are the tensors the script holds at eval_joint.py:192-203; `grid_obj` is zeroed in place
exactly like the script does.  Returns torch tensors on the device (boxes [K,8,3] f32, scores [K]
f32, classes [K] i64); `back_project_numpy` returns the `boxes / scores / probs / classes` numpy
arrays the script builds at :265-268.  There is no CPU fallback.
"""
import ctypes

import torch

from . import _lib
from .hv_cuda import _check_input, _ptr, _stream_ptr, grid_dims

_work = {}   # (device, stream) -> scratch tensor


def default_params(**overrides):
    L = _lib.load()
    p = _lib.BpParams()
    L.cvb200_bp_default_params(ctypes.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise TypeError("back_project: unknown parameter %r" % k)
        setattr(p, k, v)
    return p


def back_project(grid_obj, grid_rot, grid_scale, scan_points, xyz_pred, prob_pred, class_pred, res,
                 corner=None, return_trace=False, **params):
    """See module docstring.  `corner` (3 floats) defaults to min(scan_points, 0) like the script
    (:201); pass the value hv_cuda.forward used to avoid the extra reduction + sync."""
    for t, name in ((grid_obj, "grid_obj"), (grid_rot, "grid_rot"), (grid_scale, "grid_scale"),
                    (scan_points, "scan_points"), (xyz_pred, "xyz_pred"), (prob_pred, "prob_pred"),
                    (class_pred, "class_pred")):
        _check_input(t, name)
    if grid_obj.dtype != torch.float32 or scan_points.dtype != torch.float32:
        raise RuntimeError("back_project: float32 tensors expected")
    if class_pred.dtype != torch.int64:
        raise RuntimeError("class_pred must be int64 (torch.argmax output)")
    L = _lib.load()
    dev = grid_obj.device
    n = scan_points.shape[0]
    with torch.cuda.device(dev):
        if corner is None:
            corner, _, _ = grid_dims(scan_points, float(res))
        p = default_params(**params)
        if return_trace and p.max_trace == 0:
            p.max_trace = 1 << 16
        dims = _lib.i3(grid_obj.shape)
        need = L.cvb200_bp_work_bytes(dims)
        key = (dev.index, torch.cuda.current_stream().cuda_stream)
        w = _work.get(key)
        if w is None or w.numel() < need:
            w = torch.empty(need, dtype=torch.uint8, device=dev)
            _work[key] = w
        boxes = torch.empty((p.max_boxes, 8, 3), dtype=torch.float32, device=dev)
        scores = torch.empty((p.max_boxes,), dtype=torch.float32, device=dev)
        classes = torch.empty((p.max_boxes,), dtype=torch.int32, device=dev)
        counts = torch.zeros((2,), dtype=torch.int32, device=dev)
        trace = torch.zeros((max(p.max_trace, 1), 4), dtype=torch.int32, device=dev)
        rc = L.cvb200_back_project(_ptr(grid_obj), _ptr(grid_rot), _ptr(grid_scale), dims, _lib.f3(corner), float(res),
                                   _ptr(scan_points), _ptr(xyz_pred), _ptr(prob_pred), _ptr(class_pred), n,
                                   ctypes.byref(p), _ptr(boxes), _ptr(scores), _ptr(classes), _ptr(counts),
                                   _ptr(trace) if p.max_trace else None, _ptr(w), w.numel(), _stream_ptr())
        _lib.check(rc, "cvb200_back_project")
        k, iters = counts.tolist()          # the ONE host synchronisation of the whole loop
    out = (boxes[:k], scores[:k], classes[:k].long())
    if return_trace:
        return out + (trace[:min(iters, p.max_trace)], iters)
    return out


def back_project_numpy(*args, **kw):
    """(boxes, scores, probs, classes) numpy arrays as the script builds them (eval_joint.py:265-268;
    scores and probs are the same values there, :261-262)."""
    boxes, scores, classes = back_project(*args, **kw)
    s = scores.cpu().numpy()
    return boxes.cpu().numpy(), s, s.copy(), classes.cpu().numpy()
