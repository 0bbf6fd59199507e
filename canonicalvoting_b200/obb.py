"""Oriented-box IoU and per-class NMS on the device -- the detection post-process that follows the candidate loop in the
reference's eval scripts (eval_joint.py:75-89 `nms`, :265-280; utils/calc_map.py:6-21 `get_iou_obb`, which needs shapely).

    keep = nms_per_class(boxes, scores, classes, nclasses, 0.3)      # int64 indices, class by class, pick order
    iou = iou_matrix(boxes_a, boxes_b)                               # float64 [na, nb]

Inputs are CUDA tensors (boxes float32 [K,8,3] as back_project returns them); there is no CPU fallback."""
import torch

from . import _lib
from .hv_cuda import _ptr, _stream_ptr


def nms_per_class(boxes, scores, classes, nclasses, overlap_threshold=0.3):
    if not (boxes.is_cuda and scores.is_cuda and classes.is_cuda):
        raise RuntimeError("nms_per_class: CUDA tensors expected (there is no CPU path)")
    L = _lib.load()
    k = int(boxes.shape[0])
    boxes = boxes.to(torch.float32).contiguous().view(k, 24)
    scores = scores.to(torch.float32).contiguous()
    cls = classes.to(torch.int32).contiguous()
    pick = torch.empty(max(k, 1), dtype=torch.int32, device=boxes.device)
    n_pick = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        rc = L.cvb200_obb_nms(_ptr(boxes), _ptr(scores), _ptr(cls), k, int(nclasses), float(overlap_threshold), _ptr(pick), _ptr(n_pick),
                              _stream_ptr())
        _lib.check(rc, "cvb200_obb_nms")
    return pick[:int(n_pick.item())].long()


def iou_matrix(a, b):
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("iou_matrix: CUDA tensors expected (there is no CPU path)")
    L = _lib.load()
    na, nb = int(a.shape[0]), int(b.shape[0])
    a = a.to(torch.float32).contiguous().view(na, 24)
    b = b.to(torch.float32).contiguous().view(nb, 24)
    out = torch.empty((na, nb), dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(L.cvb200_obb_iou_matrix(_ptr(a), na, _ptr(b), nb, _ptr(out), _stream_ptr()), "cvb200_obb_iou_matrix")
    return out


def get_iou_obb(bbox1, bbox2):
    """Drop-in for utils/calc_map.get_iou_obb (numpy [8,3] corner arrays -> float) without shapely: evaluated on the device.
    For many pairs use iou_matrix."""
    import numpy as np
    a = torch.from_numpy(np.ascontiguousarray(bbox1, dtype=np.float32)).view(1, 8, 3).cuda()
    b = torch.from_numpy(np.ascontiguousarray(bbox2, dtype=np.float32)).view(1, 8, 3).cuda()
    return float(iou_matrix(a, b)[0, 0])
