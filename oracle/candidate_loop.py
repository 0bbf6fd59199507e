"""oracle/candidate_loop.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's candidate loop with the LCC-aware back-projection check,
which exists only as inline script code (eval_joint.py:195-263; same loop in
train_joint.py:355-424).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module; the product path (canonicalvoting_b200.back_project) never does.

Two restatements of the same lines:

  loop_torch()   the script's own torch op sequence on CPU tensors (argmax / slicing / matmul /
                 boolean-mask indexing / unique), i.e. what the reference executes, minus the
                 `.cuda()` calls.  The reference has no test or golden vector for this loop; both
                 restatements are pinned by golden vectors obtained from the reference script itself
                 (tools/make_ref_python_golden.py executes eval_joint.py:196-268 and
                 eval_separate.py:203-260 verbatim on CPU tensors -> tests/golden/refpy_loop_*.npz,
                 checked by tests/test_oracle_refpy.py), and the CUDA kernel is checked against them.
  loop_numpy()   float32 numpy with every arithmetic step spelled out in the order the CUDA
                 kernel uses (no library matmul, no FMA), so that integer decisions (which voxels
                 are zeroed, which points are inside, class ids) can be compared bit-exactly.

Float contract.  The script builds Rm = [[c,0,-s],[0,1,0],[s,0,c]] and evaluates
q = ((p - cand_world) @ Rm) / scale with a library matmul whose summation order is unspecified;
with the zeros of Rm the only rounding freedom is  q.x = dx*c + dz*s ,  q.z = -dx*s + dz*c
(two products, one add; an FMA would differ in the last bit).  loop_numpy() fixes
"round(product) + round(product)".  Inside/outside decisions can therefore differ from
loop_torch() only for points or voxels within 1 ulp of the box faces.

Thresholds are the script's module globals (eval_joint.py:18-21): thresh_high=60, thresh_low=10,
valid_ratio=0.2, elimination=2; the literals 0.3 (prob) and 0.3 (error) are at :245,:252.
"""
import numpy as np

DEFAULTS = dict(thresh_high=60.0, thresh_low=10, valid_ratio=0.2, elimination=2, prob_thresh=0.3,
                err_thresh=0.3, elim_hi_inclusive=True, max_boxes=4096)

# eval_joint.py:203 with l=h=w=2: rows of bbox_raw (8 corners, columns x,y,z)
BBOX_RAW = np.array([[1, 1, -1, -1, 1, 1, -1, -1],
                     [1, 1, 1, 1, -1, -1, -1, -1],
                     [1, -1, -1, 1, 1, -1, -1, 1]], dtype=np.float32).T


def _params(kw):
    p = dict(DEFAULTS)
    p.update(kw)
    return p


def loop_torch(grid_obj, grid_rot, grid_scale, points, xyz_pred, prob_pred, class_pred, res, **kw):
    """The script's op sequence (eval_joint.py:201-263) on CPU torch tensors.  grid_obj is modified in
    place like in the script.  Returns (boxes [K,8,3] f32, scores [K] f32, classes [K] i64, iterations)."""
    import torch
    p = _params(kw)
    res = float(res)
    el = int(p["elimination"])
    hi = el + 1 if p["elim_hi_inclusive"] else el
    corners = torch.stack([torch.min(points, 0)[0], torch.max(points, 0)[0]])          # :201
    bbox_raw = torch.from_numpy(BBOX_RAW.astype(np.float64)).float()                     # :203
    shape = torch.tensor(grid_obj.shape)
    boxes, scores, classes, iters = [], [], [], 0
    while True:
        flat = int(torch.argmax(grid_obj))                                                # :205
        cand = torch.tensor(np.unravel_index(flat, tuple(grid_obj.shape)))
        cand_world = torch.stack([corners[0, k] + res * cand[k] for k in range(3)])       # :206
        if grid_obj[cand[0], cand[1], cand[2]].item() < p["thresh_high"]:                 # :208
            break
        iters += 1
        lo = [max(int(cand[k]) - el, 0) for k in range(3)]
        grid_obj[lo[0]:int(cand[0]) + hi, lo[1]:int(cand[1]) + hi, lo[2]:int(cand[2]) + hi] = 0   # :211
        rot_vec = grid_rot[cand[0], cand[1], cand[2]]
        rot = torch.atan2(rot_vec[1], rot_vec[0])                                         # :214
        c, s = torch.cos(rot), torch.sin(rot)
        rm = torch.tensor([[c, 0, -s], [0, 1, 0], [s, 0, c]])                             # :215
        scale = grid_scale[cand[0], cand[1], cand[2]]                                     # :216
        bbox = (rm @ torch.diag(scale) @ bbox_raw.T).T                                    # :219
        bvol = (torch.stack([torch.min(bbox, 0)[0], torch.max(bbox, 0)[0]]) / res).int()  # :220
        rng = [torch.arange(int(bvol[0, k]), int(bvol[1, k]) + 1) for k in range(3)]
        cc = torch.stack(torch.meshgrid(*rng, indexing="ij"), -1).reshape(-1, 3) + cand    # :221-222
        cc = torch.max(torch.min(cc, shape - 1), torch.zeros(3, dtype=cc.dtype))           # :223
        inv = (((cc - cand) * res) @ rm) / scale                                           # :225
        m = ((-1 < inv) & (inv < 1)).all(-1)                                               # :226-228
        sel = cc[m]
        inv_w = ((points - cand_world) @ rm) / scale                                       # :231
        mw = ((-1 < inv_w) & (inv_w < 1)).all(-1)                                          # :232-234
        grid_obj[sel[:, 0], sel[:, 1], sel[:, 2]] = 0                                      # :243
        conf = prob_pred[mw] > p["prob_thresh"]                                            # :245
        if torch.sum(conf) < p["valid_ratio"] * torch.sum(mw) or torch.sum(mw) < p["thresh_low"]:   # :246
            continue
        gt = inv_w[mw][conf]
        err = torch.mean(torch.norm(xyz_pred[mw][conf] - gt, dim=-1) * prob_pred[mw][conf]).item()   # :250
        if err > p["err_thresh"]:                                                          # :252
            continue
        elems, counts = torch.unique(class_pred[mw][conf], return_counts=True)            # :255
        classes.append(int(elems[torch.argmax(counts)]))
        scores.append(float(torch.max(prob_pred[mw])))                                     # :258
        boxes.append((bbox + cand_world).numpy().copy())                                   # :259
        if len(boxes) >= p["max_boxes"]:
            break
    return (np.asarray(boxes, np.float32).reshape(-1, 8, 3), np.asarray(scores, np.float32),
            np.asarray(classes, np.int64), iters)


def loop_numpy(grid_obj, grid_rot, grid_scale, points, xyz_pred, prob_pred, class_pred, res, corner=None,
               return_trace=False, **kw):
    """Explicit float32 restatement (see module docstring).  Arrays are numpy; grid_obj is modified in place.
    `corner` defaults to min(points, 0) (eval_joint.py:201).  With return_trace also returns, per iteration,
    (cand flat index, n_in, n_conf, accepted) -- the integer decisions the GPU test compares exactly."""
    p = _params(kw)
    f32 = np.float32
    res32 = f32(res)
    el = int(p["elimination"])
    hi = el + 1 if p["elim_hi_inclusive"] else el
    X, Y, Z = grid_obj.shape
    points = np.ascontiguousarray(points, f32)
    xyz_pred = np.ascontiguousarray(xyz_pred, f32)
    prob_pred = np.ascontiguousarray(prob_pred, f32)
    class_pred = np.asarray(class_pred)
    if corner is None:
        corner = points.min(0)
    corner = np.asarray(corner, f32)
    boxes, scores, classes, trace, iters = [], [], [], [], 0
    while True:
        flat = int(np.argmax(grid_obj))                       # first maximum, like torch.argmax on CPU
        cx, cy, cz = np.unravel_index(flat, (X, Y, Z))
        if grid_obj[cx, cy, cz] < f32(p["thresh_high"]):
            break
        iters += 1
        cand = np.array([cx, cy, cz], np.int64)
        cand_world = corner + res32 * cand.astype(f32)        # float32 product, then float32 add  (:206)
        grid_obj[max(cx - el, 0):cx + hi, max(cy - el, 0):cy + hi, max(cz - el, 0):cz + hi] = 0
        rv = grid_rot[cx, cy, cz]
        rot = np.arctan2(rv[1], rv[0], dtype=f32)
        c, s = np.cos(rot, dtype=f32), np.sin(rot, dtype=f32)
        sc = grid_scale[cx, cy, cz].astype(f32)
        # box corners Rm @ diag(scale) @ raw  (:219): x' = c*(sx*rx) - s*(sz*rz), y' = sy*ry, z' = s*(sx*rx) + c*(sz*rz)
        ex, ey, ez = sc[0] * BBOX_RAW[:, 0], sc[1] * BBOX_RAW[:, 1], sc[2] * BBOX_RAW[:, 2]
        bbox = np.stack([c * ex + (-s) * ez, ey, s * ex + c * ez], -1).astype(f32)
        bmin = np.trunc(bbox.min(0) / res32).astype(np.int64)   # .int() truncates toward zero  (:220)
        bmax = np.trunc(bbox.max(0) / res32).astype(np.int64)
        # voxels of the clamped bounding volume whose inverse-transformed offset is strictly inside the unit box
        lo = np.clip(cand + bmin, 0, [X - 1, Y - 1, Z - 1])
        hi_ = np.clip(cand + bmax, 0, [X - 1, Y - 1, Z - 1])
        if np.all(bmax >= bmin):
            gx, gy, gz = np.meshgrid(np.arange(lo[0], hi_[0] + 1), np.arange(lo[1], hi_[1] + 1),
                                     np.arange(lo[2], hi_[2] + 1), indexing="ij")
            dx = (gx - cx).astype(f32) * res32
            dy = (gy - cy).astype(f32) * res32
            dz = (gz - cz).astype(f32) * res32
            qx = (dx * c + dz * s) / sc[0]
            qy = dy / sc[1]
            qz = (dx * (-s) + dz * c) / sc[2]
            m = (-1 < qx) & (qx < 1) & (-1 < qy) & (qy < 1) & (-1 < qz) & (qz < 1)
            grid_obj[gx[m], gy[m], gz[m]] = 0
        d = points - cand_world
        qx = (d[:, 0] * c + d[:, 2] * s) / sc[0]
        qy = d[:, 1] / sc[1]
        qz = (d[:, 0] * (-s) + d[:, 2] * c) / sc[2]
        mw = (-1 < qx) & (qx < 1) & (-1 < qy) & (qy < 1) & (-1 < qz) & (qz < 1)
        n_in = int(mw.sum())
        conf = mw & (prob_pred > f32(p["prob_thresh"]))
        n_conf = int(conf.sum())
        accepted = False
        if not (f32(n_conf) < f32(p["valid_ratio"]) * f32(n_in) or n_in < p["thresh_low"]):
            ex_ = xyz_pred[conf, 0] - qx[conf]
            ey_ = xyz_pred[conf, 1] - qy[conf]
            ez_ = xyz_pred[conf, 2] - qz[conf]
            nrm = np.sqrt(ex_ * ex_ + ey_ * ey_ + ez_ * ez_, dtype=f32)
            err = float(np.sum((nrm * prob_pred[conf]).astype(np.float64)) / n_conf)
            if not err > p["err_thresh"]:
                cls = np.bincount(class_pred[conf].astype(np.int64))
                classes.append(int(np.argmax(cls)))           # smallest class id on ties, like unique+argmax
                scores.append(float(prob_pred[mw].max()))
                boxes.append(bbox + cand_world)
                accepted = True
        trace.append((flat, n_in, n_conf, accepted))
        if len(boxes) >= p["max_boxes"]:
            break
    out = (np.asarray(boxes, f32).reshape(-1, 8, 3), np.asarray(scores, f32), np.asarray(classes, np.int64), iters)
    if return_trace:
        return out + (trace,)
    return out
