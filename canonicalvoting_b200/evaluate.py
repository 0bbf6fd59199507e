"""Detection bookkeeping after the candidate loop -- the tail of eval_joint.py (:265-312) and the part of
utils/calc_map.py it calls (:40-71 `voc_ap`, :78-168 `eval_det_cls`, :177-226 `eval_det_multiprocessing`), plus the
scene sharding of the multi-GPU eval sweep (scene i -> rank i mod world, one gather of the detection lists).

    dets = scene_detections(boxes, probs, classes, keep)          # [(category, box [8,3], prob)], eval_joint.py:270-281
    pred_all, gt_all = gather_detections(pred_local, gt_local)   # all ranks -> every rank (torch.distributed)
    ret = compute_map(pred_all, gt_all, ovthresh=0.25)           # eval_joint.py:92-110, same keys

The reference evaluates one IoU per (detection, ground-truth box) pair through shapely inside a multiprocessing pool
(calc_map.py:139-141,214).  Here every (scene, class) gets ONE device call -- `obb.iou_matrix`, float64 polygon clipping
(csrc/obb_nms.cu) -- and the greedy matching walks the score-sorted detections on the host over those matrices.  There
is no CPU fallback: without `iou_matrix_fn` the boxes go to the device.
"""
import numpy as np

# eval_joint.py:113-135 (idx2name composed with name2catname)
CATEGORIES = ("others", "display", "table", "bathtub", "trashbin", "sofa", "chair", "cabinet", "bookshelf")

# unit box corners, eval_joint.py:203 (l = h = w = 2): rows = corners, 0-3 top face (+y), 4-7 bottom face
BBOX_RAW = np.array([[1, 1, -1, -1, 1, 1, -1, -1], [1, 1, 1, 1, -1, -1, -1, -1], [1, -1, -1, 1, 1, -1, -1, 1]], dtype=np.float32).T


def gt_box(tx, ty, tz, ry, sx, sy, sz):
    """Corners [8,3] (float64) of a ground-truth box line `tx ty tz ry sx sy sz ...` (eval_joint.py:288,299)."""
    c, s = np.cos(ry), np.sin(ry)
    rot = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])
    return (rot @ np.diag([sx, sy, sz]) @ BBOX_RAW.T).T + np.array([tx, ty, tz])


def scene_detections(boxes, probs, classes, keep, categories=CATEGORIES, allowed=None):
    """eval_joint.py:270-281: the boxes that survive the per-class NMS as (category, box, prob) tuples, class by class in
    pick order.  `keep` is what hough_voting.nms_per_class returns (indices into boxes); `allowed` restricts the
    categories (the SceneNN subset, :272-273)."""
    boxes, probs, classes = (np.asarray(a.cpu() if hasattr(a, "cpu") else a) for a in (boxes, probs, classes))
    out = []
    for j in np.asarray(keep.cpu() if hasattr(keep, "cpu") else keep, dtype=np.int64):
        name = categories[int(classes[j])]
        if allowed is None or name in allowed:
            out.append((name, boxes[j], float(probs[j])))
    return out


def voc_ap(rec, prec, use_07_metric=False):
    """utils/calc_map.py:40-71."""
    rec, prec = np.asarray(rec, dtype=np.float64), np.asarray(prec, dtype=np.float64)
    if use_07_metric:
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            sel = rec >= t
            ap += (prec[sel].max() if sel.any() else 0.0) / 11.0
        return ap
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]                 # precision envelope
    i = np.nonzero(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def _device_iou_matrix(a, b):
    import torch
    from . import obb
    ta = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    tb = torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).cuda()
    return obb.iou_matrix(ta, tb).cpu().numpy()


def eval_det_cls(pred, gt, ovthresh=0.25, use_07_metric=False, iou_matrix_fn=None):
    """Precision / recall / AP of one class (utils/calc_map.py:78-168).
    pred: {scene: [(box, score)]}, gt: {scene: [box]} -> (rec [nd], prec [nd], ap)."""
    iou_matrix_fn = iou_matrix_fn or _device_iou_matrix
    npos = sum(len(v) for v in gt.values())
    scenes, conf, row = [], [], []
    best_iou, best_gt = {}, {}
    for sid, dets in pred.items():
        if not dets:
            continue
        g = gt.get(sid, [])
        if len(g):
            m = np.asarray(iou_matrix_fn(np.stack([np.asarray(b) for b, _ in dets]), np.stack([np.asarray(b) for b in g])))
            best_gt[sid] = m.argmax(1)                             # first maximum, like the strict `iou > ovmax` scan (:140-142)
            best_iou[sid] = m.max(1)
        for r, (_, score) in enumerate(dets):
            scenes.append(sid)
            conf.append(score)
            row.append(r)
    nd = len(conf)
    order = np.argsort(-np.asarray(conf, dtype=np.float64)) if nd else np.zeros(0, dtype=np.int64)   # :120
    taken = {sid: np.zeros(len(g), dtype=bool) for sid, g in gt.items()}
    tp, fp = np.zeros(nd), np.zeros(nd)
    for d, k in enumerate(order):
        sid, r = scenes[k], row[k]
        if sid in best_iou and best_iou[sid][r] > ovthresh and not taken[sid][best_gt[sid][r]]:
            tp[d] = 1.0
            taken[sid][best_gt[sid][r]] = True
        else:
            fp[d] = 1.0
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    with np.errstate(divide="ignore", invalid="ignore"):
        rec = tp / float(npos)                                     # npos == 0 gives inf / nan as in the reference (:160)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric)


def eval_det(pred_all, gt_all, ovthresh=0.25, use_07_metric=False, iou_matrix_fn=None):
    """utils/calc_map.py:177-226 without the process pool.  pred_all: {scene: [(category, box, score)]}, gt_all:
    {scene: [(category, box)]} -> (rec, prec, ap) dicts keyed by category.  A category with ground truth but no
    detection gets 0 / 0 / 0 (:219-222).  (The reference indexes the pool results by the position of the category among ALL
    ground-truth categories, :217-218, which mixes categories up as soon as one of them has no detection; results here are
    keyed by category.)"""
    pred, gt = {}, {}
    for sid, dets in pred_all.items():
        for name, box, score in dets:
            pred.setdefault(name, {}).setdefault(sid, []).append((box, score))
            gt.setdefault(name, {}).setdefault(sid, [])
    for sid, objs in gt_all.items():
        for name, box in objs:
            gt.setdefault(name, {}).setdefault(sid, []).append(box)
    rec, prec, ap = {}, {}, {}
    for name in gt:
        if name in pred:
            rec[name], prec[name], ap[name] = eval_det_cls(pred[name], gt[name], ovthresh, use_07_metric, iou_matrix_fn)
        else:
            rec[name], prec[name], ap[name] = 0, 0, 0
    return rec, prec, ap


def compute_map(pred_map_cls, gt_map_cls, ovthresh=0.5, iou_matrix_fn=None):
    """eval_joint.py:92-110: '<cat> Average Precision', '<cat> Recall', 'mAP', 'AR'."""
    rec, prec, ap = eval_det(pred_map_cls, gt_map_cls, ovthresh, iou_matrix_fn=iou_matrix_fn)
    ret = {}
    for key in sorted(ap):
        ret["%s Average Precision" % key] = ap[key]
    ret["mAP"] = np.mean(list(ap.values()))
    recalls = []
    for key in sorted(ap):
        try:
            r = rec[key][-1]
        except (TypeError, IndexError):
            r = 0
        ret["%s Recall" % key] = r
        recalls.append(r)
    ret["AR"] = np.mean(recalls)
    return ret


def gather_detections(pred_local, gt_local=None, group=None):
    """Merge the per-rank {scene: detections} maps on every rank (scenes are disjoint across ranks: scene i lives on rank
    i mod world, train.shard_scenes).  One all_gather_object of python lists -- the only communication of the eval sweep."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(pred_local), dict(gt_local or {})
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, (pred_local, gt_local or {}), group=group)
    pred_all, gt_all = {}, {}
    for p, g in parts:
        for sid, v in p.items():
            if sid in pred_all:
                raise RuntimeError("gather_detections: scene %r evaluated on two ranks" % (sid,))
            pred_all[sid] = v
        gt_all.update(g)
    return pred_all, gt_all
