"""oracle/detection_metric.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Independent restatement of the reference's detection metric: what utils/calc_map.py computes at :40-71 (AP from a
precision / recall curve, VOC-07 11-point or area under the precision envelope), :78-168 (greedy matching of score-sorted
detections of one class to ground-truth boxes, one match per box) and :177-226 (grouping by class; without the process pool,
results keyed by class), and the summary dictionary of eval_joint.py:92-110.  Written as plain loops over explicit records,
IoUs one pair at a time from oracle/obb_nms.get_iou_obb (float64 polygon clipping; shapely is absent here).  The reference
holds no golden vectors for this code: pinned by hand-computed cases (tests/test_oracle_map.py) and by golden vectors from the
reference functions `voc_ap` / `eval_det_cls` executed verbatim with the same IoU injected (tests/golden/refpy_metric.npz).
"""
import numpy as np

from .obb_nms import get_iou_obb


def average_precision(recall, precision, eleven_point=False):
    recall, precision = np.asarray(recall, float), np.asarray(precision, float)
    if eleven_point:
        total = 0.0
        for level in np.arange(0.0, 1.1, 0.1):
            reached = precision[recall >= level]
            total += (reached.max() if reached.size else 0.0) / 11.0
        return total
    r = np.r_[0.0, recall, 1.0]
    p = np.r_[0.0, precision, 0.0]
    for i in reversed(range(len(p) - 1)):          # envelope: best precision at this recall or beyond
        if p[i + 1] > p[i]:
            p[i] = p[i + 1]
    area = 0.0
    for i in range(len(r) - 1):
        if r[i + 1] != r[i]:
            area += (r[i + 1] - r[i]) * p[i + 1]
    return area


def match_class(dets_by_scene, gts_by_scene, iou_threshold=0.25, eleven_point=False):
    """dets_by_scene {scene: [(box, score)]}, gts_by_scene {scene: [box]} -> (recall [nd], precision [nd], ap)."""
    n_gt = sum(len(v) for v in gts_by_scene.values())
    records = [(scene, np.asarray(box, float), score) for scene, dets in dets_by_scene.items() for box, score in dets]
    ranking = np.argsort(-np.array([rec[2] for rec in records])) if records else []
    claimed = {scene: [False] * len(v) for scene, v in gts_by_scene.items()}
    hits, misses = [], []
    for idx in ranking:
        scene, box, _ = records[idx]
        best, best_j = -np.inf, None
        for j, gt_box in enumerate(gts_by_scene.get(scene, [])):
            v = get_iou_obb(box, np.asarray(gt_box, float))
            if v > best:                               # strict: the first of equal overlaps is kept
                best, best_j = v, j
        hit = best > iou_threshold and not claimed[scene][best_j]
        if hit:
            claimed[scene][best_j] = True
        hits.append(1.0 if hit else 0.0)
        misses.append(0.0 if hit else 1.0)
    tp, fp = np.cumsum(hits), np.cumsum(misses)
    with np.errstate(divide="ignore", invalid="ignore"):
        recall = tp / float(n_gt)                      # no ground truth at all: inf / nan, as the reference produces
    precision = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return recall, precision, average_precision(recall, precision, eleven_point)


def evaluate(pred_all, gt_all, iou_threshold=0.25, eleven_point=False):
    """pred_all {scene: [(cls, box, score)]}, gt_all {scene: [(cls, box)]} -> (recall, precision, ap) dicts keyed by class.
    Classes appear in the order the reference creates them: first through the detections, then through the ground truth."""
    dets, gts = {}, {}
    for scene, items in pred_all.items():
        for cls, box, score in items:
            dets.setdefault(cls, {}).setdefault(scene, []).append((box, score))
            gts.setdefault(cls, {}).setdefault(scene, [])
    for scene, items in gt_all.items():
        for cls, box in items:
            gts.setdefault(cls, {}).setdefault(scene, []).append(box)
    recall, precision, ap = {}, {}, {}
    for cls in gts:
        if cls in dets:
            recall[cls], precision[cls], ap[cls] = match_class(dets[cls], gts[cls], iou_threshold, eleven_point)
        else:
            recall[cls] = precision[cls] = ap[cls] = 0
    return recall, precision, ap


def compute_map(pred_all, gt_all, iou_threshold=0.5):
    """The summary of eval_joint.py:92-110: per-class AP and final recall, their means."""
    recall, _, ap = evaluate(pred_all, gt_all, iou_threshold)
    out = {"%s Average Precision" % c: ap[c] for c in sorted(ap)}
    out["mAP"] = np.mean(list(ap.values()))
    finals = []
    for c in sorted(ap):
        r = recall[c]
        finals.append(r[-1] if isinstance(r, np.ndarray) and r.size else 0)
        out["%s Recall" % c] = finals[-1]
    out["AR"] = np.mean(finals)
    return out
