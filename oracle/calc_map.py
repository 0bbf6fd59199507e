"""oracle/calc_map.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Plain restatement of the reference's detection metric, following its control flow line by line:
utils/calc_map.py:40-71 (`voc_ap`), :78-168 (`eval_det_cls`), :177-226 (`eval_det_multiprocessing`, without the
process pool and with results keyed by category instead of by pool position) and eval_joint.py:92-110 (`compute_map`).
IoUs come from oracle/obb_nms.get_iou_obb (float64 polygon clipping; shapely is absent here).  No golden vectors exist in
the reference for this code; pinned by constructed cases in tests/test_oracle_map.py: parity unpinned beyond those.
"""
import numpy as np

from .obb_nms import get_iou_obb


def voc_ap(rec, prec, use_07_metric=False):
    if use_07_metric:
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11.0
        return ap
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def eval_det_cls(pred, gt, ovthresh=0.25, use_07_metric=False):
    class_recs, npos = {}, 0
    for img_id in gt.keys():
        bbox = np.array(gt[img_id])
        npos += len(bbox)
        class_recs[img_id] = {"bbox": bbox, "det": [False] * len(bbox)}
    for img_id in pred.keys():
        if img_id not in gt:
            class_recs[img_id] = {"bbox": np.array([]), "det": []}
    image_ids, confidence, BB = [], [], []
    for img_id in pred.keys():
        for box, score in pred[img_id]:
            image_ids.append(img_id)
            confidence.append(score)
            BB.append(box)
    confidence, BB = np.array(confidence), np.array(BB)
    sorted_ind = np.argsort(-confidence)
    BB = BB[sorted_ind, ...]
    image_ids = [image_ids[x] for x in sorted_ind]
    nd = len(image_ids)
    tp, fp = np.zeros(nd), np.zeros(nd)
    for d in range(nd):
        R = class_recs[image_ids[d]]
        bb = BB[d, ...].astype(float)
        ovmax, jmax = -np.inf, -1
        BBGT = R["bbox"].astype(float)
        if BBGT.size > 0:
            for j in range(BBGT.shape[0]):
                iou = get_iou_obb(bb, BBGT[j, ...])
                if iou > ovmax:
                    ovmax, jmax = iou, j
        if ovmax > ovthresh:
            if not R["det"][jmax]:
                tp[d] = 1.0
                R["det"][jmax] = 1
            else:
                fp[d] = 1.0
        else:
            fp[d] = 1.0
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    with np.errstate(divide="ignore", invalid="ignore"):
        rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric)


def eval_det(pred_all, gt_all, ovthresh=0.25, use_07_metric=False):
    pred, gt = {}, {}
    for img_id in pred_all.keys():
        for classname, bbox, score in pred_all[img_id]:
            if classname not in pred:
                pred[classname] = {}
            if img_id not in pred[classname]:
                pred[classname][img_id] = []
            if classname not in gt:
                gt[classname] = {}
            if img_id not in gt[classname]:
                gt[classname][img_id] = []
            pred[classname][img_id].append((bbox, score))
    for img_id in gt_all.keys():
        for classname, bbox in gt_all[img_id]:
            if classname not in gt:
                gt[classname] = {}
            if img_id not in gt[classname]:
                gt[classname][img_id] = []
            gt[classname][img_id].append(bbox)
    rec, prec, ap = {}, {}, {}
    for classname in gt.keys():
        if classname in pred:
            rec[classname], prec[classname], ap[classname] = eval_det_cls(pred[classname], gt[classname], ovthresh, use_07_metric)
        else:
            rec[classname], prec[classname], ap[classname] = 0, 0, 0
    return rec, prec, ap


def compute_map(pred_map_cls, gt_map_cls, ovthresh=0.5):
    rec, prec, ap = eval_det(pred_map_cls, gt_map_cls, ovthresh)
    ret_dict = {}
    for key in sorted(ap.keys()):
        ret_dict["%s Average Precision" % str(key)] = ap[key]
    ret_dict["mAP"] = np.mean(list(ap.values()))
    rec_list = []
    for key in sorted(ap.keys()):
        try:
            ret_dict["%s Recall" % str(key)] = rec[key][-1]
            rec_list.append(rec[key][-1])
        except Exception:
            ret_dict["%s Recall" % str(key)] = 0
            rec_list.append(0)
    ret_dict["AR"] = np.mean(rec_list)
    return ret_dict
