"""oracle/obb_nms.py -- TEST INFRASTRUCTURE (never on the product path).

CPU restatement of the detection post-process that follows the candidate loop: the oriented-box IoU
`get_iou_obb` (utils/calc_map.py:6-21) and the per-class greedy `nms` (eval_joint.py:75-89, used at :270-280).
The reference computes the xz-rectangle intersection with shapely (GEOS), which is NOT installed here and not
vendored: **parity unpinned** against shapely itself.  The polygon intersection is restated as Sutherland-Hodgman
clipping of convex quadrilaterals in float64 and pinned by analytic cases (tests/test_oracle_nms.py: axis-aligned
overlaps, a 45-degree rotated square, containment, disjoint and touching boxes) and by an independent construction of the
same area with scipy / Qhull on random rectangle pairs; `nms` / `nms_per_class` are pinned by the pick lists of the
reference's own `nms` function executed verbatim (tests/golden/refpy_nms.npz, tests/test_oracle_refpy.py)."""
import numpy as np


def _signed_area(p):
    x, z = p[:, 0], p[:, 1]
    return 0.5 * float(np.sum(x * np.roll(z, -1) - np.roll(x, -1) * z))


def _clip(subject, a, b, sign):
    """Keep the part of polygon `subject` on the inner side of the directed edge a->b (sign = orientation of the clipper)."""
    out = []
    n = len(subject)
    for i in range(n):
        p, q = subject[i], subject[(i + 1) % n]
        dp = sign * ((b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]))
        dq = sign * ((b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0]))
        if dp >= 0:
            out.append(p)
        if (dp >= 0) != (dq >= 0):
            t = dp / (dp - dq)
            out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return out


def quad_intersection_area(p1, p2):
    """Area of the intersection of two convex quadrilaterals given as [4,2] arrays (any orientation)."""
    p1 = np.asarray(p1, np.float64)
    p2 = np.asarray(p2, np.float64)
    sign = 1.0 if _signed_area(p2) >= 0 else -1.0
    poly = [tuple(v) for v in p1]
    for i in range(4):
        if not poly:
            return 0.0
        poly = _clip(poly, p2[i], p2[(i + 1) % 4], sign)
    if len(poly) < 3:
        return 0.0
    return abs(_signed_area(np.asarray(poly, np.float64)))


def get_iou_obb(bbox1, bbox2):
    """utils/calc_map.py:6-21: corners 0-3 = top face, 4-7 = bottom face, rectangles in the xz plane."""
    bbox1 = np.asarray(bbox1, np.float64)
    bbox2 = np.asarray(bbox2, np.float64)
    if not (bbox1[0, 1] > bbox1[4, 1] and bbox2[0, 1] > bbox2[4, 1]):
        return 0.0
    q1 = np.stack([bbox1[:4, 0], bbox1[:4, 2]], -1)
    q2 = np.stack([bbox2[:4, 0], bbox2[:4, 2]], -1)
    inter_area = quad_intersection_area(q1, q2)
    inter_vol = inter_area * max(0.0, min(bbox1[0, 1], bbox2[0, 1]) - max(bbox1[4, 1], bbox2[4, 1]))
    a1, a2 = abs(_signed_area(q1)), abs(_signed_area(q2))
    return inter_vol / (a1 * (bbox1[0, 1] - bbox1[4, 1]) + a2 * (bbox2[0, 1] - bbox2[4, 1]) - inter_vol)


def nms(boxes, scores, overlap_threshold):
    """eval_joint.py:75-89, verbatim control flow (argsort made stable so that ties are defined)."""
    I = np.argsort(scores, kind="stable")
    pick = []
    while I.size != 0:
        last = I.size
        i = I[-1]
        pick.append(int(i))
        suppress = [last - 1]
        for pos in range(last - 1):
            j = I[pos]
            if get_iou_obb(boxes[i], boxes[j]) > overlap_threshold:
                suppress.append(pos)
        I = np.delete(I, suppress)
    return pick


def nms_per_class(boxes, scores, classes, nclasses, overlap_threshold=0.3):
    """eval_joint.py:270-280: indices into the input arrays, class by class, in pick order."""
    boxes, scores, classes = np.asarray(boxes), np.asarray(scores), np.asarray(classes)
    out = []
    for c in range(nclasses):
        sel = np.nonzero(classes == c)[0]
        if len(sel):
            out += [int(sel[j]) for j in nms(boxes[sel], scores[sel], overlap_threshold)]
    return out
