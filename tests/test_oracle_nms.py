"""oracle/obb_nms.py pinned by analytic cases (the reference's shapely is not installable here: parity unpinned vs GEOS)."""
import numpy as np

from oracle import obb_nms as O


def _box(cx, cz, sx, sz, yaw, y0=0.0, y1=1.0):
    """8 corners like eval_joint.py:203-219 builds them: corners 0-3 top face (y1), 4-7 bottom face (y0)."""
    c, s = np.cos(yaw), np.sin(yaw)
    q = np.array([[-sx, -sz], [sx, -sz], [sx, sz], [-sx, sz]], np.float64) / 2
    xz = np.stack([c * q[:, 0] + s * q[:, 1] + cx, -s * q[:, 0] + c * q[:, 1] + cz], -1)
    top = np.stack([xz[:, 0], np.full(4, y1), xz[:, 1]], -1)
    bot = np.stack([xz[:, 0], np.full(4, y0), xz[:, 1]], -1)
    return np.concatenate([top, bot], 0).astype(np.float32)


def test_quad_intersection_analytic_cases():
    sq = np.array([[0, 0], [2, 0], [2, 2], [0, 2]], np.float64)
    assert abs(O.quad_intersection_area(sq, sq) - 4.0) < 1e-12
    assert abs(O.quad_intersection_area(sq, sq[::-1]) - 4.0) < 1e-12                      # orientation does not matter
    assert abs(O.quad_intersection_area(sq, sq + [1, 1]) - 1.0) < 1e-12                   # shifted: 1 x 1 overlap
    assert O.quad_intersection_area(sq, sq + [3, 0]) == 0.0                               # disjoint
    assert O.quad_intersection_area(sq, sq + [2, 0]) == 0.0                               # touching edge
    assert abs(O.quad_intersection_area(sq, sq * 0.5 + [0.5, 0.5]) - 1.0) < 1e-12         # containment
    diamond = np.array([[1, -0.5], [2.5, 1], [1, 2.5], [-0.5, 1]], np.float64)            # square of side 1.5*sqrt(2), rotated 45 deg
    # octagon = square minus four corner triangles with legs 0.5
    assert abs(O.quad_intersection_area(sq, diamond) - (4.0 - 4 * 0.125)) < 1e-12


def test_iou_and_nms_follow_the_reference_rules():
    a = _box(0, 0, 2, 2, 0.0)
    assert abs(O.get_iou_obb(a, a) - 1.0) < 1e-12
    b = _box(1, 0, 2, 2, 0.0)                      # half overlap in x: inter 2, union 6
    assert abs(O.get_iou_obb(a, b) - 2.0 / 6.0) < 1e-7
    c = _box(0, 0, 2, 2, 0.0, y0=0.5, y1=1.5)      # half overlap in y
    assert abs(O.get_iou_obb(a, c) - (4 * 0.5) / (4 + 4 - 2)) < 1e-7
    flipped = a.copy(); flipped[:4, 1], flipped[4:, 1] = 0.0, 1.0
    assert O.get_iou_obb(a, flipped) == 0.0        # top not above bottom -> 0 (utils/calc_map.py:13-14)
    boxes = np.stack([a, b, _box(5, 5, 1, 1, 0.3), _box(1, 0, 2, 2, 0.0)])
    scores = np.array([0.9, 0.8, 0.5, 0.8], np.float32)
    # highest score first; ties: the larger index is taken first (stable argsort, take the last); b and its twin suppress each other
    assert O.nms(boxes, scores, 0.3) == [0, 2]
    assert O.nms(boxes, scores, 0.34) == [0, 3, 2]
    classes = np.array([1, 0, 1, 0])
    assert O.nms_per_class(boxes, scores, classes, 3, 0.3) == [3, 0, 2]


def test_intersection_area_against_qhull():
    """The oracle's Sutherland-Hodgman clipping against an independent construction of the same area (scipy / Qhull: the
    intersection of the 8 half-planes of two rectangles, Chebyshev centre by linear programming) on random oriented boxes --
    the stand-in for the shapely polygons of utils/calc_map.py:15-19, which cannot be installed here."""
    from scipy.optimize import linprog
    from scipy.spatial import ConvexHull, HalfspaceIntersection

    def halfplanes(q):                        # rows [a, b, c]: a x + b y + c <= 0 inside, for a quad in either orientation
        c0 = q.mean(0)
        rows = []
        for i in range(4):
            p, r = q[i], q[(i + 1) % 4]
            nrm = np.array([r[1] - p[1], -(r[0] - p[0])])
            nrm /= np.linalg.norm(nrm)
            off = -nrm @ p
            if nrm @ c0 + off > 0:
                nrm, off = -nrm, -off
            rows.append([nrm[0], nrm[1], off])
        return np.array(rows)

    def qhull_area(q1, q2):
        H = np.vstack([halfplanes(q1), halfplanes(q2)])
        # Chebyshev centre: maximise r subject to a.x + r <= -c
        res = linprog([0, 0, -1], A_ub=np.hstack([H[:, :2], np.ones((8, 1))]), b_ub=-H[:, 2], bounds=[(None, None), (None, None), (0, None)])
        if not res.success or res.x[2] < 1e-9:
            return 0.0
        return ConvexHull(HalfspaceIntersection(H, res.x[:2]).intersections).volume

    rng = np.random.default_rng(3)
    checked = overlapping = 0
    for _ in range(300):
        b1 = _box(rng.uniform(0, 2), rng.uniform(0, 2), rng.uniform(0.2, 1.5), rng.uniform(0.2, 1.5), rng.uniform(0, 2 * np.pi))
        b2 = _box(rng.uniform(0, 2), rng.uniform(0, 2), rng.uniform(0.2, 1.5), rng.uniform(0.2, 1.5), rng.uniform(0, 2 * np.pi))
        q1, q2 = b1[:4][:, [0, 2]].astype(float), b2[:4][:, [0, 2]].astype(float)
        got, want = O.quad_intersection_area(q1, q2), qhull_area(q1, q2)
        assert abs(got - want) <= 1e-9 + 1e-9 * want, (got, want)
        checked += 1
        overlapping += want > 1e-6
    assert checked == 300 and overlapping > 100
