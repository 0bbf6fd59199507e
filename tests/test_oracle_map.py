"""CPU tests of the detection metric: oracle/detection_metric.py (the reference's control flow restated, utils/calc_map.py:40-226,
eval_joint.py:92-110) on constructed cases with known answers, the host bookkeeping of canonicalvoting_b200/evaluate.py
against it (IoU matrices injected from the oracle -- the product computes them on the device), and the world_size-2 gloo
run of the sharded evaluation (scene i -> rank i mod 2, one all_gather_object)."""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from canonicalvoting_b200 import evaluate as E
from canonicalvoting_b200 import train
from oracle import detection_metric as OM
from oracle import obb_nms as ON


def oracle_iou_matrix(a, b):
    return np.array([[ON.get_iou_obb(np.asarray(x, float), np.asarray(y, float)) for y in b] for x in a])


def random_eval_case(seed, n_scenes=6, cats=("chair", "table", "sofa")):
    """Ground-truth boxes per scene and detections = jittered ground truth + clutter, with score ties."""
    rng = np.random.default_rng(seed)
    pred_all, gt_all = {}, {}
    for s in range(n_scenes):
        gts, dets = [], []
        for _ in range(rng.integers(0, 6)):
            cat = cats[rng.integers(len(cats))]
            p = (rng.uniform(0, 4), rng.uniform(0.3, 0.8), rng.uniform(0, 4), rng.uniform(0, 2 * np.pi), *rng.uniform(0.2, 0.7, 3))
            gts.append((cat, E.gt_box(*p)))
            for _ in range(rng.integers(0, 3)):                      # 0-2 detections near this object (duplicates -> FP)
                q = np.array(p) + rng.normal(0, [0.05, 0.02, 0.05, 0.1, 0.03, 0.03, 0.03])
                dets.append((cat if rng.uniform() < 0.85 else cats[rng.integers(len(cats))], E.gt_box(*q).astype(np.float32),
                             float(np.round(rng.uniform(0.3, 1.0), 1))))
        for _ in range(rng.integers(0, 3)):                          # clutter
            q = (rng.uniform(0, 4), 0.5, rng.uniform(0, 4), rng.uniform(0, 6), *rng.uniform(0.2, 0.7, 3))
            dets.append((cats[rng.integers(len(cats))], E.gt_box(*q).astype(np.float32), float(np.round(rng.uniform(0.3, 1.0), 1))))
        if s == 2:
            dets.append(("bathtub", E.gt_box(1, 0.5, 1, 0, 0.4, 0.4, 0.4).astype(np.float32), 0.9))   # detections of a class without any GT
        if s == 3:
            gts.append(("bookshelf", E.gt_box(2, 0.5, 2, 0.3, 0.4, 0.8, 0.2)))                          # GT of a class never detected
        pred_all["scene%04d" % s] = dets
        gt_all["scene%04d" % s] = gts
    return pred_all, gt_all


def test_voc_ap_known_values():
    rec, prec = np.array([0.5, 0.5, 1.0]), np.array([1.0, 0.5, 2.0 / 3.0])
    assert abs(OM.average_precision(rec, prec) - (0.5 * 1.0 + 0.5 * 2.0 / 3.0)) < 1e-12
    assert abs(E.voc_ap(rec, prec) - OM.average_precision(rec, prec)) < 1e-12
    assert abs(OM.average_precision(rec, prec, True) - (6 * 1.0 + 5 * 2.0 / 3.0) / 11.0) < 1e-12
    assert abs(E.voc_ap(rec, prec, True) - OM.average_precision(rec, prec, True)) < 1e-12
    assert E.voc_ap(np.zeros(0), np.zeros(0)) == 0.0 == OM.average_precision(np.zeros(0), np.zeros(0))


def test_constructed_scene_has_the_expected_ap():
    box = lambda x: E.gt_box(x, 0.5, 0.0, 0.0, 0.5, 0.5, 0.5)
    gt = {"a": [box(0.0), box(3.0)], "b": [box(0.0)]}
    # scores descending: exact hit (TP), duplicate of the same object (FP), far miss (FP), hit in scene b (TP); object at x=3 never found
    pred = {"a": [(box(0.0), 0.9), (box(0.05), 0.8), (box(10.0), 0.7)], "b": [(box(0.1), 0.6)]}
    rec, prec, ap = OM.match_class(pred, gt, 0.25)
    np.testing.assert_allclose(rec, [1 / 3, 1 / 3, 1 / 3, 2 / 3])
    np.testing.assert_allclose(prec, [1.0, 0.5, 1 / 3, 0.5])
    assert abs(ap - (1 / 3 * 1.0 + 1 / 3 * 0.5)) < 1e-12
    rec2, prec2, ap2 = E.eval_det_cls(pred, gt, 0.25, iou_matrix_fn=oracle_iou_matrix)
    np.testing.assert_array_equal(rec2, rec)
    np.testing.assert_array_equal(prec2, prec)
    assert abs(ap2 - ap) < 1e-12


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("thresh", [0.25, 0.5])
def test_host_bookkeeping_matches_oracle(seed, thresh):
    pred_all, gt_all = random_eval_case(seed)
    want = OM.compute_map(pred_all, gt_all, thresh)
    got = E.compute_map(pred_all, gt_all, thresh, iou_matrix_fn=oracle_iou_matrix)
    assert list(got) == list(want)
    for k in want:
        # the two sum the area under the envelope in different orders: AP may differ in the last bit
        np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), np.asarray(want[k], dtype=np.float64), rtol=0, atol=1e-12, err_msg=k)
    assert want["bookshelf Average Precision"] == 0 and "bathtub Average Precision" in want


def test_gt_box_and_scene_detections():
    b = E.gt_box(1.0, 2.0, 3.0, 0.0, 0.5, 0.25, 0.125)
    np.testing.assert_allclose(b.max(0) - b.min(0), [1.0, 0.5, 0.25])
    np.testing.assert_allclose(b.mean(0), [1.0, 2.0, 3.0])
    assert (b[:4, 1] > b[4:, 1]).all()                               # corners 0-3 are the top face (what get_iou_obb assumes)
    boxes = np.stack([E.gt_box(i, 0.5, 0, 0, 0.3, 0.3, 0.3) for i in range(4)]).astype(np.float32)
    dets = E.scene_detections(boxes, np.array([0.9, 0.8, 0.7, 0.6]), np.array([6, 2, 6, 1]), np.array([2, 0, 1]))
    assert [d[0] for d in dets] == ["chair", "chair", "table"] and [d[2] for d in dets] == [0.7, 0.9, 0.8]
    assert E.scene_detections(boxes, np.ones(4), np.array([6, 2, 6, 1]), np.array([2, 0, 1]), allowed=("table",))[0][0] == "table"


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pred_all, gt_all = random_eval_case(7, n_scenes=7)
        ids = sorted(pred_all)
        mine = [ids[i] for i in train.shard_scenes(len(ids), rank, world)]
        p, g = E.gather_detections({s: pred_all[s] for s in mine}, {s: gt_all[s] for s in mine})
        assert sorted(p) == ids and sorted(g) == ids
        ret = E.compute_map({s: p[s] for s in ids}, {s: g[s] for s in ids}, 0.25, iou_matrix_fn=oracle_iou_matrix)
        if rank == 0:
            out.put({k: float(np.asarray(v)) for k, v in ret.items()})
    finally:
        dist.destroy_process_group()


def test_world_size_2_sharded_evaluation_equals_single_process():
    pred_all, gt_all = random_eval_case(7, n_scenes=7)
    want = OM.compute_map(pred_all, gt_all, 0.25)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 30100 + os.getpid() % 400          # disjoint from the range tests/test_train_host.py uses
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got.keys() == want.keys()
    for k in want:
        np.testing.assert_allclose(got[k], float(np.asarray(want[k])), rtol=0, atol=1e-12, err_msg=k)
