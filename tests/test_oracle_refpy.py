"""The oracles against golden vectors produced by the reference's OWN python code (tools/make_ref_python_golden.py cuts the
line ranges out of eval_joint.py / utils/calc_map.py where they lie in /root/reference and executes them on CPU tensors with
a device shim only; tests/golden/refpy_*.npz).  This is what pins the candidate-loop, head-decode, NMS and detection-metric
oracles to the reference itself rather than to a reading of it."""
import hashlib
import os

import numpy as np
import pytest
import torch

from canonicalvoting_b200 import synthetic
from oracle import candidate_loop as CL
from oracle import detection_metric as OM
from oracle import hv_oracle as O
from oracle import obb_nms as ON

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RES = 0.03


def _loop_case(tag):
    g = np.load(os.path.join(GOLD, "refpy_loop_%s.npz" % tag))
    sc = synthetic.make_scene(int(g["n"]), int(g["G"]), int(g["R"]), seed=int(g["seed"]), n_objects=int(g["n_objects"]))
    go, gr, gs = O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(RES), int(g["R"]), threads=1)
    digest = hashlib.sha1(go.tobytes() + gr.tobytes() + gs.tobytes()).hexdigest()
    assert digest == str(g["grids_sha1"]), "the single-threaded vote oracle does not reproduce the grids the golden loop ran on"
    return g, sc, go, gr, gs


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_candidate_loop_oracles_match_the_reference_script(tag):
    """eval_joint.py:196-268 executed verbatim vs oracle/candidate_loop.py: boxes, scores, classes and the set of zeroed voxels."""
    g, sc, go, gr, gs = _loop_case(tag)
    thresh = float(g["thresh_high"])
    assert len(g["boxes"]) >= 2
    # the explicit float32 restatement (what the GPU parity tests compare the CUDA kernel with)
    grid = go.copy()
    b, s, c, it = CL.loop_numpy(grid, gr, gs, sc["points"], sc["xyz"], sc["obj"], sc["class_pred"], RES, thresh_high=thresh)
    assert len(b) == len(g["boxes"])
    np.testing.assert_allclose(b, g["boxes"], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(np.asarray(s, np.float32), g["scores"].astype(np.float32))
    np.testing.assert_array_equal(c, g["classes"])
    np.testing.assert_array_equal(np.flatnonzero(grid.reshape(-1) != go.reshape(-1)).astype(np.int32), g["zeroed"])
    # the lifted torch op sequence
    t = torch.from_numpy
    grid_t = t(go.copy())
    b2, s2, c2, it2 = CL.loop_torch(grid_t, t(gr), t(gs), t(sc["points"]), t(sc["xyz"]), t(sc["obj"]), t(sc["class_pred"]), RES, thresh_high=thresh)
    np.testing.assert_allclose(np.asarray(b2, np.float32).reshape(-1, 8, 3), g["boxes"], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(np.asarray(c2), g["classes"])
    np.testing.assert_array_equal(np.flatnonzero(grid_t.numpy().reshape(-1) != go.reshape(-1)).astype(np.int32), g["zeroed"])
    assert it == it2


def test_head_decode_matches_the_reference_script():
    """eval_joint.py:173-190 executed verbatim vs canonicalvoting_b200.minkunet.decode_heads (the torch restatement the CUDA
    head-decode kernel is tested against)."""
    from canonicalvoting_b200.minkunet import decode_heads
    g = np.load(os.path.join(GOLD, "refpy_decode.npz"))
    xyz, scale, cls, prob = decode_heads(torch.from_numpy(g["feats"]), 9, True)
    np.testing.assert_array_equal(xyz.numpy(), g["xyz"])
    np.testing.assert_array_equal(scale.numpy(), g["scale"])
    np.testing.assert_array_equal(cls.numpy(), g["cls"])
    np.testing.assert_array_equal(prob.numpy(), g["prob"])


def test_per_class_nms_matches_the_reference_function():
    """eval_joint.py:75-89 `nms` executed verbatim, class by class (:270-281), vs oracle/obb_nms.nms_per_class."""
    g = np.load(os.path.join(GOLD, "refpy_nms.npz"))
    boxes, scores, classes = g["boxes"], g["scores"], g["classes"]
    want = []
    for k in np.unique(classes):
        local = [int(i) for kk, i in g["picks"] if kk == k]
        want += [int(np.flatnonzero(classes == k)[i]) for i in local]
    assert len(want) < len(boxes)                                   # something was suppressed
    assert ON.nms_per_class(boxes, scores, classes, 9, 0.3) == want


def test_detection_metric_matches_the_reference_functions():
    """utils/calc_map.py `eval_det_cls` + `voc_ap` executed verbatim (IoU injected) vs oracle/detection_metric.py."""
    from tests.test_oracle_map import random_eval_case
    rows = np.load(os.path.join(GOLD, "refpy_metric.npz"))["rows"]
    assert len(rows) == 24
    cats = ("chair", "table", "sofa")
    cache = {}
    for seed, ci, thr, ap, rec_last, prec_last, ap07 in rows:
        seed, cat = int(seed), cats[int(ci)]
        if seed not in cache:
            cache[seed] = random_eval_case(seed)
        pred_all, gt_all = cache[seed]
        pred = {s: [(b, sc) for c, b, sc in v if c == cat] for s, v in pred_all.items()}
        pred = {s: v for s, v in pred.items() if v}
        gt = {s: [b for c, b in v if c == cat] for s, v in gt_all.items()}
        rec, prec, got_ap = OM.match_class(pred, gt, float(thr))
        assert abs(got_ap - ap) < 1e-12 and rec[-1] == rec_last and prec[-1] == prec_last
        assert abs(OM.average_precision(rec, prec, True) - ap07) < 1e-12


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_proposal_sampler_oracles_match_the_reference_module(tag):
    """sunrgbd/brnetcanon.py:119-161 (HoughVotingModule.forward after the vote) executed verbatim with recorded multinomial
    draws vs oracle/proposals.py: proposal locations, scales and the number of rejection trials."""
    from oracle import proposals as P
    from tests.test_oracle_proposals import vote_case
    g = np.load(os.path.join(GOLD, "refpy_proposals.npz"))
    n, G, R, seed, num_proposal, n_seeds = (int(v) for v in g["args_" + tag])
    hv_map, hv_scale, corner0, seeds, draws = vote_case(n, G, R, seed, n_seeds=n_seeds, n_draw=int(num_proposal * 1.5), trials=40)
    c, s, used, dmins = P.sample_numpy(hv_map, hv_scale, RES, corner0, seeds, [d.numpy() for d in draws], num_proposal)
    assert used == int(g["trials_" + tag]) and (tag != "c" or used > 1)
    np.testing.assert_array_equal(c, g["cand_" + tag])
    np.testing.assert_array_equal(s, g["scales_" + tag])
    t = torch.from_numpy
    c2, s2, used2 = P.sample_torch(t(hv_map), t(hv_scale), RES, t(corner0), t(seeds), draws, num_proposal)
    assert used2 == used
    np.testing.assert_array_equal(c2.numpy(), g["cand_" + tag])
    np.testing.assert_array_equal(s2.numpy(), g["scales_" + tag])


def test_joint_loss_matches_the_reference_script():
    """train_joint.py:253-282 executed verbatim (value and gradient w.r.t. the network output) vs train.joint_loss."""
    from canonicalvoting_b200 import train
    g = np.load(os.path.join(GOLD, "refpy_loss.npz"))
    out = torch.from_numpy(g["out"]).requires_grad_(True)
    loss = train.joint_loss(out, torch.from_numpy(g["xyz"]), torch.from_numpy(g["scale"]), torch.from_numpy(g["cls"]))
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    np.testing.assert_allclose(out.grad.numpy(), g["grad"], rtol=1e-5, atol=1e-8)


def test_schedules_match_the_reference_script():
    """get_current_lr (train_joint.py:128-133) and the BN-momentum lambda (:224) executed verbatim for epochs 0..200."""
    from canonicalvoting_b200 import sparse as ME
    from canonicalvoting_b200 import train
    g = np.load(os.path.join(GOLD, "refpy_schedules.npz"))
    for e, lr, bn in zip(g["epochs"], g["lr"], g["bn"]):
        assert train.learning_rate(int(e)) == lr and train.bn_momentum(int(e)) == bn
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    assert train.adjust_learning_rate(opt, 125) == opt.param_groups[0]["lr"] == g["lr"][125]
    model = torch.nn.Sequential(ME.MinkowskiBatchNorm(8), torch.nn.ReLU())
    sched = train.BNMomentumScheduler(model, last_epoch=39)
    assert model[0].momentum == g["bn"][40] and model[0].bn.momentum == 0.1          # the wrapped BatchNorm1d is not reached
    sched.step()
    assert sched.last_epoch == 40 and model[0].momentum == g["bn"][40]
    sched.step(60)
    assert model[0].momentum == g["bn"][60]


def test_per_category_loop_variant_matches_eval_separate():
    """eval_separate.py:203-260 executed verbatim (zeroes [c-2, c+2), no class vote) vs the oracle with elim_hi_inclusive=False --
    the switch cvb200_bp_params carries for this script."""
    g = np.load(os.path.join(GOLD, "refpy_loop_sep.npz"))
    sc = synthetic.make_scene(int(g["n"]), int(g["G"]), int(g["R"]), seed=int(g["seed"]), n_objects=int(g["n_objects"]))
    go, gr, gs = O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(RES), int(g["R"]), threads=1)
    assert hashlib.sha1(go.tobytes() + gr.tobytes() + gs.tobytes()).hexdigest() == str(g["grids_sha1"])
    grid = go.copy()
    b, s, c, it = CL.loop_numpy(grid, gr, gs, sc["points"], sc["xyz"], sc["obj"], sc["class_pred"], RES, thresh_high=60.0, elim_hi_inclusive=False)
    assert len(b) == len(g["boxes"]) >= 3
    np.testing.assert_allclose(b, g["boxes"], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(np.asarray(s, np.float32), g["scores"].astype(np.float32))
    np.testing.assert_array_equal(np.flatnonzero(grid.reshape(-1) != go.reshape(-1)).astype(np.int32), g["zeroed"])
    # the inclusive variant (eval_joint.py) zeroes a different set: the switch matters on this scene
    grid2 = go.copy()
    CL.loop_numpy(grid2, gr, gs, sc["points"], sc["xyz"], sc["obj"], sc["class_pred"], RES, thresh_high=60.0)
    assert not np.array_equal(grid2, grid)


@pytest.mark.parametrize("name", ["MinkUNet14A", "MinkUNet34C"])
def test_unet_wiring_matches_the_reference_forward(name):
    """utils/minkunet.py:122-180 -- the reference class's own forward(), run on CPU with its convolutions routed to the oracle
    (tools/make_ref_python_golden.py::unet_wiring_golden) -- vs OracleNet on this repository's model built from the same
    seed: layer order, skip connections and concatenation order of the two restatements are the reference's."""
    import canonicalvoting_b200.minkunet as M
    from oracle import sparse_oracle as SO
    g = np.load(os.path.join(GOLD, "refpy_unet_wiring.npz"))
    torch.manual_seed(11)
    model = getattr(M, name)(3, 20).eval()
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1, generator=gen)
                m.running_var.uniform_(0.5, 1.5, generator=gen)
    coords, feats = torch.from_numpy(g[name + "_coords"]), torch.from_numpy(g[name + "_feats"])
    with torch.no_grad():
        y = SO.OracleNet(model).forward(coords, feats)
    want = g[name + "_out"]
    assert y.shape == want.shape
    np.testing.assert_allclose(y.numpy(), want, rtol=0, atol=1e-5 * float(np.abs(want).max()))


def test_collate_matches_the_reference_function():
    """train_joint.py:78-90 `collate_fn` executed verbatim vs train.collate on the same two scenes."""
    from canonicalvoting_b200 import train
    g = np.load(os.path.join(GOLD, "refpy_collate.npz"))
    coords, feats, xyz, scale, cls = train.collate([synthetic.make_scene(300, 16, 4, seed=s) for s in (0, 1)])
    assert coords.dtype == torch.int32 and cls.dtype == torch.int64
    for got, key in ((coords, "coords"), (feats, "feats"), (xyz, "xyz"), (scale, "scale"), (cls, "cls")):
        np.testing.assert_array_equal(got.numpy(), g[key])


def test_scene_detection_tuples_match_the_reference_script():
    """eval_joint.py:265-281 executed verbatim (arrays -> per-class nms -> (category, box, prob) tuples, with the script's
    idx2name / name2catname tables) vs oracle nms_per_class + evaluate.scene_detections and evaluate.CATEGORIES."""
    from canonicalvoting_b200 import evaluate as E
    g = np.load(os.path.join(GOLD, "refpy_scene_tuples.npz"))
    assert tuple(g["categories"]) == E.CATEGORIES
    keep = np.asarray(ON.nms_per_class(g["boxes"], g["scores"], g["classes"], 9, 0.3), dtype=np.int64)
    dets = E.scene_detections(g["boxes"], g["scores"], g["classes"], keep)
    assert 0 < len(dets) < len(g["boxes"])
    assert [d[0] for d in dets] == list(g["names"])
    np.testing.assert_array_equal(np.stack([d[1] for d in dets]), g["out_boxes"])
    np.testing.assert_array_equal(np.array([d[2] for d in dets]), g["out_probs"])
