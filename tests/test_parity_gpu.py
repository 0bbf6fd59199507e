"""GPU parity tests at BASELINE.json sizes and at the level of the product's results ("close the parity holes"):

  (b) MinkUNetEngine (tcgen05 kind::tf32, BatchNorm folded) on MinkUNet34C at config C2 (50 000 voxels) against the CPU oracle
      `oracle/sparse_oracle.py FastCpuNet` (fp32), with the tolerance that was measured, not a generous one;
  (c) detection-level equivalence: network -> head decode (eval_joint.py:173-190) -> hv_cuda.forward -> candidate loop with the
      LCC back-projection check (eval_joint.py:195-263) with the TF32 engine, with the exact-fp32 module path and with the CPU
      oracle chain: same class_pred (near-ties counted and reported), same boxes;
  (d) one training step in tf32 mode (train_joint.py:244-288): loss and every parameter gradient against torch autograd through
      the oracle network (oracle/sparse_oracle.py GradCpuNet, float64).
"""
import numpy as np
import pytest
import torch

from canonicalvoting_b200 import synthetic
from oracle import sparse_oracle as SO

pytestmark = pytest.mark.gpu

NC = 9
RES = 0.03


def _model34(seed=0):
    import bench
    return bench.make_model()            # MinkUNet34C(3, 64), seeded init, non-trivial BatchNorm statistics


def _scene_tensors(sc):
    coords = torch.cat([torch.zeros(len(sc["coords"]), 1, dtype=torch.int32), torch.from_numpy(sc["coords"])], 1).contiguous()
    feats = (torch.from_numpy(sc["feats"]) * 2.0 - 1.0).contiguous()          # eval_joint.py:168
    return coords, feats


def test_engine_C2_matches_cpu_oracle():
    """(b) 50 000-voxel scene, MinkUNet34C, the engine's TF32 program vs the fp32 CPU oracle.  Measured on B200: max error
    3.4e-3 of the output scale, rms 3.3e-4 of it; class_pred (argmax over the object logits, eval_joint.py:188) differs at
    15-17 of 50 000 points, every one a near-tie of the two best logits.  Asserted: 5e-3 / 1e-3 / near-ties only, <= 0.2 %."""
    from canonicalvoting_b200.engine import MinkUNetEngine
    from canonicalvoting_b200.minkunet import decode_heads
    sc = synthetic.make_config("C2", seed=0)
    coords, feats = _scene_tensors(sc)
    model = _model34()
    with torch.no_grad():
        want = SO.FastCpuNet(model).forward(coords, feats)
    eng = MinkUNetEngine(model.cuda(), NC, True)
    got = eng(coords.cuda(), feats.cuda()).cpu()
    assert got.shape == want.shape == (50000, 64)
    scale = float(want.abs().max())
    err = (got - want).abs()
    rms = float(err.square().mean().sqrt())
    print("engine C2 vs FastCpuNet: max err %.3e  rms %.3e  scale %.3e  (max/scale %.2e)" % (float(err.max()), rms, scale, float(err.max()) / scale))
    assert float(err.max()) <= 5e-3 * scale, "max err %.3e vs scale %.3e" % (float(err.max()), scale)
    assert rms <= 1e-3 * scale
    # integer outputs of the decode: class_pred may only differ where the two best object logits are within the error bound
    _, _, cls_w, prob_w = decode_heads(want, NC, True)
    _, _, cls_g, prob_g = decode_heads(got, NC, True)
    diff = torch.nonzero(cls_w != cls_g)[:, 0]
    if len(diff):
        top2 = want[diff, 6 * NC:7 * NC].topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1])
        assert float(margin.max()) <= 2 * float(err.max()), "class_pred differs away from a near-tie (margin %.3e)" % float(margin.max())
    print("class_pred: %d of %d points differ (all near-ties)" % (len(diff), len(cls_w)))
    assert len(diff) <= 0.002 * len(cls_w)
    assert float((prob_w - prob_g).abs().max()) <= 5e-3


def _planted_heads(sc):
    """64-channel head tensor P with decode(P) == the scene's planted per-point predictions (eval_joint.py:173-190 inverted):
    xyz[k] / log scale[k] in the slots of the point's class k, logits such that softmax(logits)[k] = obj with the background
    logit as the alternative."""
    n = len(sc["points"])
    P = torch.zeros(n, 7 * NC + 1)
    k = torch.from_numpy(sc["class_pred"]).long()
    obj = torch.from_numpy(sc["obj"]).double().clamp(1e-4, 1 - 1e-4)
    is_obj = obj > 0.5
    logits = torch.full((n, NC + 1), -20.0)
    a = torch.log(obj / (1 - obj)).float()
    logits[torch.arange(n), k] = a
    logits[:, NC] = 0.0
    P[:, 6 * NC:] = logits
    # decode picks the slot of argmax over all 10 logits (background -> slot 0, eval_joint.py:178)
    slot = torch.where(is_obj, k, torch.zeros_like(k))
    xyz, scale = torch.from_numpy(sc["xyz"]), torch.from_numpy(sc["scale"])
    for d in range(3):
        P[torch.arange(n), 3 * slot + d] = xyz[:, d]
        P[torch.arange(n), 3 * NC + 3 * slot + d] = torch.log(scale[:, d])
    return P


def _boxes_match(a, b, tol):
    """Greedy one-to-one matching of two box lists ([K,8,3] corner arrays) by max corner distance."""
    if len(a) != len(b):
        return False, "counts %d vs %d" % (len(a), len(b))
    used = set()
    worst = 0.0
    for i in range(len(a)):
        d = [float(np.abs(a[i] - b[j]).max()) if j not in used else np.inf for j in range(len(b))]
        j = int(np.argmin(d))
        if d[j] > tol:
            return False, "box %d has no partner within %.3g (best %.3g)" % (i, tol, d[j])
        used.add(j)
        worst = max(worst, d[j])
    return True, worst


@pytest.mark.parametrize("n,G,R,seed", [(20000, 64, 12, 1), (30000, 96, 12, 4)])
def test_detections_survive_tf32(n, G, R, seed):
    """(c) What TF32 does to the product's RESULT.  No trained checkpoint exists offline and a randomly initialised network detects
    nothing, so the head tensor of a planted-object scene (decode(P) = the planted predictions) is perturbed by the network's own
    output, H = P + alpha * net(scene), once with the TF32 engine, once with the exact-fp32 module path and once with the fp32 CPU
    oracle; alpha scales the perturbation to 0.1 (the LCC noise level of the generator).  The three head tensors differ exactly by
    the arithmetic of the network path; everything downstream must give the same detections."""
    import hv_cuda
    from canonicalvoting_b200 import sparse as ME
    from canonicalvoting_b200.engine import MinkUNetEngine
    from canonicalvoting_b200.hough_voting import back_project
    from canonicalvoting_b200.minkunet import decode_heads
    from oracle import candidate_loop as CL
    from oracle import hv_oracle as O
    sc = synthetic.make_scene(n, G, R, seed=seed)
    coords, feats = _scene_tensors(sc)
    model = _model34()
    with torch.no_grad():
        f_cpu = SO.FastCpuNet(model).forward(coords, feats)
    model = model.cuda()
    eng = MinkUNetEngine(model, NC, True)
    f_tf32 = eng(coords.cuda(), feats.cuda())
    ME.set_forward_mode("fp32")
    with torch.no_grad():
        f_fp32 = model(ME.SparseTensor(feats, coords, device="cuda")).F
    alpha = 0.1 / float(f_cpu.abs().max())
    P = _planted_heads(sc)
    thresh = 60.0 * R / 120
    res_t = torch.tensor(RES, dtype=torch.float32).cuda()
    rots_t = torch.tensor(R, dtype=torch.int32).cuda()
    pts = torch.from_numpy(sc["points"]).cuda()

    def gpu_arm(f):
        H = (P.cuda() + alpha * f).contiguous()
        xyz, scale, cls, prob = eng.decode(H)
        go, gr, gs = hv_cuda.forward(pts, xyz, scale, prob, res_t, rots_t)
        boxes, scores, classes = back_project(go, gr, gs, pts, xyz, prob, cls, RES, thresh_high=thresh)
        return cls.cpu(), boxes.cpu().numpy(), scores.cpu().numpy(), classes.cpu().numpy()

    cls_t, box_t, sco_t, kls_t = gpu_arm(f_tf32)
    cls_f, box_f, sco_f, kls_f = gpu_arm(f_fp32)
    # CPU oracle chain on the fp32 CPU features
    Hc = P + alpha * f_cpu
    xyz_c, scale_c, cls_c, prob_c = decode_heads(Hc, NC, True)
    go, gr, gs = O.forward(sc["points"], xyz_c.numpy(), scale_c.numpy(), prob_c.numpy(), np.float32(RES), R)
    box_c, sco_c, kls_c, iters = CL.loop_numpy(go.copy(), gr, gs, sc["points"], xyz_c.numpy(), prob_c.numpy(), cls_c.numpy(), RES,
                                               thresh_high=thresh)
    assert len(box_c) >= 3, "the planted scene must produce detections (got %d)" % len(box_c)
    # class_pred: integer output of the decode
    n_tf = int((cls_t != cls_f).sum())
    n_cpu = int((cls_f != cls_c).sum())
    print("class_pred differing points: tf32 vs fp32 %d, fp32 GPU vs CPU oracle %d (of %d)" % (n_tf, n_cpu, n))
    assert n_tf <= 0.001 * n and n_cpu <= 0.001 * n
    # box lists: same count, same classes, corners within 1.5 cm (half a voxel), scores within 1e-3
    for name, (b, s, k) in (("tf32 engine", (box_t, sco_t, kls_t)), ("fp32 module path", (box_f, sco_f, kls_f))):
        ok, info = _boxes_match(np.asarray(box_c), b, 0.015)
        assert ok, "%s vs CPU oracle: %s" % (name, info)
        assert sorted(np.asarray(kls_c).tolist()) == sorted(k.tolist()), name
        assert np.allclose(np.sort(np.asarray(sco_c)), np.sort(s), atol=1e-3), name
        print("%s: %d boxes identical to the oracle chain (max corner deviation %.2e m)" % (name, len(b), info))
    ok, info = _boxes_match(box_f, box_t, 0.015)
    assert ok, "tf32 vs fp32 on the GPU: %s" % info


def _grad_report(model, want):
    """Deviation of model.grad from the reference gradients: per tensor ||g - w|| / ||w|| (worst, median), and over all parameters
    together the relative L2 error and the cosine -- the direction an optimiser step takes."""
    rel, num, den, dot, gg = {}, 0.0, 0.0, 0.0, 0.0
    for name, p in model.named_parameters():
        assert name in want and want[name] is not None, name
        g, w = p.grad.detach().cpu().double(), want[name]
        assert g.shape == w.shape, name
        rel[name] = float((g - w).norm()) / max(float(w.norm()), 1e-30)
        num += float((g - w).square().sum()); den += float(w.square().sum())
        dot += float((g * w).sum()); gg += float(g.square().sum())
    worst = max(rel, key=rel.get)
    srt = sorted(rel.values())
    return {"worst": rel[worst], "worst_name": worst, "median": srt[len(srt) // 2], "global_rel": (num / den) ** 0.5,
            "cosine": dot / (gg * den) ** 0.5}


@pytest.mark.parametrize("bn_training", [False, True], ids=["bn-eval", "bn-train"])
def test_training_step_tf32_gradients_match_oracle_autograd(bn_training):
    """(d) One training step of the joint model (train_joint.py:244-288) with forward, input gradient and weight gradient on
    tcgen05 (tf32 mode): loss and parameter gradients vs torch autograd through the float64 oracle network.

    Two conditionings.  `bn-eval`: BatchNorm with running statistics -- the gradient of every tensor is a plain sum of products,
    so the deviation shows the arithmetic of the three kernels: fp32 mode agrees to rounding, tf32 mode to TF32's 10-bit
    operands.  `bn-train` (what the script does): batch statistics make the loss invariant to the scale of every convolution
    in front of a BatchNorm, their gradients are differences of large nearly cancelling sums, and ANY rounding is amplified --
    the exact-fp32 path itself deviates from float64 by 4e-3 per tensor there (measured, printed), tf32 by 1e-1 on single
    tensors and 7e-2 over the whole gradient, whose direction stays within cosine 0.997.  (That is TF32, not this kernel: the
    library default under autograd is therefore the exact-fp32 path; tf32 training is opt-in, sparse.set_forward_mode.)"""
    from canonicalvoting_b200 import sparse as ME
    from canonicalvoting_b200 import train as T
    from canonicalvoting_b200.minkunet import MinkUNet14A
    scenes = [synthetic.make_scene(12000, 80, 4, seed=s) for s in (21, 22)]       # >= 125 voxels per scene on the coarsest level
    coords, feats, xyz_l, scale_l, class_l = T.collate(scenes)
    torch.manual_seed(5)
    model = MinkUNet14A(3, 7 * NC + 1).cuda()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.05); m.running_var.uniform_(0.8, 1.2)
    model.train(bn_training)
    oracle = SO.GradCpuNet(model)
    out_o = oracle.forward(coords, (feats * 2.0 - 1.0).double())
    loss_o = T.joint_loss(out_o, xyz_l.double(), scale_l.double(), class_l)
    loss_o.backward()
    want, loss_ref = oracle.grads(), float(loss_o.detach())
    rep = {}
    for mode in ("fp32", "tf32"):
        ME.set_forward_mode(mode)
        try:
            model.zero_grad(set_to_none=True)
            out = model(ME.SparseTensor((feats * 2.0 - 1.0).cuda(), coords.cuda(), device="cuda"))
            loss = T.joint_loss(out.F, xyz_l.cuda(), scale_l.cuda(), class_l.cuda())
            loss.backward()
        finally:
            ME.set_forward_mode("fp32")
        r = _grad_report(model, want)
        r["loss"] = float(loss.detach())
        rep[mode] = r
        print("%s %s: loss %.6f vs oracle %.6f; gradients: global rel. error %.3e, cosine %.6f, per tensor worst %.3e (%s) median %.3e" % (
            "bn-train" if bn_training else "bn-eval", mode, r["loss"], loss_ref, r["global_rel"], r["cosine"], r["worst"], r["worst_name"], r["median"]))
    f, t = rep["fp32"], rep["tf32"]
    assert abs(f["loss"] - loss_ref) <= 1e-5 * abs(loss_ref) and abs(t["loss"] - loss_ref) <= 2e-3 * abs(loss_ref)
    # envelopes = what was measured on B200 (printed above) with a margin of 2-3x:
    #   bn-eval   fp32 1.1e-6 global / 2.2e-4 worst tensor;   tf32 1.4e-3 global, cosine 0.999999, 8.7e-2 worst tensor (a BatchNorm bias:
    #             a sum over all rows with cancellation), 2.1e-2 median
    #   bn-train  fp32 2.6e-3 global, cosine 0.999997;         tf32 7.1e-2 global, cosine 0.9975, 1.5e-1 worst tensor (ill-conditioned: wide margin)
    if not bn_training:
        assert f["global_rel"] <= 1e-5 and f["worst"] <= 1e-3, f
        assert t["global_rel"] <= 5e-3 and t["worst"] <= 0.3 and t["median"] <= 6e-2 and t["cosine"] >= 0.99999, t
    else:
        assert f["global_rel"] <= 8e-3 and f["cosine"] >= 0.9999, f
        assert t["cosine"] >= 0.99 and t["global_rel"] <= 0.2 and t["worst"] <= 0.5, t
