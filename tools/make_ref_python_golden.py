"""tools/make_ref_python_golden.py -- golden vectors from the reference's OWN python code, executed here on the CPU.

Runs in the build container (needs /root/reference).  The candidate loop with the LCC back-projection check, the head
decode, the per-class NMS and the detection metric are inline script code / module-level functions of files that cannot be
imported offline (eval_joint.py imports hydra + MinkowskiEngine + hv_cuda at the top, utils/calc_map.py imports shapely).
This script therefore cuts the relevant LINE RANGES out of the reference files where they lie, wraps each in a function and
executes it with a device shim only: the literal substrings `.to('cuda')` and `.cuda()` are removed so that the code runs
on CPU tensors; no arithmetic, control flow or constant is touched, and nothing of the reference is written into this
repository except the resulting numbers.

    eval_joint.py:173-190   head decode                       -> tests/golden/refpy_decode.npz
    eval_joint.py:196-268   candidate loop (+ :19-22 thresholds, :60-66 unravel_index)   -> tests/golden/refpy_loop_*.npz
    eval_separate.py:203-260 the per-category variant of the loop (zeroes [c-2, c+2), threshold 60 hard-coded, no class vote)
                                                              -> tests/golden/refpy_loop_sep.npz
    eval_joint.py:265-281   detection tuples of a scene (arrays, per-class nms, idx2name / name2catname :113-135)
                                                              -> tests/golden/refpy_scene_tuples.npz
    eval_joint.py:75-89     nms  (IoU injected: oracle/obb_nms.get_iou_obb, shapely is absent)
    utils/calc_map.py:40-71,78-168  voc_ap, eval_det_cls (IoU injected likewise)          -> tests/golden/refpy_metric.npz
    train_joint.py:253-282          joint loss of the training step (xyz_component_weights 1,1,1; factors of config.yaml)
                                                                                          -> tests/golden/refpy_loss.npz
    train_joint.py:128-133,200-206,224  learning-rate and BN-momentum schedules           -> tests/golden/refpy_schedules.npz
    train_joint.py:78-90            collate_fn (ME.utils.batched_coordinates = this repository's)  -> tests/golden/refpy_collate.npz
    utils/minkunet.py:122-180       MinkUNetBase.forward, the wiring of the U-Net: the reference class (imported unmodified on the
                                    MinkowskiEngine/ compat package) runs on CPU with the sparse ops routed to the convolution oracle
                                                                                          -> tests/golden/refpy_unet_wiring.npz
    sunrgbd/brnetcanon.py:119-161   vote-map proposal sampler (torch.multinomial replaced by recorded draws; :86-91 unravel_index)
                                                                                          -> tests/golden/refpy_proposals.npz

tests/test_oracle_refpy.py checks oracle/candidate_loop.py, canonicalvoting_b200/minkunet.decode_heads, oracle/obb_nms.py
and oracle/detection_metric.py against these files on every CPU run.
"""
import os
import re
import sys
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from canonicalvoting_b200 import synthetic  # noqa: E402
from oracle import hv_oracle as O  # noqa: E402
from oracle import obb_nms as ON  # noqa: E402


def cut(path, first, last):
    """Lines first..last (1-based, inclusive) of a reference file, dedented, with the device shim applied."""
    lines = open(os.path.join(REF, path)).read().splitlines()[first - 1:last]
    src = textwrap.dedent("\n".join(lines))
    return src.replace(".to('cuda')", "").replace(".cuda()", "")


def make_function(name, args, body, returns, env):
    src = "def %s(%s):\n%s\n    return %s\n" % (name, ", ".join(args), textwrap.indent(body, "    "), returns)
    exec(compile(src, "<reference:%s>" % name, "exec"), env)
    return env[name]


def reference_env():
    env = {"torch": torch, "np": np, "cfg": types.SimpleNamespace(scannet_res=0.03, log_scale=True), "nclasses": 9}
    ej = open(os.path.join(REF, "eval_joint.py")).read()
    # module-level pieces: thresholds (:19-22), unravel_index, nms -- located by their text, executed unchanged
    for pat in (r"^thresh_high = .*$", r"^thresh_low = .*$", r"^valid_ratio = .*$", r"^elimination = .*$"):
        exec(re.search(pat, ej, re.M).group(0), env)
    for fn in ("unravel_index", "nms"):
        m = re.search(r"^def %s\(.*?(?=^\S)" % fn, ej, re.M | re.S)
        exec(m.group(0), env)
    env["get_iou_obb"] = ON.get_iou_obb
    return env


class CpuSparse:
    """Stand-in for ME.SparseTensor on the CPU: features + the shared {tensor stride: coordinates} table."""

    def __init__(self, F, cm, ts):
        self.F, self.coordinate_manager, self.tensor_stride = F, cm, ts

    def _like(self, F, tensor_stride=None):
        return CpuSparse(F, self.coordinate_manager, self.tensor_stride if tensor_stride is None else tensor_stride)

    def __add__(self, other):
        return self._like(self.F + other.F)

    __iadd__ = __add__


def unet_wiring_golden():
    """Import utils/minkunet.py of the reference UNMODIFIED (on this repository's MinkowskiEngine/ compat package, whose module
    constructors work on the CPU), route the two convolution module types to oracle/sparse_oracle.py for the duration of the
    call, and run the reference's own forward().  What this pins: the order of layers, the skip connections and the
    concatenation order of MinkUNetBase.forward (utils/minkunet.py:122-180) as OracleNet / canonicalvoting_b200.minkunet
    restate them.  What it cannot pin: MinkowskiEngine's own operator semantics (the oracle's)."""
    sys.path.insert(1, REF)
    import utils.minkunet as ref_unet
    from canonicalvoting_b200.sparse import modules as M
    from oracle import sparse_oracle as SO

    def conv_forward(self, x):
        w = self.kernel.detach()
        b = self.bias.detach() if self.bias is not None else None
        C, ts = x.coordinate_manager, x.tensor_stride
        if self.kernel_size == 1:
            return x._like(x.F @ w + (b if b is not None else 0))
        if self.stride == 2:
            coarse, out = SO.conv_down(C[ts], x.F, w, ts, b)
            C[2 * ts] = coarse
            return x._like(out, 2 * ts)
        return x._like(SO.conv_same(C[ts], x.F, w, self.kernel_size, ts, b))

    def convtr_forward(self, x):
        C, ts = x.coordinate_manager, x.tensor_stride
        return x._like(SO.conv_up(C[ts // 2], x.F, self.kernel.detach(), ts, self.bias.detach() if self.bias is not None else None), ts // 2)

    saved = (M.MinkowskiConvolution.forward, M.MinkowskiConvolutionTranspose.forward)
    M.MinkowskiConvolution.forward, M.MinkowskiConvolutionTranspose.forward = conv_forward, convtr_forward
    try:
        out = {}
        for name, n, G in (("MinkUNet14A", 500, 20), ("MinkUNet34C", 400, 18)):
            torch.manual_seed(11)
            model = getattr(ref_unet, name)(3, 20).eval()
            g = torch.Generator().manual_seed(5)
            with torch.no_grad():
                for m in model.modules():
                    if isinstance(m, torch.nn.BatchNorm1d):
                        m.running_mean.normal_(0, 0.1, generator=g)
                        m.running_var.uniform_(0.5, 1.5, generator=g)
            lin = torch.randperm(G ** 3, generator=g)[:n]
            coords = torch.stack([lin % 2, lin // (G * G), (lin // G) % G, lin % G], 1).int()      # two scenes in the batch column
            feats = torch.randn(n, 3, generator=g)
            with torch.no_grad():
                y = model(CpuSparse(feats, {1: coords}, 1)).F
            print("unet wiring", name, tuple(y.shape), float(y.abs().max()))
            out.update({name + "_coords": coords.numpy(), name + "_feats": feats.numpy(), name + "_out": y.numpy()})
        np.savez_compressed(os.path.join(OUT, "refpy_unet_wiring.npz"), **out)
    finally:
        M.MinkowskiConvolution.forward, M.MinkowskiConvolutionTranspose.forward = saved


def main():
    env = reference_env()
    # ---- candidate loop: eval_joint.py:196-268, inputs are the names the script holds at that point
    loop_body = cut("eval_joint.py", 196, 268)
    ref_loop = make_function("ref_loop", ["grid_obj", "grid_rot", "grid_scale", "curr_points", "xyz_pred", "prob_pred", "class_pred"],
                             loop_body, "boxes, scores, probs, classes, grid_obj", env)
    cases = {"a": dict(n=8000, G=40, R=8, seed=3, n_objects=6), "b": dict(n=12000, G=48, R=8, seed=5, n_objects=12),
             "c": dict(n=6000, G=32, R=8, seed=2, n_objects=4)}
    for tag, c in cases.items():
        sc = synthetic.make_scene(c["n"], c["G"], c["R"], seed=c["seed"], n_objects=c["n_objects"])
        go, gr, gs = O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(0.03), c["R"], threads=1)
        env["thresh_high"] = 60.0 * c["R"] / 120            # the script's 60 assumes num_rots = 120 (eval_joint.py:18)
        t = torch.from_numpy
        boxes, scores, probs, classes, grid_after = ref_loop(t(go.copy()), t(gr), t(gs), t(sc["coords"].astype(np.int64)), t(sc["xyz"]),
                                                            t(sc["obj"]), t(sc["class_pred"]))
        print("loop", tag, "boxes", len(boxes), "classes", classes)
        # the vote grids are regenerated by the test from the seed (single-threaded C oracle, -ffp-contract=off: deterministic);
        # their digest is stored so that a platform on which they differ fails with a clear message
        import hashlib
        digest = hashlib.sha1(go.tobytes() + gr.tobytes() + gs.tobytes()).hexdigest()
        np.savez_compressed(os.path.join(OUT, "refpy_loop_%s.npz" % tag), grids_sha1=digest,
                            n=c["n"], G=c["G"], R=c["R"], seed=c["seed"], n_objects=c["n_objects"], thresh_high=env["thresh_high"],
                            boxes=np.asarray(boxes, np.float32).reshape(-1, 8, 3), scores=np.asarray(scores, np.float64),
                            classes=np.asarray(classes, np.int64),
                            zeroed=np.flatnonzero(grid_after.numpy().reshape(-1) != go.reshape(-1)).astype(np.int32))
        if tag == "a":      # per-class NMS of the script (eval_joint.py:270-281 calls nms(:75-89) class by class)
            rng = np.random.default_rng(0)
            many = np.concatenate([boxes + rng.normal(0, 0.03, boxes.shape).astype(np.float32) for _ in range(4)])
            sc_ = rng.uniform(0.3, 1.0, len(many))
            cl_ = np.tile(classes, 4)
            picks = {int(k): [int(i) for i in env["nms"](many[cl_ == k], sc_[cl_ == k], 0.3)] for k in np.unique(cl_)}
            np.savez_compressed(os.path.join(OUT, "refpy_nms.npz"), boxes=many, scores=sc_, classes=cl_,
                                picks=np.array([(k, i) for k, v in picks.items() for i in v], np.int64))
    # ---- the loop of eval_separate.py (:203-260): names as the script holds them at :188-201
    senv2 = dict(env)
    senv2.update(scannet_res=0.03, elimination=2,
                 bbox_raw=torch.tensor([[1, 1, -1, -1, 1, 1, -1, -1], [1, 1, 1, 1, -1, -1, -1, -1], [1, -1, -1, 1, 1, -1, -1, 1]]).float().T)
    # (bbox_raw above is the value eval_separate.py builds at :148-150 with l = h = w = 2, scannet_res the one of :151)
    sep_loop = make_function("ref_loop_sep", ["grid_obj", "grid_rot", "grid_scale", "scan_points", "corners", "xyz_pred", "prob_pred"],
                             "boxes = []\nscores = []\nprobs = []\n" + cut("eval_separate.py", 203, 260), "boxes, scores, grid_obj", senv2)
    c = dict(n=8000, G=40, R=120, seed=3, n_objects=6)
    sc = synthetic.make_scene(c["n"], c["G"], c["R"], seed=c["seed"], n_objects=c["n_objects"])
    go, gr, gs = O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(0.03), c["R"], threads=1)
    t = torch.from_numpy
    pts = t(sc["coords"].astype(np.int64)) * 0.03                      # eval_separate.py:188
    corners = torch.stack([torch.min(pts, 0)[0], torch.max(pts, 0)[0]])
    boxes, scores, grid_after = sep_loop(t(go.copy()), t(gr), t(gs), pts, corners, t(sc["xyz"]), t(sc["obj"]))
    print("loop sep boxes", len(boxes))
    import hashlib
    np.savez_compressed(os.path.join(OUT, "refpy_loop_sep.npz"), grids_sha1=hashlib.sha1(go.tobytes() + gr.tobytes() + gs.tobytes()).hexdigest(),
                        boxes=np.asarray(boxes, np.float32).reshape(-1, 8, 3), scores=np.asarray(scores, np.float64),
                        zeroed=np.flatnonzero(grid_after.numpy().reshape(-1) != go.reshape(-1)).astype(np.int32), **c)

    # ---- detection tuples of a scene: eval_joint.py:265-281 (arrays, per-class nms, category names) on the boxes of loop case b
    denv = dict(env)
    ej = open(os.path.join(REF, "eval_joint.py")).read()
    for name in ("idx2name", "name2catname"):
        exec(re.search(r"^%s = \{.*?^\}" % name, ej, re.M | re.S).group(0), denv)
    denv["SCENENN"] = False
    sc = synthetic.make_scene(12000, 48, 8, seed=5, n_objects=12)
    go, gr, gs = O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], np.float32(0.03), 8, threads=1)
    env["thresh_high"] = 60.0 * 8 / 120
    t = torch.from_numpy
    boxes, scores, probs, classes, _ = ref_loop(t(go.copy()), t(gr), t(gs), t(sc["coords"].astype(np.int64)), t(sc["xyz"]), t(sc["obj"]), t(sc["class_pred"]))
    rng = np.random.default_rng(1)                     # duplicate every box with jitter so that the NMS has something to suppress
    boxes = [b for b in boxes] + [b + rng.normal(0, 0.02, b.shape).astype(np.float32) for b in boxes]
    scores = list(scores) + [float(s_ - 0.01 * (i + 1)) for i, s_ in enumerate(scores)]
    probs, classes = list(scores), list(classes) + list(classes)
    tuples = make_function("ref_scene_tuples", ["boxes", "scores", "probs", "classes"], "map_scene = []\n" + cut("eval_joint.py", 265, 281),
                           "map_scene", denv)(boxes, scores, probs, classes)
    cats = sorted(set(denv["name2catname"][denv["idx2name"][i]] for i in range(9)))
    np.savez_compressed(os.path.join(OUT, "refpy_scene_tuples.npz"), boxes=np.asarray(boxes, np.float32), scores=np.asarray(scores, np.float64),
                        classes=np.asarray(classes, np.int64), names=np.array([c_ for c_, _, _ in tuples]),
                        out_boxes=np.asarray([b_ for _, b_, _ in tuples], np.float32), out_probs=np.asarray([p_ for _, _, p_ in tuples], np.float64),
                        categories=np.array([denv["name2catname"][denv["idx2name"][i]] for i in range(9)]))
    print("scene tuples", len(tuples), "of", len(boxes))

    # ---- head decode: eval_joint.py:173-190 on random network outputs
    decode = make_function("ref_decode", ["scan_output", "scan_points"], cut("eval_joint.py", 173, 190), "xyz_pred, scale_pred, class_pred, prob_pred", env)
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(500, 64, generator=g)
    xyz, scale, cls, prob = decode(types.SimpleNamespace(F=feats), torch.zeros(500, 4, dtype=torch.int32))
    np.savez_compressed(os.path.join(OUT, "refpy_decode.npz"), feats=feats.numpy(), xyz=xyz.numpy(), scale=scale.numpy(), cls=cls.numpy(),
                        prob=prob.numpy())
    # ---- detection metric: voc_ap + eval_det_cls of utils/calc_map.py, IoU = oracle polygon clipping
    menv = {"np": np, "get_iou_obb": ON.get_iou_obb}
    cm = open(os.path.join(REF, "utils/calc_map.py")).read()
    for fn in ("voc_ap", "get_iou_main", "eval_det_cls"):
        exec(re.search(r"^def %s\(.*?(?=^\S)" % fn, cm, re.M | re.S).group(0), menv)
    from tests.test_oracle_map import random_eval_case
    rows = []
    for seed in (0, 1, 2, 3):
        pred_all, gt_all = random_eval_case(seed)
        for cat in ("chair", "table", "sofa"):
            pred = {s: [(b, sc_) for c_, b, sc_ in v if c_ == cat] for s, v in pred_all.items()}
            pred = {s: v for s, v in pred.items() if v}
            gt = {s: [b for c_, b in v if c_ == cat] for s, v in gt_all.items()}
            for s in pred:
                gt.setdefault(s, [])
            if not pred:
                continue
            for thr in (0.25, 0.5):
                rec, prec, ap = menv["eval_det_cls"](pred, gt, thr, False, ON.get_iou_obb)
                rows.append((seed, ("chair", "table", "sofa").index(cat), thr, ap, rec[-1], prec[-1], menv["voc_ap"](rec, prec, True)))
    np.savez_compressed(os.path.join(OUT, "refpy_metric.npz"), rows=np.array(rows, np.float64))
    print("metric rows", len(rows))

    # ---- joint loss: train_joint.py:253-282 (class gather, masked MSE x 2, cross entropy, sum)
    lenv = {"torch": torch, "nclasses": 9, "cfg": types.SimpleNamespace(log_scale=True, xyz_factor=1.0, scale_factor=1.0),
            "obj_criterion": torch.nn.CrossEntropyLoss(), "xyz_weights": torch.tensor([1.0, 1.0, 1.0]), "losses": {}}
    ref_loss = make_function("ref_loss", ["scan_output", "scan_xyz_labels", "scan_scale_labels", "scan_class_labels"],
                             cut("train_joint.py", 253, 282), "loss", lenv)
    g = torch.Generator().manual_seed(0)
    n = 300
    out_f = torch.randn(n, 64, generator=g, requires_grad=True)
    xyz_l, scale_l = torch.randn(n, 3, generator=g), torch.rand(n, 3, generator=g) + 0.1
    cls_l = torch.randint(0, 10, (n,), generator=g)
    loss = ref_loss(types.SimpleNamespace(F=out_f), xyz_l, scale_l, cls_l)
    loss.backward()
    np.savez_compressed(os.path.join(OUT, "refpy_loss.npz"), out=out_f.detach().numpy(), xyz=xyz_l.numpy(), scale=scale_l.numpy(), cls=cls_l.numpy(),
                        loss=float(loss), grad=out_f.grad.numpy())
    print("loss", float(loss))

    # ---- schedules: get_current_lr (train_joint.py:128-133) and the bn_lbmd lambda (:224) with the constants of :200-206 / config.yaml
    tj = open(os.path.join(REF, "train_joint.py")).read()
    senv = {"BN_MOMENTUM_INIT": 0.5, "BN_MOMENTUM_MAX": 0.001, "BN_DECAY_STEP": 20, "BN_DECAY_RATE": 0.5, "BASE_LEARNING_RATE": 1e-3,
            "LR_DECAY_STEPS": [80, 120, 160], "LR_DECAY_RATES": [0.1, 0.1, 0.1]}
    exec(re.search(r"^def get_current_lr\(.*?(?=^\S)", tj, re.M | re.S).group(0), senv)
    exec(textwrap.dedent(re.search(r"^\s*bn_lbmd = lambda.*$", tj, re.M).group(0)), senv)
    epochs = np.arange(0, 201)
    np.savez_compressed(os.path.join(OUT, "refpy_schedules.npz"), epochs=epochs, lr=np.array([senv["get_current_lr"](int(e)) for e in epochs]),
                        bn=np.array([senv["bn_lbmd"](int(e)) for e in epochs]))
    print("schedules", senv["get_current_lr"](130), senv["bn_lbmd"](45))

    # ---- collate_fn of train_joint.py (:78-90) on two synthetic scenes
    import MinkowskiEngine as ME_compat
    cenv = {"torch": torch, "np": np, "ME": ME_compat}
    exec(re.search(r"^def collate_fn\(.*?(?=^\S)", tj, re.M | re.S).group(0), cenv)
    scenes = [synthetic.make_scene(300, 16, 4, seed=s_) for s_ in (0, 1)]
    batch = [("id%d" % i, torch.from_numpy(s_["coords"]), s_["feats"], s_["xyz_labels"], s_["scale_labels"], s_["class_labels"]) for i, s_ in enumerate(scenes)]
    _, cb, fb, xb, sb, lb = cenv["collate_fn"](batch)
    np.savez_compressed(os.path.join(OUT, "refpy_collate.npz"), coords=cb.numpy(), feats=fb.numpy(), xyz=xb.numpy(), scale=sb.numpy(), cls=lb.numpy())
    print("collate", tuple(cb.shape), cb.dtype)

    # ---- U-Net wiring: the reference's MinkUNetBase.forward on CPU
    unet_wiring_golden()

    # ---- proposal sampler: the body of HoughVotingModule.forward after the vote, draws injected
    from tests.test_oracle_proposals import vote_case

    class TorchWithDraws:
        """`torch` as the reference code sees it: everything is the real module except multinomial, which replays draws."""
        def __init__(self, draws):
            self._draws, self.used = list(draws), 0

        def multinomial(self, dist, n, replacement=True):
            d = self._draws[self.used]
            assert len(d) == n
            self.used += 1
            return d

        def __getattr__(self, name):
            return getattr(torch, name)

    bc = open(os.path.join(REF, "sunrgbd/brnetcanon.py")).read()
    penv = {"np": np}
    exec(re.search(r"^def unravel_index\(.*?(?=^\S)", bc, re.M | re.S).group(0), penv)
    body = cut("sunrgbd/brnetcanon.py", 119, 161)
    out = {}
    for tag, (n, G, R, seed, num_proposal, n_seeds) in {"a": (5000, 32, 4, 0, 64, 256), "b": (12000, 48, 8, 5, 150, 256),
                                                          "c": (5000, 32, 4, 1, 64, 4)}.items():     # c: few seeds -> several rejection trials
        hv_map, hv_scale, corner0, seeds, draws = vote_case(n, G, R, seed, n_seeds=n_seeds, n_draw=int(num_proposal * 1.5), trials=40)     # :133
        fake = TorchWithDraws(draws)
        penv["torch"] = fake
        fn = make_function("ref_sampler", ["self", "hv_map", "hv_scale", "corners", "vote_points", "pow"], body, "candidates, probs, scales", penv)
        t = torch.from_numpy
        me = types.SimpleNamespace(num_proposal=num_proposal, res=torch.tensor(0.03, dtype=torch.float32))
        cand, probs, scales = fn(me, t(hv_map), t(hv_scale), torch.stack([t(corner0), t(corner0) + 1.0]), t(seeds), 0.5)
        print("proposals", tag, tuple(cand.shape), "trials", fake.used)
        out.update({"cand_" + tag: cand.numpy(), "scales_" + tag: scales.numpy(), "trials_" + tag: fake.used,
                    "args_" + tag: np.array([n, G, R, seed, num_proposal, n_seeds])})
    np.savez_compressed(os.path.join(OUT, "refpy_proposals.npz"), **out)


if __name__ == "__main__":
    main()
