#!/usr/bin/env python
"""bench.py -- headline benchmark of the CanonicalVoting hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C1|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one synthetic scene: BASELINE.json configs[1] ("C2"): MinkUNet34C
forward on a 50 000-voxel scan -> head decode -> Hough voting (hv_cuda.forward) into a 128^3 grid with
num_rots = 12.  Prints ONE JSON line (rank 0).  DESIGN.md "Measurement" defines every field.

  value        scenes/s, whole job (all ranks), inputs resident in HBM, CUDA-event timed per step, L2 flushed between
               timed steps, max over ranks
  e2e          scenes/s through the reference-facing API starting from pinned HOST buffers: H2D of coords + feats ->
               engine -> decode -> hv_cuda.forward (with its geometry sync) -> D2H of the step's result; three scenes in
               flight (two host syncs per scene leave more bubbles to fill than the resident loop's two scenes)
  passes       the K-step loop runs three times back to back; `value` is the median pass, all three are in passes_ms_per_step
  roofline     the dominant kernel group of the step = the tcgen05 sparse-convolution program of the U-Net:
               algorithmic FLOPs 2 * sum(pairs * cin * cout) / its CUDA-event time, against the measured dense
               tensor peak; `vote` holds the HBM roofline of the vote op (40 N + 24 G bytes)
  cpu_baseline the oracle ports (oracle/sparse_oracle.py FastCpuNet + oracle/hv_oracle.c OpenMP) on the host cores,
               bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NCLASSES = 9
# DRAM bytes per call of the vote op (sum over its kernels) from this round's ncu --set full capture; None where no capture exists
VOTE_DRAM_BYTES_NCU = {"C2": 21725184 + 80819456}      # scatter + write-out, profiles/r2g_launches_scene_C2.csv (third scene)
# same for the convolution program: sum over its 63 launches of the third scene of that launch list (cold caches between launches)
CONV_DRAM_BYTES_NCU = {"C2": 569001472}


# ----------------------------------------------------------------------------- helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  ONE long-lived
    `nvidia-smi -lms` process started before the region: forking a fresh nvidia-smi per sample from this (large, GIL-holding)
    process stalled the launching threads for milliseconds and showed up as 2-3x outliers of a 50 ms measurement."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=50):
        self.idx, self.period, self.samples, self.proc = gpu_index, period_ms, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", str(self.period)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def mark(self):
        """Samples taken before this call are discarded (call right before the timed region)."""
        self._t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            out = ""
        lines = [ln for ln in out.splitlines() if ln.strip()]
        # keep the samples of the timed region: the process started `lead` seconds before mark()
        skip = int(max(0.0, (self._t_mark - self._t_start)) * 1000 / self.period) if hasattr(self, "_t_mark") else 0
        for ln in lines[min(skip, max(len(lines) - 1, 0)):]:
            self.samples.append([x.strip() for x in ln.split(",")])

    def __enter__(self):
        self._t_start = time.time()
        return self.start()

    def __exit__(self, *a):
        self.stop()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].lower().startswith("active")
                                                          for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def scene_for(workload, seed):
    from canonicalvoting_b200 import synthetic
    return synthetic.make_config(workload, seed=seed)


def vote_bytes(n, dims):
    """SURVEY.md 8d contract figure: read every input once (10 floats/point), write every output voxel once (6 floats)."""
    return 40 * n + 24 * int(dims[0]) * int(dims[1]) * int(dims[2])


def make_model():
    import torch
    from canonicalvoting_b200.minkunet import MinkUNet34C
    torch.manual_seed(0)
    model = MinkUNet34C(3, 6 * NCLASSES + NCLASSES + 1).eval()       # train_joint.py:218: random init, there is no checkpoint offline
    with torch.no_grad():                                             # non-trivial BN statistics so that folding is exercised
        g = torch.Generator().manual_seed(1)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.05, generator=g)
                m.running_var.uniform_(0.8, 1.2, generator=g)
    return model


def scene_tensors(sc):
    import torch
    coords = torch.cat([torch.zeros(len(sc["coords"]), 1, dtype=torch.int32), torch.from_numpy(sc["coords"])], 1).contiguous()
    feats = (torch.from_numpy(sc["feats"]) * 2.0 - 1.0).contiguous()      # eval_joint.py:168
    return coords, feats


def workload_config(workload, sc):
    G = sc["grid"]
    return {"workload": "%s: MinkUNet34C(3,64) forward + head decode + Hough voting (hv_cuda.forward) on one synthetic room "
                        "scene, N=%d voxels, vote grid %d^3, num_rots=%d" % (workload, len(sc["points"]), G, sc["num_rots"]),
            "points": len(sc["points"]), "grid": [G, G, G], "num_rots": sc["num_rots"], "res": sc["res"],
            "weights": "random init (no checkpoint offline), BatchNorm in eval mode",
            "l2": "value/e2e: inputs larger than L2 (rotation of 4 resident scenes, ~200 MB working set each); "
                  "roofline kernel_ms: L2 flushed (256 MiB memset) right before the convolution program of every timed step"}


# ----------------------------------------------------------------------------- CPU legs (oracle ports)
def cpu_step_seconds(model, sc, frac=1.0):
    """One pass of the CPU ports over (a spatial crop holding `frac` of) the scene: U-Net, decode, vote."""
    import torch
    from canonicalvoting_b200.minkunet import decode_heads
    from oracle import hv_oracle as O
    from oracle import sparse_oracle as SO
    coords, feats = scene_tensors(sc)
    pts, R, res = sc["points"], sc["num_rots"], np.float32(sc["res"])
    if frac < 1.0:   # crop along x: the U-Net cost is linear in the voxel count
        order = np.argsort(sc["coords"][:, 0], kind="stable")[: max(int(len(pts) * frac), 512)]
        coords, feats, pts = coords[order], feats[order], pts[order]
    t0 = time.perf_counter()
    with torch.no_grad():
        f = SO.FastCpuNet(model).forward(coords, feats)
        xyz, scale, cls, prob = decode_heads(f, NCLASSES, True)
    t1 = time.perf_counter()
    O.forward(pts, xyz.numpy(), scale.numpy(), prob.numpy(), res, R, acc64=False, threads=O.num_threads())
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, t2 - t1, len(pts)


def cpu_baseline(model, workload, sc, max_seconds):
    import torch
    from oracle import hv_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frac = 0.25
    t, tu, tv, n = cpu_step_seconds(model, sc, frac)            # calibration pass (also warms the thread pools)
    frac = float(min(1.0, max(0.05, frac * 0.5 * max_seconds / max(t, 1e-3))))
    t, tu, tv, n = cpu_step_seconds(model, sc, frac)
    scenes_per_s = (n / len(sc["points"])) / t
    return {"value": scenes_per_s, "unit": "scenes/s", "cores": cores, "kind": "port",
            "sample": "1 pass over a %.0f%% spatial crop of the %s scene (%d voxels) in %.1f s (U-Net %.1f s, vote %.2f s), "
                      "extrapolated linearly in the voxel count; oracle/sparse_oracle.py FastCpuNet (torch CPU, %d threads) + "
                      "oracle/hv_oracle.c (OpenMP, %d threads)" % (100 * frac, workload, n, t, tu, tv, cores, O.num_threads())}


def run_reference_arm(args):
    """--impl reference.  The reference has NO CPU implementation of this path (hv_cuda.cpp:26-28 rejects CPU tensors;
    MinkowskiEngine is not installable offline), so per the task contract the arm times the oracle PORTS on the host
    cores, all threads.  Each step processes a spatial crop of the scene sized so that the whole run takes ~2 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc = scene_for(args.workload, 0)
    model = make_model()
    t, _, _, _ = cpu_step_seconds(model, sc, 0.1)
    t, _, _, n = cpu_step_seconds(model, sc, 0.1)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    frac = float(min(1.0, max(0.02, 0.1 * budget / max(t, 1e-3))))
    per_step, n_used = [], 0
    for it in range(args.warmup + args.steps):
        t, _, _, n_used = cpu_step_seconds(model, sc, frac)
        if it >= args.warmup:
            per_step.append(t)
    scene_frac = n_used / len(sc["points"])
    ms_scene = 1e3 * float(np.mean(per_step)) / scene_frac
    val = 1e3 / ms_scene
    sample = "each step = a %.0f%% spatial crop (%d voxels) of the scene; scenes/s extrapolated linearly in the voxel count" % (
        100 * scene_frac, n_used)
    line = {
        "impl": "reference", "metric": "scenes_per_sec", "value": val, "unit": "scenes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_scene, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, sc),
        "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def train_c3(args, dev, rank, world, local):
    """BASELINE configs[2]/[3]: the training step (train_joint.py:244-288) on 8 synthetic 50k-voxel scenes per GPU -- MinkUNet34C
    forward + joint loss + backward + Adam, convolutions / input gradients / weight gradients on tcgen05 (tf32 mode); under
    torchrun DistributedDataParallel all-reduces the 37.9 M gradients (151 MB) over NCCL, overlapped with the backward pass.
    Returns the extra object of the bench line (scenes/s over all ranks, max over ranks)."""
    import torch
    import torch.distributed as dist
    from canonicalvoting_b200 import sparse as ME
    from canonicalvoting_b200 import synthetic, train
    from canonicalvoting_b200.minkunet import MinkUNet34C
    scenes_per_gpu, steps, warm = 8, args.train_steps, 2
    ME.set_forward_mode("tf32")
    try:
        torch.manual_seed(0)
        model = MinkUNet34C(3, 6 * NCLASSES + NCLASSES + 1).to(dev).train()
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        mine = train.shard_scenes(scenes_per_gpu * world, rank, world)
        batch = tuple(t.pin_memory() for t in train.collate([synthetic.make_scene(50000, 128, 12, seed=1000 + i) for i in mine]))
        for _ in range(warm):
            loss = train.train_step(ddp, opt, batch, dev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = train.train_step(ddp, opt, batch, dev)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t) / steps
        nparam = sum(p.numel() for p in model.parameters())
        out = {"metric": "train_scenes_per_sec", "value": scenes_per_gpu * world / sec, "unit": "scenes/s", "ms_per_step": 1e3 * sec,
               "steps": steps, "scenes_per_gpu": scenes_per_gpu, "rows_per_gpu": int(batch[0].shape[0]), "loss": float(loss),
               "grad_allreduce_bytes": 4 * nparam if world > 1 else 0, "scaling": "weak",
               "what": "config C3/C4: MinkUNet34C fwd + joint loss + bwd + Adam on 8 x 50k-voxel scenes per GPU, tf32 tcgen05 convolutions"
                       + (", DDP/NCCL gradient all-reduce" if world > 1 else "")}
        del ddp, opt, model, batch
        torch.cuda.empty_cache()
        return out
    finally:
        ME.set_forward_mode("auto")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C5"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--streams", type=int, default=3, help="scenes in flight per GPU (one CUDA-graph lane + stream each)")
    ap.add_argument("--train-steps", type=int, default=6, help="timed steps of the training workload (config C3/C4); 0 = skip")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "ours" else 1)

    if args.impl == "reference":
        if args.steps > 30:
            args.steps = 30          # the CPU arm is sized for a ~2 minute run; its per-step sample shrinks with the step count
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import hv_cuda
    from canonicalvoting_b200 import hv_cuda as H
    from canonicalvoting_b200.engine import MinkUNetEngine
    from canonicalvoting_b200.hough_voting import back_project

    if os.environ.get("CVB200_CONV_OPTS"):      # A/B switch for measurements: "<allow_split>,<launch bits>"
        from canonicalvoting_b200 import engine as _engine
        o = [int(x) for x in os.environ["CVB200_CONV_OPTS"].split(",")]
        _engine.set_conv_options(o[0], o[1])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a rank that dies takes the job down within minutes instead of leaving the others in a collective until somebody kills them
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))

    # weak scaling: every rank owns its own scenes (scene i -> rank i mod W); no data-path collective
    sc = scene_for(args.workload, seed=rank)
    n, R, res = len(sc["points"]), sc["num_rots"], sc["res"]
    model_cpu = make_model()
    model = make_model().to(dev)
    engine = MinkUNetEngine(model, NCLASSES, True)
    coords_h, feats_h = scene_tensors(sc)
    coords_h, feats_h = coords_h.pin_memory(), feats_h.pin_memory()
    coords_d, feats_d = coords_h.to(dev), feats_h.to(dev)
    res_t = torch.tensor(res, dtype=torch.float32, device=dev)
    rots_t = torch.tensor(R, dtype=torch.int32, device=dev)
    pts_d = (coords_d[:, 1:].float() * res).contiguous()
    corner, _, dims = H.grid_dims(pts_d, res)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    thresh_high = 60.0 * R / 120           # eval_joint.py:18 assumes 120 rotations

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- stage-by-stage step of ONE scene, L2 flushed: the roofline's kernel times (plain stream launches, no graph)
    def step_resident(marks=None):
        cm = engine.build_maps(coords_d)
        arr, f, keep = engine.build(coords_d, feats_d, cm)
        flush.zero_()
        if marks is not None:
            marks[0].record(stream)
        engine.execute(arr, keep)
        if marks is not None:
            marks[1].record(stream)
        xyz, scale, cls, prob, points = engine.decode(f, coords_d, res)      # + scan_points = coords * res (eval_joint.py:193)
        if marks is not None:
            marks[2].record(stream)
        out = H.forward_host(points, xyz, scale, prob, res, R, corner, dims)
        if marks is not None:
            marks[3].record(stream)
        return out

    # ---- (1) throughput: K scenes back to back over a rotation of resident scenes whose combined working set exceeds the 126 MB
    #      L2.  One scene = ONE CUDA-graph launch (map builder + convolution program + decode + vote, SceneGraph) on one of
    #      `--streams` lanes; a single host thread copies the scene into the lane's input buffers and replays.
    n_rot = 4
    scenes = [sc] + [scene_for(args.workload, seed=rank + world * (1 + j)) for j in range(n_rot - 1)]
    dev_scenes = []
    for s_ in scenes:
        c_h, f_h = scene_tensors(s_)
        c_d, f_d = c_h.to(dev), f_h.to(dev)
        p_d = (c_d[:, 1:].float() * res).contiguous()
        cr, _, dm = H.grid_dims(p_d, res)
        assert tuple(cr) == tuple(corner) and tuple(dm) == tuple(dims), "the synthetic scenes of a workload share the vote-grid geometry"
        dev_scenes.append((c_d, f_d))
    torch.cuda.synchronize()
    L = max(1, args.streams)
    vote_cfg = dict(res=res, num_rots=R, corner=corner, dims=dims)
    lanes = [engine.graph_lane(n, vote=vote_cfg) for _ in range(L)]
    lanes_e2e = [engine.graph_lane(n) for _ in range(L)]                   # network + decode; the vote goes through hv_cuda.forward
    lane_streams = [torch.cuda.Stream(dev) for _ in range(L)]
    host_s = [0.0, 0]

    def run_value(k_total):
        start = ev()
        start.record(stream)
        for st in lane_streams:
            st.wait_event(start)
        t0 = time.perf_counter()
        for j in range(k_total):
            with torch.cuda.stream(lane_streams[j % L]):
                lanes[j % L].run(*dev_scenes[j % n_rot])
        host_s[0] += time.perf_counter() - t0
        host_s[1] += k_total
        ends = []
        for st in lane_streams:
            e_ = ev()
            e_.record(st)
            ends.append(e_)
        for e_ in ends:
            e_.synchronize()
        return max(start.elapsed_time(e_) for e_ in ends) / 1e3

    run_value(args.warmup * L + n_rot)
    barrier()
    import gc
    gc.collect()
    gc.disable()
    passes = []
    with ClockSampler(local) as clk:
        t_attach = time.time()          # let nvidia-smi attach before the timed region -- with the GPU kept busy (an idle
        while time.time() - t_attach < 0.3:   # GPU drops its clocks and the first timed pass pays the ramp-up)
            run_value(4 * L)
        clk.mark()
        host_s[0], host_s[1] = 0.0, 0
        for _ in range(3):              # K steps, three times back to back; every pass is reported, the median counts
            passes.append(run_value(args.steps))
            barrier()
    gc.enable()
    t_resident = sorted(passes)[1]
    host_us_per_scene = 1e6 * host_s[0] / max(host_s[1], 1)

    # ---- (2) diagnostics for the roofline: per-step CUDA events with an L2 flush (256 MiB write) between steps
    for _ in range(3):
        step_resident()
    barrier()
    marks = []
    for _ in range(10):
        m = [ev(), ev(), ev(), ev()]
        step_resident(m)
        marks.append(m)
    barrier()
    step_ms = [m[0].elapsed_time(m[3]) for m in marks]
    unet_ms = [m[0].elapsed_time(m[1]) for m in marks]
    vote_ms = [m[2].elapsed_time(m[3]) for m in marks]

    # ---- (3) end to end through the reference-facing API with HOST buffers, returning the product's result: H2D of the scene ->
    #      network + decode (one graph launch) -> hv_cuda.forward (its grid-geometry sync included) -> candidate loop with the
    #      LCC back-projection check (eval_joint.py:195-263) -> boxes / scores / classes copied to the host.  Single host thread,
    #      software-pipelined over the lanes: the network of scene j + L - 1 is enqueued before scene j is finished.
    d2h_bytes = [0, 0]

    def e2e_enqueue(j):
        with torch.cuda.stream(lane_streams[j % L]):
            return lanes_e2e[j % L].run(coords_h, feats_h)

    def e2e_finish(j, o):
        with torch.cuda.stream(lane_streams[j % L]):
            go, gr, gs = hv_cuda.forward(o["points"], o["xyz"], o["scale"], o["prob"], res_t, rots_t)
            boxes, scores, classes = back_project(go, gr, gs, o["points"], o["xyz"], o["prob"], o["class_pred"], res, corner=corner,
                                                  thresh_high=thresh_high)
            peak = torch.stack([go.max(), go.argmax().float()])
            hb, hs, hc, hp = boxes.cpu(), scores.cpu(), classes.cpu(), peak.cpu()      # D2H of the step's result
        d2h_bytes[0] += hb.numel() * 4 + hs.numel() * 4 + hc.numel() * 8 + 8 + 8 + 36    # + box count + peak + vote-grid geometry
        d2h_bytes[1] += 1
        return len(hb)

    def run_e2e(k_total):
        start = ev()
        start.record(stream)
        for st in lane_streams:
            st.wait_event(start)
        pend = {}
        for j in range(min(L - 1, k_total)):
            pend[j] = e2e_enqueue(j)
        nb = 0
        for j in range(k_total):
            if j + L - 1 < k_total:
                pend[j + L - 1] = e2e_enqueue(j + L - 1)
            nb += e2e_finish(j, pend.pop(j))
        ends = []
        for st in lane_streams:
            e_ = ev()
            e_.record(st)
            ends.append(e_)
        for e_ in ends:
            e_.synchronize()
        return max(start.elapsed_time(e_) for e_ in ends) / 1e3, nb

    run_e2e(args.warmup * L)
    t_e2e, n_boxes = float("inf"), 0
    for _ in range(2):            # two passes of K steps, the steadier one counts
        barrier()
        t_, n_boxes = run_e2e(args.steps)
        t_e2e = min(t_e2e, t_)
        barrier()
    h2d = coords_h.numel() * 4 + feats_h.numel() * 4
    d2h = d2h_bytes[0] / max(d2h_bytes[1], 1)

    # ---- (3b) the UNCHANGED module call of eval_joint.py:169-171 -- model(ME.SparseTensor(feats, coords)).F -- timed the same way
    #      (single scene at a time: that is how the script calls it), reported beside the engine
    t_module = None
    try:
        from canonicalvoting_b200 import sparse as ME
        with torch.no_grad():
            for _ in range(3):
                model(ME.SparseTensor(feats_d, coords_d, device=dev)).F
            torch.cuda.synchronize()
            m0, m1 = ev(), ev()
            m0.record(stream)
            for _ in range(10):
                model(ME.SparseTensor(feats_d, coords_d, device=dev)).F
            m1.record(stream)
            torch.cuda.synchronize()
            t_module = m0.elapsed_time(m1) / 10
    except Exception as e:  # pragma: no cover
        t_module = repr(e)

    launches_per_scene = lanes[0].launches if lanes else 0     # kernels of one scene graph: map builder, input pad, convolutions (decode fused), vote
    lanes, lanes_e2e = [], []
    torch.cuda.empty_cache()

    # ---- (4) training workload (configs C3 / C4) on the same clock
    train = None
    if args.train_steps > 0:
        try:
            train = train_c3(args, dev, rank, world, local)
        except Exception as e:  # pragma: no cover
            train = {"error": repr(e)}

    # ---- max over ranks
    times = torch.tensor([t_resident, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_resident, t_e2e = times.tolist()

    if rank == 0:
        hbm_gbs, bf16_tf, peak_src = load_peaks()
        # algorithmic FLOPs of the convolution program: 2 * pairs * cin * cout per op, pairs counted from the tables
        arr, _, keep = engine.build(coords_d, feats_d)
        flops, executed = program_flops(engine, keep[0], arr)
        unet_med = float(np.median(unet_ms))
        vote_med = float(np.median(vote_ms))
        tflops = flops / (unet_med * 1e-3) / 1e12
        tf32_peak = bf16_tf / 2.0
        vbytes = vote_bytes(n, dims)
        vote_gbs = vbytes / (vote_med * 1e-3) / 1e9
        # the CPU ports on the host cores: at N = 1 only (the other ranks of a multi-GPU run would idle through it)
        cpu = cpu_baseline(model_cpu, args.workload, sc, args.cpu_seconds) if world == 1 else {"skipped": "reported at N = 1 only"}
        try:   # the reference's own CUDA vote kernel (unmodified, built for sm_100a) on this same GPU, reported beside it
            from oracle import build_ref
            ref = build_ref.load_ref() if world == 1 else None
            if ref is not None:
                xyz, scale, cls, prob = engine.predict(coords_d, feats_d)
                for _ in range(3):
                    ref.forward(pts_d, xyz, scale, prob, res_t, rots_t)
                torch.cuda.synchronize()
                r0, r1 = ev(), ev()
                r0.record(stream)
                for _ in range(10):
                    ref.forward(pts_d, xyz, scale, prob, res_t, rots_t)
                r1.record(stream)
                torch.cuda.synchronize()
                cpu["reference_vote_cuda_same_gpu"] = {"ms_per_scene": r0.elapsed_time(r1) / 10,
                                                       "what": "unmodified hv_cuda.forward built for sm_100a (oracle/_ref)"}
        except Exception as e:  # pragma: no cover
            cpu["reference_vote_cuda_same_gpu"] = {"error": repr(e)}
        ms_step = 1e3 * t_resident / args.steps
        line = {
            "metric": "scenes_per_sec", "value": world * args.steps / t_resident, "unit": "scenes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "passes_ms_per_step": [1e3 * p_ / args.steps for p_ in passes],
            "step_ms_flushed": {"median": float(np.median(step_ms)), "min": float(np.min(step_ms)), "max": float(np.max(step_ms))},
            "host_us_per_scene": host_us_per_scene,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (U-Net) / f32 (vote)",
            "data": "synthetic", "config": workload_config(args.workload, sc), "scenes_in_flight_per_gpu": L,
            "e2e": {"value": world * args.steps / t_e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "result": "boxes [K,8,3] + scores + classes of the candidate loop (eval_joint.py:255-263) + vote-map peak; K = %.2f boxes per "
                              "scene with these randomly initialised weights" % (n_boxes / max(args.steps, 1)),
                    "module_api_ms_per_scene": t_module,
                    "module_api_what": "the unchanged call of eval_joint.py:169-171, model(ME.SparseTensor(feats, coords)).F in eval mode under "
                                       "no_grad (one scene at a time, device-resident inputs, network only); engine, same scope: "
                                       "step_ms_flushed minus the vote"},
            "gpu_launches": launches_per_scene * args.steps,
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": tf32_peak, "unit": "TFLOP/s", "frac": tflops / tf32_peak,
                         "traffic": CONV_DRAM_BYTES_NCU.get(args.workload),
                         "traffic_source": "sum of dram__bytes_read.sum + dram__bytes_write.sum over the 63 convolution launches of one scene, "
                                           "ncu launch list of round 2 (profiles/r2g_launches_scene_C2.csv, r2g_per_layer_C2.txt)",
                         "peak_source": peak_src + ": bf16 sustained / 2 (kind::tf32 runs at half the bf16 rate)",
                         "kernel": "sc_conv_persist_kernel program of the U-Net (%d fused convolutions), one scene alone, L2 flushed" % len(arr),
                         "algorithmic_flops": flops, "executed_flops": executed, "kernel_ms": unet_med,
                         "in_flight": {"achieved": flops * world * args.steps / t_resident / 1e12 / world,
                                       "what": "algorithmic FLOPs x scenes/s per GPU over the timed region (whole step incl. map builder, "
                                               "decode and vote; %d scenes in flight): a lower bound of the program's throughput" % L},
                         "vote": {"bound": "hbm", "achieved": vote_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": vote_gbs / hbm_gbs,
                                  "kernel": "hv_scatter_kernel + hv_finalize_kernel", "algorithmic_bytes": vbytes,
                                  "kernel_ms": vote_med, "traffic": VOTE_DRAM_BYTES_NCU.get(args.workload),
                                  "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the two kernels (profiles/r2g_launches_scene_C2.csv, "
                                                    "r2g_ncu_full_vote.csv)"}},
            "cpu_baseline": cpu,
            "train_C3": train,
            "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def program_flops(engine, cm, arr):
    """(algorithmic, executed): 2 * sum over ops of (existing (output, offset) pairs) * cin * cout, and the FLOPs of the MMAs the
    kernel actually issues -- every 128-row tile runs all K^3 offsets of all its rows (missing neighbours are zero rows)."""
    import torch
    seen, total, executed = {}, 0, 0
    tables = {t.data_ptr(): t for t in cm._nbr.values()}
    for d in cm._down.values():
        tables[d["children"].data_ptr()] = d["children"]
        tables[d["up_table"].data_ptr()] = d["up_table"]
    for o in arr:
        t = tables[o.table]
        if o.table not in seen:
            seen[o.table] = int((t >= 0).sum())
        total += 2 * seen[o.table] * (4 if o.kind == 3 else o.cin) * o.cout     # kind 3: 4-channel gather (cin field = padded K)
        rows = (o.n_out + 127) // 128 * 128
        executed += 2 * rows * (o.cin if o.kind == 3 else o.k3 * o.cin) * o.cout       # kind 3: cin field = padded K = 32 * ceil(k3 / 8)
    return total, executed


if __name__ == "__main__":
    main()
