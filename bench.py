#!/usr/bin/env python
"""bench.py -- headline benchmark of the CanonicalVoting hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C1|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one synthetic scene (BASELINE.json configs[1],
"C2": 50 000 points voting into a 128^3 grid with num_rots = 12).  Prints ONE JSON line
(rank 0).  See DESIGN.md "Measurement" for the definition of every field.

  value      scenes/s, whole job (all ranks), inputs resident in HBM, CUDA-event timed, L2 flushed
             between timed steps, max over ranks
  e2e        scenes/s through the reference-facing API (`hv_cuda.forward`) starting from pinned HOST
             buffers: H2D of the scene + geometry sync + vote + D2H of the step's result
  roofline   the vote op (scatter + write-out kernels): algorithmic bytes 40N + 24G per scene
              / measured kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port (oracle/hv_oracle.c, OpenMP, all host cores) on a bounded sample
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


# ----------------------------------------------------------------------------- helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.samples, self.stop, self.th = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.samples.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

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


def algorithmic_bytes(n, dims):
    """SURVEY.md 8d contract figure: read every input once (10 floats/point), write every output
    voxel once (6 floats/voxel)."""
    return 40 * n + 24 * int(dims[0]) * int(dims[1]) * int(dims[2])


# ----------------------------------------------------------------------------- CPU arm
def cpu_vote_scenes_per_s(workload, max_seconds, threads=None):
    """The oracle port timed on the host cores: bounded sample of the same workload."""
    from oracle import hv_oracle as O
    threads = threads or O.num_threads()
    sc = scene_for(workload, 0)
    res = np.float32(sc["res"])
    R = sc["num_rots"]
    O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], res, R, acc64=False, threads=threads)  # warm
    t0, n = time.perf_counter(), 0
    while True:
        O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], res, R, acc64=False, threads=threads)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= max_seconds or n >= 2000:
            break
    return n / dt, threads, "%d x %s scene (N=%d, R=%d) in %.1f s, oracle/hv_oracle.c OpenMP atomics" % (
        n, workload, len(sc["points"]), R, dt)


def run_reference_arm(args):
    """--impl reference.  The reference has NO CPU implementation (hv_cuda.cpp:26-28 rejects CPU
    tensors) and its CUDA build cannot run without a GPU process of its own; per the task contract
    the arm times the oracle PORT of the path on the host cores, all threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = []
    from oracle import hv_oracle as O
    threads = O.num_threads()
    sc = scene_for(args.workload, 0)
    res, R = np.float32(sc["res"]), sc["num_rots"]
    reps = {"C1": 200, "C2": 10, "C5": 2}[args.workload]   # scenes per step: bounded sample
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for _ in range(reps):
            O.forward(sc["points"], sc["xyz"], sc["scale"], sc["obj"], res, R, acc64=False, threads=threads)
        if it >= args.warmup:
            per_step.append((time.perf_counter() - t0) / reps)
    ms = 1e3 * float(np.mean(per_step))
    val = 1e3 / ms
    line = {
        "impl": "reference", "metric": "scenes_per_sec", "value": val, "unit": "scenes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, sc),
        "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": threads, "kind": "port",
                         "sample": "%d scenes per step x %d steps, oracle/hv_oracle.c OpenMP" % (reps, args.steps)},
        "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(workload, sc):
    G = sc["grid"]
    return {"workload": "%s: vote op (hv_cuda.forward) on one synthetic room scene, N=%d points, grid %d^3, "
                        "num_rots=%d; MinkUNet34C forward not yet part of the step" % (
                            workload, len(sc["points"]), G, sc["num_rots"]),
            "points": len(sc["points"]), "grid": [G, G, G], "num_rots": sc["num_rots"], "res": sc["res"],
            "l2": "flushed between timed steps (256 MiB memset)"}


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C5"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "ours" else 1)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import hv_cuda
    from canonicalvoting_b200 import hv_cuda as H

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # weak scaling: every rank owns its own scenes (scene i -> rank i mod W); no data-path collective
    sc = scene_for(args.workload, seed=rank)
    n, R, res = len(sc["points"]), sc["num_rots"], sc["res"]
    host = {k: torch.from_numpy(sc[k]).pin_memory() for k in ("points", "xyz", "scale", "obj")}
    d = {k: v.to(dev) for k, v in host.items()}
    res_t = torch.tensor(res, dtype=torch.float32, device=dev)
    rots_t = torch.tensor(R, dtype=torch.int32, device=dev)
    corner, _, dims = H.grid_dims(d["points"], res)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value, roofline): the step = vote op, async, geometry known
    def step_resident():
        return H.forward_host(d["points"], d["xyz"], d["scale"], d["obj"], res, R, corner, dims)

    for _ in range(args.warmup):
        flush.zero_()
        step_resident()
    barrier()
    evs = []
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = step_resident()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_resident = float(np.sum(step_ms)) / 1e3
    del out

    # ---- end to end through the reference-facing API with HOST buffers
    def step_e2e():
        dd = [host[k].to(dev, non_blocking=True) for k in ("points", "xyz", "scale", "obj")]
        go, gr, gs = hv_cuda.forward(dd[0], dd[1], dd[2], dd[3], res_t, rots_t)
        peak = torch.stack([go.max(), go.argmax().float()])
        return peak.cpu()          # D2H read of the step's result (peak value + voxel)

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    e1.record(stream)
    barrier()
    t_e2e = e0.elapsed_time(e1) / 1e3
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = 8

    # ---- max over ranks
    times = torch.tensor([t_resident, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_resident, t_e2e = times.tolist()

    if rank == 0:
        peak_gbs, peak_src = load_peaks()
        ms = 1e3 * t_resident / args.steps
        bytes_alg = algorithmic_bytes(n, dims)
        kern_ms = float(np.median(step_ms))
        achieved = bytes_alg / (kern_ms * 1e-3) / 1e9
        cpu_val, cpu_cores, cpu_sample = cpu_vote_scenes_per_s(args.workload, args.cpu_seconds)
        cpu = {"value": cpu_val, "unit": "scenes/s", "cores": cpu_cores, "kind": "port", "sample": cpu_sample}
        # the reference's own CUDA kernel (unmodified, built for sm_100a) on this same GPU, reported beside it
        try:
            from oracle import build_ref
            ref = build_ref.load_ref()
            if ref is not None:
                for _ in range(3):
                    ref.forward(d["points"], d["xyz"], d["scale"], d["obj"], res_t, rots_t)
                torch.cuda.synchronize()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                r0.record(stream)
                for _ in range(reps):
                    ref.forward(d["points"], d["xyz"], d["scale"], d["obj"], res_t, rots_t)
                r1.record(stream)
                torch.cuda.synchronize()
                cpu["reference_cuda_same_gpu"] = {"ms_per_scene": r0.elapsed_time(r1) / reps,
                                                  "what": "unmodified hv_cuda.forward built for sm_100a (oracle/_ref)"}
        except Exception as e:  # pragma: no cover
            cpu["reference_cuda_same_gpu"] = {"error": repr(e)}
        line = {
            "metric": "scenes_per_sec", "value": world * args.steps / t_resident, "unit": "scenes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, sc),
            "e2e": {"value": world * args.steps / t_e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": None, "peak_source": peak_src,
                         "kernel": "hv_scatter_kernel + hv_finalize_kernel (whole vote op)",
                         "algorithmic_bytes": bytes_alg, "kernel_ms": kern_ms},
            "cpu_baseline": cpu,
            "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
