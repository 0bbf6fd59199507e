#!/usr/bin/env python
"""tools/eval_sweep.py [--workload C5] [--scenes 16] -- BASELINE.json configs[4]: the eval_joint.py scene loop over a set of
synthetic scenes, sharded over the ranks of a torchrun launch (scene i -> rank i mod world):

    python tools/eval_sweep.py --workload C2 --scenes 8
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/eval_sweep.py --workload C5

per scene: MinkUNet34C engine + head decode (random-init weights: its outputs are timed, not used) -> Hough voting on the
scene's synthetic per-point predictions -> candidate loop with the LCC back-projection check -> per-class OBB NMS ->
detection tuples; then ONE all_gather_object of the detection lists and the detection metric against the planted boxes
(canonicalvoting_b200/evaluate.py).  Rank 0 prints one JSON line: scenes/s (CUDA events around each rank's loop, max over
ranks), mAP / AR at IoU 0.25 and 0.5.  Written at the end of round 1 after the GPU budget was spent: not yet run on a GPU.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import hough_voting  # noqa: E402
import hv_cuda  # noqa: E402
from canonicalvoting_b200 import evaluate, train  # noqa: E402
from canonicalvoting_b200.engine import MinkUNetEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C5", choices=["C1", "C2", "C5"])
    ap.add_argument("--scenes", type=int, default=16)
    ap.add_argument("--no-unet", action="store_true", help="skip the (random-weight) U-Net: vote + loop + NMS only")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    mine = train.shard_scenes(args.scenes, rank, world)
    scenes = {i: bench.scene_for(args.workload, seed=i) for i in mine}
    eng = None if args.no_unet else MinkUNetEngine(bench.make_model().to(dev), 9, True)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def run_scene(sc):
        R, res = sc["num_rots"], sc["res"]
        if eng is not None:
            c_h, f_h = bench.scene_tensors(sc)
            eng.predict(c_h.to(dev), f_h.to(dev))
        pts, xyz, scale, obj, cls = d(sc["points"]), d(sc["xyz"]), d(sc["scale"]), d(sc["obj"]), d(sc["class_pred"])
        res_t = torch.tensor(res, dtype=torch.float32, device=dev)
        rots_t = torch.tensor(R, dtype=torch.int32, device=dev)
        go, gr, gs = hv_cuda.forward(pts, xyz, scale, obj, res_t, rots_t)
        boxes, scores, classes = hough_voting.back_project(go, gr, gs, pts, xyz, obj, cls, res, thresh_high=60.0 * R / 120)
        keep = hough_voting.nms_per_class(boxes, scores, classes, 9, 0.3)
        return evaluate.scene_detections(boxes, scores, classes, keep)

    if mine:
        run_scene(scenes[mine[0]])                                  # warm-up (allocations, lazy initialisation)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pred = {"scene%04d" % i: run_scene(scenes[i]) for i in mine}
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gt = {"scene%04d" % i: [(evaluate.CATEGORIES[k], evaluate.gt_box(c[0], c[1], c[2], yaw, h[0], h[1], h[2]))
                            for c, h, yaw, k in scenes[i]["boxes"]] for i in mine}
    pred_all, gt_all = evaluate.gather_detections(pred, gt)
    if rank == 0:
        out = {"workload": args.workload, "scenes": args.scenes, "n_gpus": world, "unet": eng is not None,
               "seconds": float(t.item()), "scenes_per_sec": args.scenes / float(t.item()),
               "detections": sum(len(v) for v in pred_all.values()), "gt_boxes": sum(len(v) for v in gt_all.values())}
        for thr in (0.25, 0.5):
            ret = evaluate.compute_map(pred_all, gt_all, thr)
            out["mAP@%.2f" % thr], out["AR@%.2f" % thr] = float(ret["mAP"]), float(ret["AR"])
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
