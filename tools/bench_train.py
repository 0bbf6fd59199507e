#!/usr/bin/env python
"""tools/bench_train.py [--scenes-per-gpu 8] [--points 50000] [--steps 5] -- training step (BASELINE configs[2]/[3]):
MinkUNet34C forward + joint loss + backward + Adam on a batch of synthetic scenes per GPU; under torchrun the
gradients are all-reduced over NCCL by DistributedDataParallel.  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from canonicalvoting_b200 import sparse as ME  # noqa: E402
from canonicalvoting_b200 import synthetic, train  # noqa: E402
from canonicalvoting_b200.minkunet import MinkUNet34C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes-per-gpu", type=int, default=8)
ap.add_argument("--points", type=int, default=50000)
ap.add_argument("--grid", type=int, default=128)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--mode", default="tf32", choices=["fp32", "tf32"])
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ME.set_forward_mode(a.mode)
torch.manual_seed(0)
model = MinkUNet34C(3, 64).to(dev).train()
ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
mine = train.shard_scenes(a.scenes_per_gpu * world, rank, world)
batch = train.collate([synthetic.make_scene(a.points, a.grid, 12, seed=i) for i in mine])
for _ in range(a.warmup):
    loss = train.train_step(ddp, opt, batch, dev)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = train.train_step(ddp, opt, batch, dev)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    sec = float(t) / a.steps
    print(json.dumps({"metric": "train_scenes_per_sec", "value": a.scenes_per_gpu * world / sec, "unit": "scenes/s", "n_gpus": world,
                      "ms_per_step": 1e3 * sec, "scenes_per_gpu": a.scenes_per_gpu, "points_per_scene": a.points,
                      "rows_per_gpu": int(batch[0].shape[0]), "conv_mode": a.mode, "loss": float(loss),
                      "what": "MinkUNet34C fwd + joint loss + bwd + Adam, DDP/NCCL grad all-reduce" if world > 1 else
                              "MinkUNet34C fwd + joint loss + bwd + Adam"}))
if world > 1:
    dist.destroy_process_group()
