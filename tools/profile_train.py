#!/usr/bin/env python
"""tools/profile_train.py -- where the training step's GPU time goes (torch profiler, one step, config C3 on one GPU)."""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from canonicalvoting_b200 import sparse as ME  # noqa: E402
from canonicalvoting_b200 import synthetic, train  # noqa: E402
from canonicalvoting_b200.minkunet import MinkUNet34C  # noqa: E402

import os  # noqa: E402

import torch.distributed as dist  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:       # under torchrun: DistributedDataParallel, the gradient all-reduce shows up as ncclDevKernel_* rows
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ME.set_forward_mode("tf32")
torch.manual_seed(0)
model = MinkUNet34C(3, 64).to(dev).train()
ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
batch = tuple(t.pin_memory() for t in train.collate([synthetic.make_scene(50000, 128, 12, seed=8 * rank + i) for i in range(8)]))
for _ in range(2):
    train.train_step(ddp, opt, batch, dev)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    train.train_step(ddp, opt, batch, dev)
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
    nccl = [e for e in prof.key_averages() if "nccl" in e.key.lower()]
    for e in nccl:
        print("NCCL: %s  calls %d  total %.3f ms" % (e.key[:90], e.count, e.device_time_total / 1e3))
if world > 1:
    dist.destroy_process_group()
