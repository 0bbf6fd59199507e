#!/usr/bin/env python
"""tools/profile_train.py -- where the training step's GPU time goes (torch profiler, one step, config C3 on one GPU)."""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from canonicalvoting_b200 import sparse as ME  # noqa: E402
from canonicalvoting_b200 import synthetic, train  # noqa: E402
from canonicalvoting_b200.minkunet import MinkUNet34C  # noqa: E402

dev = torch.device("cuda", 0)
ME.set_forward_mode("tf32")
torch.manual_seed(0)
model = MinkUNet34C(3, 64).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
batch = train.collate([synthetic.make_scene(50000, 128, 12, seed=i) for i in range(8)])
for _ in range(2):
    train.train_step(model, opt, batch, dev)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    train.train_step(model, opt, batch, dev)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
