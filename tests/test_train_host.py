"""CPU tests of the host side of the data-parallel training step (canonicalvoting_b200/train.py): scene sharding,
collation contract, the joint loss against a literal restatement of train_joint.py:253-283, and a world_size-2
gloo run of the gradient-averaging semantics the NCCL path relies on."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from canonicalvoting_b200 import synthetic, train


def test_shard_scenes_partitions_everything_once():
    for world in (1, 2, 3, 8):
        got = sorted(sum((train.shard_scenes(17, r, world) for r in range(world)), []))
        assert got == list(range(17))


def test_collate_matches_the_loader_contract():
    scenes = [synthetic.make_scene(300, 16, 4, seed=s) for s in (0, 1)]
    coords, feats, xyz, scale, cls = train.collate(scenes)
    assert coords.dtype == torch.int32 and coords.shape == (600, 4)
    assert coords[:300, 0].eq(0).all() and coords[300:, 0].eq(1).all()
    assert torch.equal(coords[:300, 1:], torch.from_numpy(scenes[0]["coords"]))
    assert feats.shape == (600, 3) and xyz.shape == (600, 3) and scale.shape == (600, 3) and cls.dtype == torch.int64


def test_joint_loss_matches_script_formulation():
    g = torch.Generator().manual_seed(0)
    n, nc = 200, 9
    out = torch.randn(n, 64, generator=g, requires_grad=True)
    xyz_l, scale_l = torch.randn(n, 3, generator=g), torch.rand(n, 3, generator=g) + 0.1
    cls_l = torch.randint(0, 10, (n,), generator=g)
    got = train.joint_loss(out, xyz_l, scale_l, cls_l)
    # literal restatement of train_joint.py:253-283 with xyz_weights = 1, factors = 1, log_scale = True
    idx = cls_l.clone().unsqueeze(-1).unsqueeze(-1).expand(-1, -1, 3).clone()
    idx[idx == nc] = 0
    oxyz = torch.gather(out[:, :27].reshape(-1, nc, 3), 1, idx)[:, 0]
    oscale = torch.gather(out[:, 27:54].reshape(-1, nc, 3), 1, idx)[:, 0]
    mask = (cls_l < nc) & (0 <= cls_l)
    want = torch.mean((oscale[mask] - torch.log(scale_l[mask])) ** 2) + torch.mean((oxyz[mask] - xyz_l[mask]) ** 2) + \
        torch.nn.functional.cross_entropy(out[:, 54:], cls_l)
    torch.testing.assert_close(got, want)
    got.backward()
    assert torch.isfinite(out.grad).all()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Linear(64, 64)             # stands in for the U-Net: DDP only sees parameters and gradients
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    scenes = train.shard_scenes(4, rank, world)
    g = torch.Generator().manual_seed(100)
    data = [(torch.randn(50, 64, generator=g), torch.randn(50, 3, generator=g), torch.rand(50, 3, generator=g) + 0.1,
             torch.randint(0, 10, (50,), generator=g)) for _ in range(4)]
    x = torch.cat([data[i][0] for i in scenes]); xl = torch.cat([data[i][1] for i in scenes])
    sl = torch.cat([data[i][2] for i in scenes]); cl = torch.cat([data[i][3] for i in scenes])
    loss = train.joint_loss(ddp(x), xl, sl, cl)
    loss.backward()
    q.put((rank, model.weight.grad.clone().numpy(), float(loss)))
    dist.destroy_process_group()


def test_world_size_2_gradients_are_the_mean_of_the_per_rank_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=0, atol=0)      # all ranks hold the same averaged gradient
    # single-process reference: average of the two per-rank gradients
    torch.manual_seed(0)
    model = torch.nn.Linear(64, 64)
    g = torch.Generator().manual_seed(100)
    data = [(torch.randn(50, 64, generator=g), torch.randn(50, 3, generator=g), torch.rand(50, 3, generator=g) + 0.1,
             torch.randint(0, 10, (50,), generator=g)) for _ in range(4)]
    grads = []
    for r in range(2):
        model.zero_grad()
        ids = train.shard_scenes(4, r, 2)
        x = torch.cat([data[i][0] for i in ids]); xl = torch.cat([data[i][1] for i in ids])
        sl = torch.cat([data[i][2] for i in ids]); cl = torch.cat([data[i][3] for i in ids])
        train.joint_loss(model(x), xl, sl, cl).backward()
        grads.append(model.weight.grad.clone())
    np.testing.assert_allclose(res[0][1], ((grads[0] + grads[1]) / 2).numpy(), rtol=1e-5, atol=1e-6)
