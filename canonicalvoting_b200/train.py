"""Training step of the joint model (train_joint.py:244-288) on canonicalvoting_b200.sparse, data-parallel over scenes.

The reference trains on one GPU (`model.cuda()`, train_joint.py:230).  Here scenes shard across ranks (one process
per GPU): every rank builds ONE sparse tensor from its own scenes (batch index = coordinate column 0, exactly like
`ME.utils.batched_coordinates`, train_joint.py:82), runs forward / losses / backward locally, and gradients are
averaged with NCCL all-reduce (torch DistributedDataParallel buckets, overlapped with the backward pass).  BatchNorm
statistics stay per rank (the reference's BN also sees only its own 3-scene batch; SyncBN would change semantics).
Masked-mean losses make "mean of per-rank means" differ slightly from a global mean when object-point counts differ
between ranks -- the usual DDP semantics.
"""
import torch
import torch.nn.functional as F

from . import sparse as ME

NCLASSES = 9


def collate(scenes):
    """List of synthetic scenes (canonicalvoting_b200.synthetic) -> the 6-tuple of train_joint.py's collate_fn (:78-90)."""
    coords = ME.utils.batched_coordinates([torch.from_numpy(s["coords"]) for s in scenes])
    cat = lambda k, dt: torch.cat([torch.from_numpy(s[k]) for s in scenes]).to(dt)
    return (coords, cat("feats", torch.float32), cat("xyz_labels", torch.float32), cat("scale_labels", torch.float32),
            cat("class_labels", torch.int64))


def joint_loss(out_f, xyz_labels, scale_labels, class_labels, log_scale=True, xyz_factor=1.0, scale_factor=1.0, xyz_weights=None):
    """train_joint.py:253-283: per-class xyz / scale heads gathered by the GT class, masked MSE x2 + 10-way CE.  `xyz_weights`
    ([3], config xyz_component_weights, train_joint.py:241,271-272) weighs the three components of both regression terms; the
    cross-entropy term is added only when the batch holds an object point, like the script (:269-273)."""
    nc = NCLASSES
    idx = class_labels.clone()
    idx[(idx < 0) | (idx == nc)] = 0
    idx = idx.view(-1, 1, 1).expand(-1, 1, 3)
    xyz = torch.gather(out_f[:, :3 * nc].reshape(-1, nc, 3), 1, idx)[:, 0]
    scale = torch.gather(out_f[:, 3 * nc:6 * nc].reshape(-1, nc, 3), 1, idx)[:, 0]
    logits = out_f[:, 6 * nc:]
    mask = (class_labels < nc) & (class_labels >= 0)
    loss = out_f.sum() * 0.0
    if bool(mask.any()):
        w = 1.0 if xyz_weights is None else torch.as_tensor(xyz_weights, dtype=out_f.dtype, device=out_f.device).view(1, 3)
        tgt = torch.log(scale_labels[mask]) if log_scale else scale_labels[mask]
        loss = loss + scale_factor * torch.mean((scale[mask] - tgt) ** 2 * w) + xyz_factor * torch.mean((xyz[mask] - xyz_labels[mask]) ** 2 * w)
        loss = loss + F.cross_entropy(logits, class_labels)
    return loss


def train_step(model, optimizer, batch, device):
    """One optimisation step on this rank's batch; `model` may be wrapped in DistributedDataParallel."""
    coords, feats, xyz_l, scale_l, class_l = batch
    optimizer.zero_grad(set_to_none=True)
    nb = dict(non_blocking=True)          # pinned batches (a DataLoader with pin_memory) upload asynchronously
    feats = feats.to(device, **nb).clone()
    feats[:, -3:] = feats[:, -3:] * 2.0 - 1.0                                  # train_joint.py:248-249: only the rgb columns are recentred
    out = model(ME.SparseTensor(feats, coords.to(device, **nb), device=device))
    loss = joint_loss(out.F, xyz_l.to(device, **nb), scale_l.to(device, **nb), class_l.to(device, **nb))
    loss.backward()
    optimizer.step()
    return loss.detach()


def shard_scenes(n_scenes, rank, world):
    """Scene i -> rank i mod world."""
    return [i for i in range(n_scenes) if i % world == rank]


# ---- schedules of the training script (train_joint.py:100-138,200-206,224-225; config/config.yaml:30-36)
def learning_rate(epoch, base=1e-3, decay_steps=(80, 120, 160), decay_rates=(0.1, 0.1, 0.1)):
    """get_current_lr (train_joint.py:128-133): the base rate times every decay factor whose epoch has been reached."""
    lr = base
    for step, rate in zip(decay_steps, decay_rates):
        if epoch >= step:
            lr *= rate
    return lr


def adjust_learning_rate(optimizer, epoch, **schedule):
    """train_joint.py:135-138."""
    lr = learning_rate(epoch, **schedule)
    for group in optimizer.param_groups:
        group["lr"] = lr
    return lr


def bn_momentum(epoch, init=0.5, floor=0.001, decay_step=20, decay_rate=0.5):
    """The `bn_lbmd` lambda of train_joint.py:224: init * rate^(epoch // step), not below the floor."""
    return max(init * decay_rate ** int(epoch / decay_step), floor)


class BNMomentumScheduler:
    """train_joint.py:100-125.  Like the reference's setter (:92-97) it writes `momentum` on the MinkowskiBatchNorm WRAPPER
    modules; the wrapped `bn` (nn.BatchNorm1d) keeps its own momentum of 0.1, which is what the reference observably trains
    with."""

    def __init__(self, model, bn_lambda=bn_momentum, last_epoch=-1):
        if not isinstance(model, torch.nn.Module):
            raise RuntimeError("Class '%s' is not a PyTorch nn Module" % type(model).__name__)
        self.model, self.lmbd = model, bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        epoch = self.last_epoch + 1 if epoch is None else epoch
        self.last_epoch = epoch
        value = self.lmbd(epoch)
        for m in self.model.modules():
            if isinstance(m, ME.MinkowskiBatchNorm):
                m.momentum = value
