"""oracle/proposals.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the vote-map proposal sampler, sunrgbd/brnetcanon.py:104-162 (`HoughVotingModule.forward` after the
vote): `sample_torch` is the script's own torch op sequence on CPU tensors with the random draw injected (`draws`: the
int64 cell indices torch.multinomial would return, one tensor per trial); `sample_numpy` is an explicit float32
restatement with the distance written out ((a-b)^2 summed, then sqrt) instead of torch.cdist.  The two agree exactly
away from the 0.3 m rejection boundary (tests/test_oracle_proposals.py).  No reference test exists for this module and
mmdet3d/BRNet (its caller) is absent; both functions are pinned by golden vectors from the reference module itself
(tools/make_ref_python_golden.py executes sunrgbd/brnetcanon.py:119-161 verbatim with recorded multinomial draws ->
tests/golden/refpy_proposals.npz, tests/test_oracle_refpy.py).
"""
import numpy as np
import torch


def unravel_index(index, shape):
    """sunrgbd/brnetcanon.py:85-90."""
    out = []
    for dim in reversed(shape):
        out.append(index % dim)
        index = index // dim
    return tuple(reversed(out))


def distribution(hv_map, pow=0.5):
    """:120-126 -> (dist [X*Z], hv_map_yidx [X,Z])."""
    hv_map_y = hv_map.max(1)[0] + 1e-7
    hv_map_y = torch.pow(hv_map_y, pow)
    hv_map_yidx = torch.argmax(hv_map, 1)
    dist = hv_map_y.reshape(-1)
    if (not torch.all(torch.isfinite(dist))) or (dist.sum() < 1e-7):
        dist = torch.ones_like(dist)
    return dist, hv_map_yidx


def sample_torch(hv_map, hv_scale, res, corner0, vote_points, draws, num_proposal, pow=0.5):
    """:118-161 with torch.multinomial replaced by the given draws.  Returns (candidates, scales, trials used)."""
    res = torch.tensor(res, dtype=torch.float32)
    dist, hv_map_yidx = distribution(hv_map, pow)
    shape_xz = (hv_map.shape[0], hv_map.shape[2])
    cnt, loc, scales, used = 0, [], [], 0
    while cnt < num_proposal:
        sample = draws[used]
        used += 1
        sample_idx = unravel_index(sample, shape_xz)
        world_loc = torch.stack([sample_idx[0], hv_map_yidx[sample_idx[0], sample_idx[1]], sample_idx[1]], -1) * res + corner0
        scale = hv_scale[sample_idx[0], hv_map_yidx[sample_idx[0], sample_idx[1]], sample_idx[1], :]
        dist2seed = torch.min(torch.cdist(world_loc, vote_points), -1)[0]
        if torch.sum(dist2seed < 0.3) == 0:
            loc.append(world_loc)
            scales.append(scale)
        else:
            loc.append(world_loc[dist2seed < 0.3])
            scales.append(scale[dist2seed < 0.3])
        cnt += loc[-1].shape[0]
    return torch.cat(loc)[:num_proposal], torch.cat(scales)[:num_proposal], used


def project_y_numpy(hv_map):
    """(max over y, first arg-max over y) of a [X,Y,Z] array."""
    return hv_map.max(1), hv_map.argmax(1).astype(np.int32)


def sample_numpy(hv_map, hv_scale, res, corner0, vote_points, draws, num_proposal, radius=0.3):
    """Explicit float32 restatement of one forward() given the draws; distance = sqrt(dx*dx + dy*dy + dz*dz) in float32,
    left-to-right.  Returns (candidates [P,3], scales [P,3], trials used, min distance of every draw per trial)."""
    hv_map = np.asarray(hv_map, np.float32)
    hv_scale = np.asarray(hv_scale, np.float32)
    seeds = np.asarray(vote_points, np.float32)
    res = np.float32(res)
    corner0 = np.asarray(corner0, np.float32)
    _, arg = project_y_numpy(hv_map)
    Z = hv_map.shape[2]
    cnt, loc, scales, used, dmins = 0, [], [], 0, []
    while cnt < num_proposal:
        s = np.asarray(draws[used], np.int64)
        used += 1
        ix, iz = s // Z, s % Z
        iy = arg[ix, iz].astype(np.int64)
        w = (np.stack([ix, iy, iz], -1).astype(np.float32) * res).astype(np.float32) + corner0
        sc = hv_scale[ix, iy, iz, :]
        d = (w[:, None, :] - seeds[None, :, :]).astype(np.float32)
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]
        dmin = np.sqrt(d2.min(1) if seeds.shape[0] else np.full(len(s), np.inf, np.float32)).astype(np.float32)
        dmins.append(dmin)
        keep = dmin < np.float32(radius)
        if keep.sum() == 0:
            keep[:] = True
        loc.append(w[keep])
        scales.append(sc[keep])
        cnt += int(keep.sum())
    return np.concatenate(loc)[:num_proposal], np.concatenate(scales)[:num_proposal], used, dmins
