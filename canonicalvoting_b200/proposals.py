"""Vote-map proposal sampler -- `HoughVotingModule` of the SUN RGB-D variant (sunrgbd/brnetcanon.py:94-162).

The reference votes with the 7-argument `hv_cuda.forward(..., corners)` (:99), projects the vote map along the height
axis (max + argmax, :120-122), draws 1.5 x num_proposal cells with `torch.multinomial` from the max-projection raised
to `pow` (:133) and keeps the draws that lie within 0.3 m of a VoteNet seed (:139-149), repeating until num_proposal
locations are collected.  Here the projection is one pass over the grid (cvb200_hv_project_y) and a whole trial after
the draw -- unravel, look-ups, world location, nearest-seed distance, rejection, ordered compaction, append -- is one
launch (cvb200_hv_proposals); the draw itself stays `torch.multinomial` so that the random stream is the reference's.
One host synchronisation per trial (the running count) plus the reference's own `if` on the distribution.

CUDA tensors only; there is no CPU fallback.
"""
import torch
from torch import nn

from . import _lib, hv_cuda
from .hv_cuda import _ptr, _stream_ptr

SEED_RADIUS = 0.3                     # sunrgbd/brnetcanon.py:142-149


def project_y(hv_map):
    """(hv_map.max(1)[0], torch.argmax(hv_map, 1)) of a [X,Y,Z] float32 CUDA grid; argmax as int32."""
    if not hv_map.is_cuda:
        raise RuntimeError("project_y: CUDA tensor expected (there is no CPU path)")
    L = _lib.load()
    g = hv_map.to(torch.float32).contiguous()
    X, Y, Z = (int(d) for d in g.shape)
    vmax = torch.empty((X, Z), dtype=torch.float32, device=g.device)
    arg = torch.empty((X, Z), dtype=torch.int32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(L.cvb200_hv_project_y(_ptr(g), _lib.i3((X, Y, Z)), _ptr(vmax), _ptr(arg), _stream_ptr()), "cvb200_hv_project_y")
    return vmax, arg


class _Trial:
    """Device state of one forward() call: output rows, running count, scratch."""

    def __init__(self, num_proposal, n_draw, device):
        self.loc = torch.zeros((num_proposal, 3), dtype=torch.float32, device=device)
        self.scale = torch.zeros((num_proposal, 3), dtype=torch.float32, device=device)
        self.count = torch.zeros(1, dtype=torch.int32, device=device)
        self.work = torch.empty(max(int(_lib.load().cvb200_hv_proposals_work_bytes(n_draw)), 1), dtype=torch.uint8, device=device)


def append_proposals(state, sample, arg_y, hv_scale, res, corner, vote_points, radius=SEED_RADIUS):
    """One rejection trial (sunrgbd/brnetcanon.py:134-152) appended to `state`; asynchronous."""
    L = _lib.load()
    X, Y, Z = (int(d) for d in hv_scale.shape[:3])
    sample = sample.to(torch.int64).contiguous()
    seeds = vote_points.to(torch.float32).contiguous()
    with torch.cuda.device(sample.device):
        rc = L.cvb200_hv_proposals(_ptr(sample), int(sample.numel()), _ptr(arg_y), _ptr(hv_scale), _lib.i3((X, Y, Z)), float(res),
                                   _lib.f3(corner), _ptr(seeds), int(seeds.shape[0]), float(radius), int(state.loc.shape[0]),
                                   _ptr(state.loc), _ptr(state.scale), _ptr(state.count), _ptr(state.work), state.work.numel(),
                                   _stream_ptr())
        _lib.check(rc, "cvb200_hv_proposals")


class HoughVotingModule(nn.Module):
    """sunrgbd/brnetcanon.py:104-162, same constructor and forward signature (plus an optional `sampler`, the draw
    function `(dist, n) -> int64 cell indices`, for reproducible tests; default torch.multinomial with replacement)."""

    def __init__(self, res=0.03, num_rots=36, nms_size=0.15, thresh=0, num_proposal=256, no_grad=True):
        super().__init__()
        self.res = torch.tensor(res, dtype=torch.float32, device="cuda")
        self.num_rots = torch.tensor(num_rots, dtype=torch.int32, device="cuda")
        self.no_grad = no_grad
        self.nms_size_grid = int(nms_size // res)
        self.num_proposal = num_proposal
        self.thresh = thresh

    def forward(self, pc, xyz, scale, prob, corners, vote_points, pow=0.5, sampler=None):
        with torch.no_grad():       # the reference's 7-input HVFunction has no backward (:94-102)
            hv_map, _, hv_scale = hv_cuda.forward(pc, xyz.contiguous(), scale.contiguous(), prob.contiguous(), self.res, self.num_rots,
                                                  corners)
            hv_map_y, arg_y = project_y(hv_map)
            dist = torch.pow(hv_map_y + 1e-7, pow).reshape(-1)                                    # :120-121,124
            if bool(((~torch.isfinite(dist)).any() | (dist.sum() < 1e-7)).item()):               # :125-126
                dist = torch.ones_like(dist)
            n_draw = int(self.num_proposal * 1.5)
            state = _Trial(self.num_proposal, n_draw, hv_map.device)
            res_h = float(hv_cuda._host_scalar(self.res))
            corner = tuple(float(v) for v in corners[0].detach().float().cpu())
            hv_scale = hv_scale.to(torch.float32).contiguous()
            cnt = 0
            while cnt < self.num_proposal:                                                        # :131-152
                sample = sampler(dist, n_draw) if sampler is not None else torch.multinomial(dist, n_draw, replacement=True)
                append_proposals(state, sample, arg_y, hv_scale, res_h, corner, vote_points)
                cnt = int(state.count.item())
            candidates, scales = state.loc, state.scale
            probs = torch.zeros_like(candidates)[..., 0]                                          # :161
        return candidates, probs, scales
