"""hough_voting -- the op names BASELINE.json's north_star uses, plus the autograd glue the
reference copy-pastes into every script (HVFunction / HoughVoting, train_joint.py:22-56).

    vote(points, xyz, scale, obj, res, num_rots)  == hv_cuda.forward with the same arguments
    vote_host(...)                                 fully asynchronous variant: python scalars + known geometry
    HVFunction / HoughVoting                       as in the reference scripts
    back_project(...)                              the candidate loop + LCC-aware back-projection check
                                                   (eval_joint.py:195-263) as one device-resident loop
"""
import torch

from . import hv_cuda
from .back_project import back_project, back_project_numpy  # noqa: F401


def vote(points, xyz, scale, obj, res, num_rots, corners=None):
    """Alias of hv_cuda.forward (houghvoting/src/hv_cuda.cpp:30-45)."""
    return hv_cuda.forward(points, xyz, scale, obj, res, num_rots, corners)


def vote_host(points, xyz, scale, obj, res, num_rots, corner, dims):
    """Sync-free vote: `res`/`num_rots` python numbers, `corner`/`dims` known on the host (e.g. from the
    integer voxel coordinates the loader already has on the CPU, eval_joint.py:182,193)."""
    return hv_cuda.forward_host(points, xyz, scale, obj, res, num_rots, corner, dims)


class HVFunction(torch.autograd.Function):
    """train_joint.py:22-37."""

    @staticmethod
    def forward(ctx, points, xyz, scale, obj, res, num_rots):
        ctx.save_for_backward(points, xyz, scale, obj, res, num_rots)
        grid_obj, grid_rot, grid_scale = hv_cuda.forward(points, xyz, scale, obj, res, num_rots)
        return grid_obj, grid_rot, grid_scale

    @staticmethod
    def backward(ctx, grad_obj, grad_rot, grad_scale):
        points, xyz, scale, obj, res, num_rots = ctx.saved_tensors
        d_xyz, d_scale, d_obj = hv_cuda.backward(grad_obj.contiguous(), points, xyz, scale, obj, res, num_rots)
        return None, d_xyz, d_scale, d_obj, None, None


class HoughVoting(torch.nn.Module):
    """train_joint.py:48-56."""

    def __init__(self, res=0.03, num_rots=120):
        super().__init__()
        self.res = torch.tensor(res, dtype=torch.float32).cuda()
        self.num_rots = torch.tensor(num_rots, dtype=torch.int32).cuda()

    def forward(self, points, xyz, scale, obj):
        return HVFunction.apply(points, xyz, scale, obj, self.res, self.num_rots)


# the detection post-process that follows back_project in the eval scripts (eval_joint.py:265-280)
from .obb import get_iou_obb, iou_matrix, nms_per_class  # noqa: E402,F401
