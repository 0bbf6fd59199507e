"""ME.utils subset used by the reference: batched_coordinates (train_joint.py:82, eval_joint.py:64),
sparse_quantize (utils/dataloader.py:197; device='cuda' form at sunrgbd/brnetcanon.py:218) and kaiming_normal_
(utils/resnet.py:112).  Semantics from recollection of MinkowskiEngine 0.5.x [ME-recall].  CPU inputs are handled on the
host with numpy (the data loader's workers); CUDA float32 inputs (or device='cuda') are voxelised on the device by
cvb200_sc_quantize -- same result: the first point of every voxel, in input order."""
import math

import numpy as np
import torch


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """list of [Ni, 3] arrays / tensors -> int32 [sum Ni, 4] rows (batch index, x, y, z); float inputs are floored."""
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(np.asarray(c) if not isinstance(c, torch.Tensor) else c)
        if c.is_floating_point():
            c = torch.floor(c)
        c = c.to(torch.int64)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int64, device=c.device), c], 1))
    res = torch.cat(out, 0).to(dtype) if out else torch.zeros((0, 4), dtype=dtype)
    return res.to(device) if device is not None else res


def sparse_quantize(coordinates, features=None, labels=None, quantization_size=None, return_index=False,
                    return_inverse=False, device=None, **_ignored):
    """floor(coordinates / quantization_size), one representative row per occupied voxel (the first in input
    order; ME leaves the choice unspecified).  Returns what ME returns for the argument combination used by the
    reference: (unique_coords[, features][, labels][, index][, inverse])."""
    want_cuda = device is not None and torch.device(device).type == "cuda"
    if isinstance(coordinates, torch.Tensor) and (coordinates.is_cuda or want_cuda) and coordinates.dtype == torch.float32:
        return _sparse_quantize_cuda(coordinates.to(device) if want_cuda and not coordinates.is_cuda else coordinates, features, labels,
                                     quantization_size, return_index, return_inverse)
    c = np.asarray(coordinates.cpu() if isinstance(coordinates, torch.Tensor) else coordinates)
    if quantization_size is not None:
        c = np.floor(c / quantization_size)
    elif np.issubdtype(c.dtype, np.floating):
        c = np.floor(c)                          # like batched_coordinates: float inputs are floored, not truncated [ME-recall]
    c = c.astype(np.int64)
    _, index, inverse = np.unique(c, axis=0, return_index=True, return_inverse=True)
    order = np.sort(index)                      # keep input order among the representatives
    remap = np.empty(len(index), np.int64)
    remap[np.argsort(index)] = np.arange(len(index))
    outs = [torch.from_numpy(c[order].astype(np.int32))]
    if features is not None:
        outs.append(features[order])
    if labels is not None:
        outs.append(labels[order])
    if return_index:
        outs.append(torch.from_numpy(order))
    if return_inverse:
        outs.append(torch.from_numpy(remap[inverse.reshape(-1)]))
    return outs[0] if len(outs) == 1 else tuple(outs)


def _sparse_quantize_cuda(xyz, features, labels, quantization_size, return_index, return_inverse):
    """Device path of sparse_quantize: one hash-map pass (first point per voxel by atomicMin of the row index) + a scan."""
    import ctypes

    from .. import _lib
    from .coords import _ptr, _stream
    L = _lib.load()
    xyz = xyz.contiguous()
    n, dev = xyz.shape[0], xyz.device
    cap = int(L.cvb200_sc_hash_capacity(n))
    keys = torch.empty(cap, dtype=torch.int64, device=dev)
    vals = torch.empty(cap, dtype=torch.int32, device=dev)
    voxel = torch.empty((n, 4), dtype=torch.int32, device=dev)
    rep = torch.empty(n, dtype=torch.int32, device=dev)
    flag = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.cvb200_sc_quantize(_ptr(xyz), n, ctypes.c_float(float(quantization_size) if quantization_size is not None else 0.0), 0,
                                  _ptr(keys), _ptr(vals), cap, _ptr(voxel), _ptr(rep), _ptr(flag), _stream())
        _lib.check(rc, "cvb200_sc_quantize")
    index = torch.nonzero(flag, as_tuple=False).view(-1)            # ascending = input order (one host sync: the voxel count)
    outs = [voxel[index, 1:].contiguous()]
    if features is not None:
        outs.append(features[index.to(features.device)] if isinstance(features, torch.Tensor) else features[index.cpu().numpy()])
    if labels is not None:
        outs.append(labels[index.to(labels.device)] if isinstance(labels, torch.Tensor) else labels[index.cpu().numpy()])
    if return_index:
        outs.append(index)
    if return_inverse:
        excl = torch.cumsum(flag, 0) - flag
        outs.append(excl[rep.long()].long())
    return outs[0] if len(outs) == 1 else tuple(outs)


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """Kaiming normal for a sparse-convolution kernel [K^D, Cin, Cout] (fan = channels * kernel volume)."""
    if tensor.dim() == 3:
        kv, cin, cout = tensor.shape
    else:
        kv, (cin, cout) = 1, tensor.shape
    fan = cin * kv if mode == "fan_in" else cout * kv
    std = torch.nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)


def permute_kernel_offsets(state_dict, order="zyx"):
    """The one-line hook for checkpoints whose kernel-offset order differs from the one this package restates from recollection
    of MinkowskiEngine 0.5.x (offset index k = ix + K (iy + K iz): x fastest).
    Returns a copy of `state_dict` with axis 0 of every K^3-offset convolution kernel ([K^3, Cin, Cout], K^3 in {8, 27, 125})
    re-ordered from `order` to the package's order:

        order="zyx"     the checkpoint numbers offsets with z fastest (k' = iz + K (iy + K ix))
        order="mirror"  the checkpoint's offsets are point-mirrored (k' = K^3 - 1 - k)
        order=callable  perm = order(K) -> LongTensor [K^3] with new_kernel[k] = old_kernel[perm[k]]

        model.load_state_dict(permute_kernel_offsets(torch.load("pretrained/joint.pth")))      # eval_joint.py:152

    Applying "zyx" (or "mirror") twice gives the original back.  1x1x1 kernels ([Cin, Cout]) and all other entries pass through."""
    import torch
    out = {}
    for name, t in state_dict.items():
        k3 = t.shape[0] if (name.endswith(".kernel") and t.dim() == 3) else 0
        K = round(k3 ** (1.0 / 3)) if k3 else 0
        if K and K ** 3 == k3 and K > 1:
            if callable(order):
                perm = order(K)
            elif order == "zyx":
                k = torch.arange(k3)
                ix, iy, iz = k % K, (k // K) % K, k // (K * K)
                perm = iz + K * (iy + K * ix)          # our offset (ix, iy, iz) lives at index iz + K (iy + K ix) of the checkpoint
            elif order == "mirror":
                perm = torch.arange(k3 - 1, -1, -1)
            else:
                raise ValueError("order must be 'zyx', 'mirror' or a callable K -> permutation")
            out[name] = t[perm.to(t.device)].clone()
        else:
            out[name] = t
    return out
