"""Seeded synthetic scenes with the reference loader's output contract.

The reference's datasets (ScanNet / Scan2CAD, utils/dataloader.py:89-210) are not
available offline; this generator emits the same 6-tuple contract
(utils/dataloader.py:202-208: integer voxel coords, rgb feats, LCC xyz labels, scale
labels, class labels) plus per-point "network predictions" shaped like the head decode
output (eval_joint.py:173-190), so that the vote op, the candidate loop and the sparse
U-Net can be driven by identical inputs in the CUDA path, the oracle and the benchmark.

Scene: a room with floor y=0 and walls x=0, z=0, plus `n_objects`
oriented boxes resting on the floor; surface points are floored to the integer lattice
[0, G-1]^3, de-duplicated, shuffled and cut/padded to exactly N rows including the two
anchor voxels (0,0,0) and (G-1,G-1,G-1), so that the vote grid is exactly G^3
(points = coords * 0.03f lie on the 0.03 m lattice, eval_joint.py:182,193).
"""
import numpy as np

RES = np.float32(0.03)  # config/config.yaml:10 scannet_res

# BASELINE.json configs: name -> (N points, grid G, num_rots)
CONFIGS = {
    "C1": (5_000, 32, 4),
    "C2": (50_000, 128, 12),
    "C5": (200_000, 256, 24),
}


def _rot_y(yaw):
    """Rm = [[c,0,-s],[0,1,0],[s,0,c]] (eval_joint.py:215)."""
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]], dtype=np.float64)


def _box_surface(rng, centre, half, yaw, n):
    """n points uniform on the surface of an oriented box (voxel units)."""
    areas = np.array([half[1] * half[2], half[1] * half[2], half[0] * half[2], half[0] * half[2],
                      half[0] * half[1], half[0] * half[1]])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    u = rng.uniform(-1.0, 1.0, size=(n, 3))
    axis = face // 2
    sign = np.where(face % 2 == 0, -1.0, 1.0)
    u[np.arange(n), axis] = sign
    local = u * half
    return centre + local @ _rot_y(yaw).T, u


def make_scene(n_points=50_000, grid=128, num_rots=12, seed=0, n_objects=12, snap_yaw=False,
               uniform=False):
    """Returns a dict of numpy arrays (float32 unless noted):
        coords int32 [N,3], points [N,3] = coords*0.03f, feats [N,3] rgb in [0,1],
        xyz [N,3], scale [N,3], obj [N], class_pred int64 [N] in [0,8],
        xyz_labels, scale_labels, class_labels int32 (9 = background)  (loader contract),
        boxes: list of (centre_m[3], half_m[3], yaw, class)
    `uniform=True` is the no-contention control distribution of SURVEY.md 8d (all fields i.i.d.).
    """
    rng = np.random.default_rng(seed)
    G, N = int(grid), int(n_points)
    assert N >= 2 and N <= G ** 3
    boxes = []
    if uniform:
        lin = rng.choice(G ** 3 - 2, size=N - 2, replace=False) + 1
        coords = np.stack(np.unravel_index(lin, (G, G, G)), -1).astype(np.int64)
        owner = np.full(N - 2, -1)
        local = np.zeros((N - 2, 3))
    else:
        pts, own, loc = [], [], []
        budget = max(N * 3, 4096)
        # room shell: floor y=0 and walls x=0, z=0
        for axis in (1, 0, 2):
            p = rng.uniform(0, G, size=(budget // 4, 3))
            p[:, axis] = 0.0
            pts.append(p); own.append(np.full(len(p), -1)); loc.append(np.zeros((len(p), 3)))
        for k in range(n_objects):
            half = rng.uniform(0.08, 0.22, size=3) * G / 2.0
            yaw = rng.uniform(0.0, 2.0 * np.pi)
            if snap_yaw:
                yaw = np.round(yaw / (2.0 * np.pi / num_rots)) * (2.0 * np.pi / num_rots)
            r = float(np.hypot(half[0], half[2]))
            cx, cz = rng.uniform(r + 1, max(G - r - 1, r + 2), size=2)
            centre = np.array([cx, half[1], cz])
            cls = int(rng.integers(0, 9))
            boxes.append((centre, half, yaw, cls))
            p, u = _box_surface(rng, centre, half, yaw, budget // (2 * n_objects))
            pts.append(p); own.append(np.full(len(p), k)); loc.append(u)
        allp = np.concatenate(pts)
        owner = np.concatenate(own)
        local = np.concatenate(loc)
        coords = np.floor(allp).astype(np.int64)
        ok = np.all((coords >= 0) & (coords <= G - 1), axis=1)
        coords, owner, local = coords[ok], owner[ok], local[ok]
        # drop the anchors (re-added below) and de-duplicate voxels (loader: dataloader.py:197-204)
        lin = (coords[:, 0] * G + coords[:, 1]) * G + coords[:, 2]
        keep = (lin != 0) & (lin != G ** 3 - 1)
        coords, owner, local, lin = coords[keep], owner[keep], local[keep], lin[keep]
        perm = rng.permutation(len(lin))
        coords, owner, local, lin = coords[perm], owner[perm], local[perm], lin[perm]
        _, first = np.unique(lin, return_index=True)
        first = rng.permutation(first)[: N - 2]
        coords, owner, local, lin = coords[first], owner[first], local[first], lin[first]
        if len(coords) < N - 2:  # small grids: pad with random clutter voxels
            need = N - 2 - len(coords)
            taken = set(lin.tolist()) | {0, G ** 3 - 1}
            extra = []
            while len(extra) < need:
                cand = rng.integers(1, G ** 3 - 1, size=2 * need)
                for c in cand.tolist():
                    if c not in taken:
                        taken.add(c); extra.append(c)
                        if len(extra) == need:
                            break
            ec = np.stack(np.unravel_index(np.array(extra), (G, G, G)), -1)
            coords = np.concatenate([coords, ec])
            owner = np.concatenate([owner, np.full(need, -1)])
            local = np.concatenate([local, np.zeros((need, 3))])
    anchors = np.array([[0, 0, 0], [G - 1, G - 1, G - 1]], dtype=np.int64)
    coords = np.concatenate([coords, anchors])
    owner = np.concatenate([owner, [-1, -1]])
    local = np.concatenate([local, np.zeros((2, 3))])
    perm = rng.permutation(N)
    coords, owner, local = coords[perm], owner[perm], local[perm]

    n = N
    xyz = rng.uniform(-1.0, 1.0, size=(n, 3))
    scale = 0.3 * np.exp(rng.normal(0.0, 0.3, size=(n, 3)))
    obj = rng.uniform(0.0, 0.1, size=n)
    cls_pred = rng.integers(0, 9, size=n)
    cls_lbl = np.full(n, 9, dtype=np.int32)
    xyz_lbl = np.zeros((n, 3))
    scale_lbl = np.ones((n, 3))
    for k, (centre, half, yaw, cls) in enumerate(boxes):
        m = owner == k
        cnt = int(m.sum())
        if cnt == 0:
            continue
        half_m = half * float(RES)
        # LCC of the voxel centre actually emitted: Rm^T (p - c) / s  (see hv_cuda_kernel.cu:38-39)
        lcc = ((coords[m] - centre) * float(RES)) @ _rot_y(yaw) / half_m
        xyz_lbl[m] = lcc
        scale_lbl[m] = half_m
        cls_lbl[m] = cls
        xyz[m] = np.clip(lcc + rng.normal(0.0, 0.05, size=(cnt, 3)), -1.2, 1.2)
        scale[m] = half_m * np.exp(rng.normal(0.0, 0.05, size=(cnt, 3)))
        obj[m] = rng.uniform(0.6, 1.0, size=cnt)
        cls_pred[m] = cls
    if uniform:
        obj = rng.uniform(0.0, 1.0, size=n)
    coords32 = coords.astype(np.int32)
    return {
        "coords": coords32,
        "points": coords32.astype(np.float32) * RES,
        "feats": rng.uniform(0.0, 1.0, size=(n, 3)).astype(np.float32),
        "xyz": xyz.astype(np.float32),
        "scale": scale.astype(np.float32),
        "obj": obj.astype(np.float32),
        "class_pred": cls_pred.astype(np.int64),
        "xyz_labels": xyz_lbl.astype(np.float32),
        "scale_labels": scale_lbl.astype(np.float32),
        "class_labels": cls_lbl,
        "boxes": [(c * float(RES), h * float(RES), float(y), int(k)) for (c, h, y, k) in boxes],
        "grid": G, "num_rots": int(num_rots), "res": float(RES),
    }


def make_config(name, seed=0, **kw):
    n, g, r = CONFIGS[name]
    return make_scene(n, g, r, seed=seed, **kw)
