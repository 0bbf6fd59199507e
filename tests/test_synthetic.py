"""The synthetic scene generator (canonicalvoting_b200/synthetic.py) is shared by the kernels, the oracles, the golden vectors
and bench.py: its contract is checked here -- determinism by seed, unique voxels, anchor voxels that make the
vote grid exactly G^3 under the reference's float32 geometry, value ranges of the per-point predictions, and the loader's
label contract (utils/dataloader.py: class 9 = background)."""
import numpy as np
import pytest

from canonicalvoting_b200 import synthetic
from oracle import hv_oracle as O


@pytest.mark.parametrize("n,G,R,seed", [(5000, 32, 4, 0), (20000, 64, 12, 3), (3000, 28, 12, 1)])
def test_scene_contract(n, G, R, seed):
    a = synthetic.make_scene(n, G, R, seed=seed)
    b = synthetic.make_scene(n, G, R, seed=seed)
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert np.array_equal(a[k], b[k]), k                                   # same seed, same scene
    c = a["coords"]
    assert c.shape == (n, 3) and c.dtype == np.int32 and c.min() == 0 and c.max() == G - 1
    assert len(np.unique(c.astype(np.int64) @ np.array([G * G, G, 1]))) == n       # one point per voxel (dataloader.py:197-204)
    assert (c == 0).all(1).any() and (c == G - 1).all(1).any()                     # anchors
    assert np.array_equal(a["points"], c.astype(np.float32) * np.float32(0.03))
    corner, dims = O.grid_dims(a["points"], np.float32(0.03))
    assert tuple(dims) == (G, G, G) and tuple(corner) == (0.0, 0.0, 0.0)            # float32 int((max-min)/res)+1 (hv_cuda_kernel.cu:131-134)
    assert a["obj"].min() >= 0 and a["obj"].max() <= 1 and (a["scale"] > 0).all() and np.abs(a["xyz"]).max() <= 1.2
    assert a["class_pred"].dtype == np.int64 and a["class_pred"].min() >= 0 and a["class_pred"].max() <= 8
    lbl = a["class_labels"]
    assert lbl.dtype == np.int32 and set(np.unique(lbl)) <= set(range(10)) and (lbl == 9).any() and (lbl < 9).any()
    obj = lbl < 9
    assert np.array_equal(a["class_pred"][obj], lbl[obj].astype(np.int64))          # object points predict their own class
    assert (a["obj"][obj] >= 0.6).all() and (a["obj"][~obj] <= 0.1).all()
    assert len(a["boxes"]) == 12 and a["num_rots"] == R and a["grid"] == G


def test_named_configs_and_uniform_control():
    assert synthetic.CONFIGS["C1"] == (5000, 32, 4) and synthetic.CONFIGS["C2"] == (50000, 128, 12) and synthetic.CONFIGS["C5"] == (200000, 256, 24)
    u = synthetic.make_scene(3000, 28, 12, seed=1, uniform=True)
    assert u["boxes"] == [] and (u["class_labels"] == 9).all() and u["obj"].max() > 0.9       # i.i.d. control: no objects
    assert synthetic.make_scene(2000, 24, 4, seed=5)["coords"].tolist() != synthetic.make_scene(2000, 24, 4, seed=6)["coords"].tolist()
