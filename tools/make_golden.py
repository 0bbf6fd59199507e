"""tools/make_golden.py -- generate tests/golden/hv_*.npz from the UNMODIFIED reference kernel.

Runs ON THE GPU BOX (the reference op has no CPU path, hv_cuda.cpp:26-28):
    python tools/make_golden.py gpurun_out/golden
It imports the prebuilt oracle/_ref/hv_cuda_ref.so (oracle/build_ref.py: the reference sources
compiled where they lie, sm_100a, arithmetic untouched), feeds it seeded synthetic scenes
(canonicalvoting_b200/synthetic.py) and stores inputs + reference outputs + the device cos/sin
table.  The files are then committed under tests/golden/ and pin the CPU oracle in the
`-m "not gpu"` suite (tests/test_oracle_vote.py::test_oracle_matches_reference_golden).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from canonicalvoting_b200 import hv_cuda as H  # noqa: E402
from canonicalvoting_b200 import synthetic  # noqa: E402
from oracle import build_ref  # noqa: E402

CASES = {
    "C1": dict(n_points=5000, grid=32, num_rots=4, seed=0),
    "ragged": dict(n_points=777, grid=20, num_rots=7, seed=2),
    "R120": dict(n_points=2000, grid=24, num_rots=120, seed=3),
    "uniform": dict(n_points=3000, grid=28, num_rots=12, seed=1, uniform=True),
}


def main(out_dir):
    ref = build_ref.load_ref()
    assert ref is not None, "oracle/_ref/hv_cuda_ref.so missing"
    os.makedirs(out_dir, exist_ok=True)
    for name, kw in CASES.items():
        sc = synthetic.make_scene(**kw)
        R = kw["num_rots"]
        p, x, s, o = (torch.from_numpy(sc[k]).cuda() for k in ("points", "xyz", "scale", "obj"))
        res_t = torch.tensor(0.03, dtype=torch.float32).cuda()
        rots_t = torch.tensor(R, dtype=torch.int32).cuda()
        go, gr, gs = ref.forward(p, x, s, o, res_t, rots_t)
        gen = torch.Generator(device="cpu").manual_seed(1234)
        grad = torch.randn(go.shape, generator=gen).cuda()
        d_xyz, d_scale, d_obj = ref.backward(grad, p, x, s, o, res_t, rots_t)
        ct, st = H.theta_table(R)
        torch.cuda.synchronize()
        path = os.path.join(out_dir, "hv_%s.npz" % name)
        np.savez_compressed(
            path, points=sc["points"], xyz=sc["xyz"], scale=sc["scale"], obj=sc["obj"], num_rots=R,
            theta_cos=ct.cpu().numpy(), theta_sin=st.cpu().numpy(),
            ref_grid_obj=go.cpu().numpy(), ref_grid_rot=gr.cpu().numpy(), ref_grid_scale=gs.cpu().numpy(),
            grad_grid=grad.cpu().numpy(), ref_d_xyz=d_xyz.cpu().numpy(), ref_d_scale=d_scale.cpu().numpy(),
            ref_d_obj=d_obj.cpu().numpy())
        print(name, tuple(go.shape), "max", float(go.max()), "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
