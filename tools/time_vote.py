#!/usr/bin/env python
"""tools/time_vote.py -- the vote op alone (hv_cuda.forward_host, geometry known): CUDA events per call, L2 flushed or warm,
(GPU box)."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from canonicalvoting_b200 import hv_cuda as H, synthetic  # noqa: E402

out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for wl in ("C2", "C5") + (("R120",) if "all" in sys.argv else ()):
    sc = synthetic.make_scene(50000, 128, 120, seed=0) if wl == "R120" else synthetic.make_config(wl, seed=0)
    p, x, s, o = (torch.from_numpy(sc[k]).cuda() for k in ("points", "xyz", "scale", "obj"))
    res, R = sc["res"], sc["num_rots"]
    corner, _, dims = H.grid_dims(p, res)
    nbytes = 40 * len(sc["points"]) + 24 * dims[0] * dims[1] * dims[2]
    for impl in (0,):
        for cold in (True, False):
            ts = []
            for it in range(13):
                if cold:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                H.forward_host(p, x, s, o, res, R, corner, dims)
                b.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            key = "%s %s" % (wl, "L2 cold" if cold else "L2 warm")
            out[key] = {"median_us": ts[len(ts) // 2], "min_us": ts[0], "GBps": nbytes / ts[len(ts) // 2] / 1e3}
            print("%-28s median %8.1f us  min %8.1f us  -> %7.0f GB/s of 40N+24G" % (key, ts[len(ts) // 2], ts[0], nbytes / ts[len(ts) // 2] / 1e3), flush=True)
json.dump(out, open("gpurun_out/time_vote.json", "w"), indent=1)
