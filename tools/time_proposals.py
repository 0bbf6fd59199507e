"""tools/time_proposals.py -- CUDA-event timing of the proposal sampler's device passes (csrc/hv_proposals.cu) on a C2 /
C5-sized vote map: the y-projection (HBM-bound, 4 G bytes) and one 768-draw rejection trial against 1024 seeds."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from canonicalvoting_b200 import proposals  # noqa: E402


def ev(fn, n=50):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


out = {}
for G in (128, 256):
    g = torch.rand(G, G, G, device="cuda")
    gs = torch.rand(G, G, G, 3, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t_proj = ev(lambda: (flush.zero_(), proposals.project_y(g))) - ev(lambda: flush.zero_())
    _, arg = proposals.project_y(g)
    dist = torch.rand(G * G, device="cuda")
    seeds = torch.rand(1024, 3, device="cuda") * G * 0.03
    st = proposals._Trial(512, 768, g.device)
    sample = torch.multinomial(dist, 768, replacement=True)

    def trial():
        st.count.zero_()
        proposals.append_proposals(st, sample, arg, gs, 0.03, (0.0, 0.0, 0.0), seeds)
    out["G%d" % G] = {"project_y_us_l2_cold": t_proj, "project_y_GBps": 4 * G ** 3 / t_proj / 1e3, "trial_us": ev(trial),
                      "torch_max_plus_argmax_us": ev(lambda: (g.max(1)[0], torch.argmax(g, 1)))}
print(json.dumps(out))
