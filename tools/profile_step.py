"""cProfile of the host side of the C2 step (GPU box): where does the CPU time per scene go?"""
import cProfile
import pstats
import sys
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench
from canonicalvoting_b200 import hv_cuda as H
from canonicalvoting_b200.engine import MinkUNetEngine
sc = bench.scene_for("C2", 0)
model = bench.make_model().cuda()
eng = MinkUNetEngine(model, 9, True, pipeline=True)
c_h, f_h = bench.scene_tensors(sc)
c, f = c_h.cuda(), f_h.cuda()
res, R = sc["res"], sc["num_rots"]
pts = (c[:, 1:].float() * res).contiguous()
corner, _, dims = H.grid_dims(pts, res)


def step():
    xyz, scale, cls, prob = eng.predict(c, f)
    p = (c[:, 1:].float() * res).contiguous()
    return H.forward_host(p, xyz, scale, prob, res, R, corner, dims)


for _ in range(10):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
