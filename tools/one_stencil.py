"""Developer tool: a few stencil sweeps at N x N (target for ncu).  usage: one_stencil.py N"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pbml_mantle_convection_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
grid = bench.eng_grid(n, n, dev)
g = torch.Generator(device=dev).manual_seed(7)
T = torch.rand(1, n, n, device=dev, generator=g)
u = torch.randn(1, n, n, device=dev, generator=g) * 100
v = torch.randn(1, n, n, device=dev, generator=g) * 100
members = ops.make_members([bench.PARAMS0], dev)
uv = [ops.uvmax_reduce(u, v), torch.zeros(1, dtype=torch.int32, device=dev)]
Tb = [T, torch.empty_like(T)]
dto = torch.empty(1, dtype=torch.float64, device=dev)
for i in range(6):
    uv[(i + 1) % 2].zero_()
    ops.advect_diffuse(Tb[i % 2], u, v, grid.xcoef, grid.ycoef, members, uv[i % 2], grid.dx_min, 0.99, T_out=Tb[(i + 1) % 2],
                       dt_out=dto, uv_out=uv[(i + 1) % 2])
torch.cuda.synchronize()
print("ok")
