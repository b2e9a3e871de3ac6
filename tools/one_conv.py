"""Developer tool: launch one conv shape a few times (target for ncu).  usage: one_conv.py H W xform nsrc impl [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbml_mantle_convection_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
H, W, xf, nsrc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
impl = sys.argv[5] if len(sys.argv) > 5 else "row_f16x2"
B = int(sys.argv[6]) if len(sys.argv) > 6 else 1
g = torch.Generator(device=dev).manual_seed(5)
x = torch.randn(B, 4, H, W, 4, device=dev, generator=g)
stats = torch.stack([x.double().sum((2, 3, 4)), (x.double() ** 2).sum((2, 3, 4))], -1).contiguous()
gam, bet, bias = torch.ones(16, device=dev), torch.zeros(16, device=dev), torch.zeros(16, device=dev)
ch = [16] * nsrc
w = torch.randn(16, sum(ch), 3, 3, device=dev, generator=g) / 12
wpk, wrow, wum = ops.pack_conv_weight(w, ch), ops.pack_conv_weight_row(w, ch), ops.pack_conv_weight_umma(w, ch)
srcs = [ops.Source(x, xf, stats if xf else None, gam if xf else None, bet if xf else None)] + \
       [ops.Source(torch.randn_like(x)) for _ in range(nsrc - 1)]
o, st = torch.empty_like(x), torch.zeros_like(stats)
for _ in range(5):
    ops.conv_fwd(srcs, wpk, bias, 16, 3, "replicate", impl=impl, wpk_row=wrow, wpk_umma=wum, out=o, stats=st)
torch.cuda.synchronize()
print("ok")
