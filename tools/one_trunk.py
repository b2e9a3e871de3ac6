"""Developer tool: launch the persistent trunk kernel (pbmc_trunk_fwd, R layers of one level) a few times -- the target
for ncu.  usage: one_trunk.py H W [R] [B] [max_ctas] [impl] [loader = bulk | threads]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbml_mantle_convection_b200 import _lib as L  # noqa: E402
from pbml_mantle_convection_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
H, W = int(sys.argv[1]), int(sys.argv[2])
R = int(sys.argv[3]) if len(sys.argv) > 3 else 4
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
max_ctas = int(sys.argv[5]) if len(sys.argv) > 5 else 0
impl = sys.argv[6] if len(sys.argv) > 6 else "mux_f16x2"
loader = sys.argv[7] if len(sys.argv) > 7 else "threads"
g = torch.Generator(device=dev).manual_seed(5)
x = torch.randn(B, 4, H, W, 4, device=dev, generator=g)
stats = torch.stack([x.double().sum((2, 3, 4)), (x.double() ** 2).sum((2, 3, 4))], -1).contiguous()


class Lay:
    def __init__(self):
        w = torch.randn(16, 16, 3, 3, device=dev, generator=g) / 12
        self.wpk, self.wpk_row = ops.pack_conv_weight(w, [16]), ops.pack_conv_weight_row(w, [16])
        self.bias, self.gamma, self.beta = torch.zeros(16, device=dev), torch.ones(16, device=dev), torch.zeros(16, device=dev)
        self.cin_blks, self.cout, self.ksize = 4, 16, 3


lays = [Lay() for _ in range(R)]
src = ops.Source(x, L.XFORM_GN_GELU, stats, torch.ones(16, device=dev), torch.zeros(16, device=dev))
ping = [torch.empty_like(x), torch.empty_like(x)]
st = torch.empty(R, B, 4, 2, dtype=torch.float64, device=dev)
sync = torch.empty(B, dtype=torch.int32, device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
ev = []
for _ in range(6):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.trunk_fwd(src, lays, "replicate", impl=impl, max_ctas=max_ctas, ping=ping, stats=st, sync=sync, loader=loader)
    b.record()
    ev.append((a, b))
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev[2:])
print(f"trunk {H}x{W} R={R} B={B} max_ctas={max_ctas} {impl} loader={loader}: {ms[len(ms) // 2] * 1e3:.1f} us per launch, {ms[len(ms) // 2] * 1e3 / R:.1f} us per layer")
