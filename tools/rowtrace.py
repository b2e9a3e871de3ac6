"""Developer tool: in-kernel timeline of conv_row_kernel's CTA (0,0,0).  Needs a trace build:
   PBMC_EXTRA_NVCC_FLAGS=-DPBMC_ROW_TRACE python pbml_mantle_convection_b200/build.py --force
usage: python tools/rowtrace.py H W [xform]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbml_mantle_convection_b200 import _lib as L  # noqa: E402
from pbml_mantle_convection_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
H, W = int(sys.argv[1]), int(sys.argv[2])
xf = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nsrc = int(sys.argv[4]) if len(sys.argv) > 4 else 1
lib = C.CDLL(L.LIB_PATH)
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
lib.pbmc_debug_set_row_trace(C.c_void_p(buf.data_ptr()))
g = torch.Generator(device=dev).manual_seed(5)
torch.manual_seed(0)
x = torch.randn(1, 4, H, W, 4, device=dev, generator=g)
stats = torch.stack([x.double().sum((2, 3, 4)), (x.double() ** 2).sum((2, 3, 4))], -1).contiguous()
gam, bet, bias = torch.ones(16, device=dev), torch.zeros(16, device=dev), torch.zeros(16, device=dev)
ch = [16] * nsrc
w = torch.randn(16, sum(ch), 3, 3, device=dev, generator=g) / 12
wpk, wrow = ops.pack_conv_weight(w, ch), ops.pack_conv_weight_row(w, ch)
srcs = [ops.Source(x, xf, stats if xf else None, gam if xf else None, bet if xf else None)] + [ops.Source(torch.randn_like(x)) for _ in range(nsrc - 1)]
o, st = torch.empty_like(x), torch.zeros_like(stats)
for _ in range(3):
    buf.zero_()
    ops.conv_fwd(srcs, wpk, bias, 16, 3, "replicate", impl="row_f16x2", wpk_row=wrow, out=o, stats=st)
    torch.cuda.synchronize()
t = buf.cpu().numpy()
t0 = t[0]
rel = lambda i: int(t[i] - t0) if t[i] else None
print("setup done", rel(1), " epilogue loop done", rel(3), " end", rel(2))
NPG = int(sys.argv[5]) if len(sys.argv) > 5 else 3
for pg in range(NPG):
    print(f"P{pg}: role start {rel(4+pg)}  prologue done {rel(8+pg)}")
    k = 0
    while t[100 + pg * 300 + 8 * k]:
        b = 100 + pg * 300 + 8 * k
        print(f"P{pg} stage {NPG*k+pg:3d}: top {rel(b)}  xformed {rel(b+1)}  slot-free {rel(b+2)}  stored {rel(b+4)}  fenced {rel(b+5)}  arrived {rel(b+3)}  next-load-issued {rel(b+6)}")
        k += 1
st_ = 0
while t[1400 + 2 * st_]:
    print(f"MMA stage {st_:3d}: full {rel(1400+2*st_)}  issued {rel(1401+2*st_)}")
    st_ += 1
yo = 0
while t[1200 + 3 * yo]:
    print(f"EPI row {yo:3d}: d_full {rel(1200+3*yo)}  loaded {rel(1201+3*yo)}  stored {rel(1202+3*yo)}")
    yo += 1
