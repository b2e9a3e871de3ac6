"""Developer micro-benchmark: CUDA-event timings of single conv launches on the rollout's shapes.
usage: python tools/kbench.py [impl ...]   (default: row_f16x2 umma_f16x2)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbml_mantle_convection_b200 import _lib as L  # noqa: E402
from pbml_mantle_convection_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)


def timeit(fn, iters=30, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    impls = sys.argv[1:] or ["mux_f16x2", "row_f16x2"]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    w = torch.randn(16, 16, 3, 3, device=dev, generator=g) / 12
    w1 = torch.randn(16, 103, 3, 3, device=dev, generator=g) / 30
    bias = torch.zeros(16, device=dev)
    gam, bet = torch.ones(16, device=dev), torch.zeros(16, device=dev)
    pk = dict(wpk=ops.pack_conv_weight(w, [16]), wpk_umma=ops.pack_conv_weight_umma(w, [16]), wpk_row=ops.pack_conv_weight_row(w, [16]))
    ch1 = [16] * 6 + [7]
    pk1 = dict(wpk=ops.pack_conv_weight(w1, ch1), wpk_umma=ops.pack_conv_weight_umma(w1, ch1), wpk_row=ops.pack_conv_weight_row(w1, ch1))
    for B, H, W in [(1, 512, 512), (1, 256, 256), (1, 128, 128), (1, 64, 64), (1, 32, 32), (1, 16, 16), (32, 256, 256), (1, 128, 506)]:
        x = torch.randn(B, 4, H, W, 4, device=dev, generator=g)
        stats = torch.stack([x.double().sum((2, 3, 4)), (x.double() ** 2).sum((2, 3, 4))], -1).contiguous()
        o, st = torch.empty_like(x), torch.zeros_like(stats)
        for xf, name in [(L.XFORM_GN_GELU, "gn+gelu"), (L.XFORM_NONE, "plain")]:
            src = ops.Source(x, xf, stats if xf else None, gam if xf else None, bet if xf else None)
            for impl in impls:
                f = lambda: ops.conv_fwd([src], pk["wpk"], bias, 16, 3, "replicate", impl=impl, wpk_umma=pk["wpk_umma"],
                                         wpk_row=pk["wpk_row"], out=o, stats=st)
                cold, _ = timeit(f, flush=flush)
                warm, best = timeit(f)
                cells = B * H * W
                print(f"conv16x16 {name:8s} B{B} {H}x{W} {impl:12s} cold {cold:8.1f} us  warm {warm:8.1f} us (best {best:.1f})  "
                      f"warm: {cells * 128 / warm / 1e3:7.1f} GB/s  {cells * 4608 / warm / 1e6:6.1f} TF/s", flush=True)
        if (H, W) in [(512, 512), (256, 256), (128, 506)]:
            srcs = [ops.Source(x, L.XFORM_GN_GELU, stats, gam, bet)] + [ops.Source(torch.randn_like(x)) for _ in range(5)] + \
                   [ops.Source(torch.randn(B, 2, H, W, 4, device=dev, generator=g))]
            for impl in impls:
                f = lambda: ops.conv_fwd(srcs, pk1["wpk"], bias, 16, 3, "replicate", impl=impl, wpk_umma=pk1["wpk_umma"],
                                         wpk_row=pk1["wpk_row"], out=o, stats=st)
                cold, _ = timeit(f, flush=flush)
                warm, best = timeit(f)
                cells = B * H * W
                print(f"conv1 103->16      B{B} {H}x{W} {impl:12s} cold {cold:8.1f} us  warm {warm:8.1f} us (best {best:.1f})  "
                      f"{cells * 29664 / warm / 1e6:6.1f} TF/s", flush=True)
            if "row_f16x2" in impls:
                # the five up-sampled levels as fp16 hi|lo operand images (bulk-copied into the stage)
                half = ops.Source(torch.randn(B, 4, H // 2, W // 2, 4, device=dev, generator=g))
                ssrcs = [srcs[0]] + [ops.bicubic_up(half, H, W, staged=True) for _ in range(5)] + [srcs[-1]]
                f = lambda: ops.conv_fwd(ssrcs, pk1["wpk"], bias, 16, 3, "replicate", impl="row_f16x2", wpk_row=pk1["wpk_row"], out=o, stats=st)
                cold, _ = timeit(f, flush=flush)
                warm, best = timeit(f)
                print(f"conv1 103->16 TMA  B{B} {H}x{W} row_f16x2    cold {cold:8.1f} us  warm {warm:8.1f} us (best {best:.1f})  "
                      f"{B * H * W * 29664 / warm / 1e6:6.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
