"""Developer probe (torchrun, >= 1 GPU): is the stencil slower when T lives in torch symmetric memory (peer-mapped VMM
allocations) than in ordinary cudaMalloc'ed tensors?  Times the plain single-domain kernel both ways on every rank."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pbml_mantle_convection_b200 import ops  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 8192)
grid = bench.eng_grid(H, W, dev)
g = torch.Generator(device=dev).manual_seed(1)
u = torch.randn(1, H, W, device=dev, generator=g)
v = torch.randn(1, H, W, device=dev, generator=g)
members = ops.make_members([bench.PARAMS0], dev)
uv = ops.uvmax_reduce(u, v)
dto = torch.empty(1, dtype=torch.float64, device=dev)


def timeit(Ta, Tb, tag):
    Ta.copy_(torch.rand(1, H, W, device=dev, generator=g))
    bufs = [Ta, Tb]
    for i in range(4):
        ops.advect_diffuse(bufs[i % 2], u, v, grid.xcoef, grid.ycoef, members, uv, grid.dx_min, 0.99, T_out=bufs[(i + 1) % 2], dt_out=dto)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(40):
        ops.advect_diffuse(bufs[i % 2], u, v, grid.xcoef, grid.ycoef, members, uv, grid.dx_min, 0.99, T_out=bufs[(i + 1) % 2], dt_out=dto)
    b.record()
    torch.cuda.synchronize()
    print(f"[rank {rank}] {tag}: {a.elapsed_time(b) / 40 * 1e3:.1f} us per {H}x{W} sweep", flush=True)


timeit(torch.empty(1, H, W, device=dev), torch.empty(1, H, W, device=dev), "T in ordinary memory ")
buf = symm_mem.empty((2, 1, H, W), dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(buf, group=dist.group.WORLD)
timeit(buf[0], buf[1], "T in symmetric memory")
dist.barrier()
dist.destroy_process_group()
