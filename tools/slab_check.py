"""Multi-GPU check of the slab-decomposed stencil (config 5): run under torchrun on N GPUs of one box,
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py [H W steps]
every rank advances its slab with NCCL halo exchange + dt all-reduce; rank 0 also runs the whole grid alone and
the gathered result must be bit-identical.  Prints device-timed cell-updates/s (max over ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbml_mantle_convection_b200 import multigpu as MG  # noqa: E402
from pbml_mantle_convection_b200 import rollout as RO  # noqa: E402


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    halo = sys.argv[4] if len(sys.argv) > 4 else "nccl"
    dt_sync = sys.argv[5] if len(sys.argv) > 5 else "flags"  # p2p only: "flags" = reduction inside the kernel, "nccl" = all_reduce
    graph_steps = int(sys.argv[6]) if len(sys.argv) > 6 else 2  # time steps per captured CUDA graph (0: no graph, eager launches)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    xc, yc = RO.synthetic_grid(H, W)
    T = RO.synthetic_T0(H, W, seed=1).astype(np.float32)
    import bench  # the stable cellular flow of the slab workloads (round 1's field blew up: NaN != NaN looked like a mismatch)
    ut, vt = bench.slab_velocity(torch.tensor(xc[0], dtype=torch.float32), torch.tensor(yc[:, 0], dtype=torch.float32), H, W)
    u, v = ut.numpy(), vt.numpy()
    st = MG.SlabStencil(H, W, xc[0], yc[:, 0], rank, world, dev, raq=2.0, halo=halo, dt_sync=dt_sync)
    st.graph_steps = graph_steps
    st.scatter(T, u, v)
    st.step(3)  # warm-up (NCCL communicators, module load)
    st.scatter(T, u, v)
    st.step(max(4, 2 * graph_steps), use_graph=graph_steps > 0)  # untimed: capture
    st.scatter(T, u, v)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st.step(steps, use_graph=graph_steps > 0)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    full = st.gather()
    ok = True
    if rank == 0:
        ref = MG.SlabStencil(H, W, xc[0], yc[:, 0], 0, 1, dev, raq=2.0)
        ref.scatter(T, u, v)
        ref.step(steps)
        ok = bool(torch.equal(ref.gather(), full)) and bool(torch.isfinite(full).all())
        print(f"slab_check H={H} W={W} world={world} steps={steps} halo={halo} dt_sync={st.dt_sync} graph_steps={graph_steps}: identical_to_single_gpu={ok} "
              f"{H * W * steps / (ms.item() * 1e-3):.4g} cell-updates/s ({ms.item() / steps:.3f} ms/step)", flush=True)
    st.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
