"""Developer tool (2+ GPUs, torchrun): step the flag-synchronised slab kernel one launch at a time and dump every rank's
pbmc_slab_sync block (tags, maxima, counters) after each launch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/slab_debug.py"""
import os
import struct
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def dump(st, tag):
    torch.cuda.synchronize()
    w = st._sync.cpu().numpy().view(np.uint32)
    slots = w[:64].reshape(2, 16, 2)  # [par][rank][lo = float bits, hi = tag]
    world = st.slab.world
    txt = " ".join(f"p{p}r{r}:(tag {int(slots[p, r, 1])}, {struct.unpack('f', struct.pack('I', int(slots[p, r, 0])))[0]:.4g})"
                   for p in range(2) for r in range(world))
    print(f"[rank {st.slab.rank}] {tag}: {txt} | steps_done {w[64]} ctas_done {w[65]} local_max {w[66]:#x} failed {w[67]:#x}",
          flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    st = bench._slab_setup(64 * world, 1024, rank, world, dev, "p2p", "flags")
    print(f"[rank {rank}] sync ptrs {[hex(p) for p in st._sync_ptrs]} own data_ptr {hex(st._sync.data_ptr())} "
          f"handle ptrs {[hex(int(p)) for p in st._sync_handle.buffer_ptrs]}", flush=True)
    dist.barrier()
    dump(st, "after publish")
    dist.barrier()
    for i in range(4):
        st._step_once()
        dump(st, f"after step {i + 1}")
        dist.barrier()
    st.step(8)
    dump(st, "after 8 more (graph)")
    full = st.gather()
    if rank == 0:
        one = bench._slab_setup(64 * world, 1024, 0, 1, dev, "nccl", "nccl")
        one.step(12)
        print("identical_to_single_gpu", bool(torch.equal(one.gather(), full)), flush=True)
    st.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
