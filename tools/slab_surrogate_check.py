"""Multi-GPU check of the slab-decomposed SURROGATE step (config 5 "stretch"; pbml_mantle_convection_b200/slab_surrogate.py):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
        tools/slab_surrogate_check.py [H W steps levels]
every rank advances its slab (NCCL halo exchange + all-reduces); rank 0 also runs the whole grid alone with the fused
single-GPU rollout and compares.  Prints ms/step of both (host-driven layer-by-layer decomposition vs the fused graph)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbml_mantle_convection_b200 as P  # noqa: E402
from pbml_mantle_convection_b200 import slab_surrogate as SS  # noqa: E402

PARAMS = (6.79733173, 475523342.0, 2.58574662)


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    levels = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = P.NewFluidNet(levels, 7, 16, 2, dev, act_fn="gelu", r_p="replicate", loss_type="curl", use_symm=True, a_bound=10,
                        repeats=4, f=3, p_pred=True).to(dev).eval()
    xc, yc = P.synthetic_grid(H, W)
    T0 = P.synthetic_T0(H, W, seed=1).astype(np.float32)
    s = SS.SlabSurrogate(net, H, W, xc[0], yc[:, 0], PARAMS, SS.DistComm(), dev)
    s.set_T(T0)
    s.step()  # warm-up (NCCL communicators, module load)
    s.set_T(T0)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    dts = [s.step() for _ in range(steps)]
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    T = s.gather(s.T)[0]
    u = s.gather(s.u)[0]
    ok = True
    if rank == 0:
        ens = P.EnsembleRollout(net, H, W, [PARAMS], dev, xc=xc, yc=yc, cn_max=0.99, per_member_dt=False)
        ens.set_T(T0[None])
        ens.step(steps)
        torch.cuda.synchronize()
        ens.n_done = 0  # BEFORE set_T: it writes the slot of the current step parity
        ens.set_T(T0[None])
        t0 = time.perf_counter()
        ens.step(steps)
        torch.cuda.synchronize()
        ms1 = (time.perf_counter() - t0) * 1e3 / steps
        eT = (T - ens.T[0]).abs().max().item()
        eu = ((u - ens.fields()[0][0]).abs().max() / ens.fields()[0][0].abs().max()).item()
        ok = eT <= 5e-6 and eu <= 2e-4 and bool(torch.isfinite(T).all())  # see tests/test_gpu_slab_surrogate.py for the bounds
        rows_bad = ((u - ens.fields()[0][0]).abs().max(dim=1).values > 1e-3 * float(ens.fields()[0][0].abs().max())).nonzero().flatten().tolist()
        print(f"  rows with |du| > 1e-3 max|u|: {len(rows_bad)} {rows_bad[:16]}{' ...' if len(rows_bad) > 16 else ''}", flush=True)
        print(f"slab_surrogate_check H={H} W={W} world={world} steps={steps} levels={levels}: matches_single_gpu={ok} "
              f"(max|dT| {eT:.2e}, rel max|du| {eu:.2e}); {ms:.2f} ms/step decomposed (host-driven, wall clock) vs {ms1:.2f} ms/step "
              f"fused single GPU", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
