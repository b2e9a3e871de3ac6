"""Developer tool (ONE GPU): the slab-decomposed surrogate step with the ranks as threads of one process (ThreadComm) at an
arbitrary size / rank count, against the fused single-GPU rollout.  usage: slab_surrogate_threads.py H W world levels [steps]"""
import os
import sys
import threading

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbml_mantle_convection_b200 as P  # noqa: E402
from pbml_mantle_convection_b200 import slab_surrogate as SS  # noqa: E402

PARAMS = (6.79733173, 475523342.0, 2.58574662)
H, W, world, levels = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = P.NewFluidNet(levels, 7, 16, 2, dev, act_fn="gelu", r_p="replicate", loss_type="curl", use_symm=True, a_bound=10, repeats=4,
                    f=3, p_pred=True).to(dev).eval()
xc, yc = P.synthetic_grid(H, W)
T0 = P.synthetic_T0(H, W, seed=1).astype(np.float32)
ens = P.EnsembleRollout(net, H, W, [PARAMS], dev, xc=xc, yc=yc, cn_max=0.99, per_member_dt=False)
ens.set_T(T0[None])
ens.step(steps)
u_r, T_r = ens.fields()[0][0].clone(), ens.T[0].clone()
shared = SS.ThreadComm.Shared(world)
out, err = [None] * world, []


def run(r):
    try:
        torch.cuda.set_device(dev)
        s = SS.SlabSurrogate(net, H, W, xc[0], yc[:, 0], PARAMS, SS.ThreadComm(shared, r), dev)
        s.set_T(T0)
        for _ in range(steps):
            s.step()
        out[r] = (s.gather(s.T)[0], s.gather(s.u)[0])
    except Exception as e:  # noqa: BLE001
        err.append(e)
        shared.barrier.abort()


th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
[t.start() for t in th]
[t.join() for t in th]
if err:
    raise err[0]
T, u = out[0]
eu = (u - u_r).abs()
rows = eu.max(dim=1).values
bad = (rows > 1e-3 * float(u_r.abs().max())).nonzero().flatten().tolist()
print(f"slab_surrogate_threads H={H} W={W} world={world} levels={levels} steps={steps}: max|dT| {(T - T_r).abs().max().item():.2e}, "
      f"rel max|du| {(eu.max() / u_r.abs().max()).item():.2e}; rows with |du| > 1e-3 max|u|: {len(bad)} "
      f"{bad[:12]}{' ...' if len(bad) > 12 else ''}", flush=True)
