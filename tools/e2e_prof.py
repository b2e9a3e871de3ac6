"""Developer tool: where does the end-to-end TS.forward(ts=1) step spend its time?  (host call, copies, waits)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN
import pbml_mantle_convection_b200 as PK
from pbml_mantle_convection_b200 import pytorch_networks_convae as P

dev = torch.device("cuda:0")
H = W = 512
net = BN.primary_net(dev)
ts = P.TS(net, P.ADNet(dev, CN_max=0.99), dev, ts=1, scale=True, p_pred=True, net="newfluidnet")
xc, yc = PK.synthetic_grid(H, W)
t64 = lambda a_: torch.tensor(a_, dtype=torch.float64)
xc_t, yc_t = t64(xc).view(1, 1, H, W), t64(yc).view(1, 1, H, W)
P0 = BN.PARAMS0
nd = ((P0[0] - 0.12624371) / (9.70723344 - 0.12624371), (np.log10(P0[1]) - 6.00352841978384) / (9.888820429862925 - 6.00352841978384),
      (np.log10(P0[2]) - 0.005251646002323797) / (1.9927988938926755 - 0.005251646002323797))
args_ts = (None, None, yc_t, t64(nd[0]), t64(nd[1]), t64(nd[2]), t64(P0[0]), t64(P0[1]), t64(P0[2]), xc_t, yc_t)
rng = np.random.default_rng(1)
T0 = (1 - yc) + 0.01 * rng.random((H, W))
Tp = t64(T0).view(1, 1, H, W).pin_memory()
pin = lambda: torch.empty(1, 1, H, W, dtype=torch.float64).pin_memory()
hostT, hostF, host_dt = [pin(), pin()], [pin(), pin(), pin()], torch.empty(1, dtype=torch.float64).pin_memory()
acc = {"call": 0.0, "enqueue_d2h": 0.0, "sync": 0.0}
def step(Tp_, k):
    t0 = time.perf_counter()
    x, dts, u, v, p, V = ts(Tp_, *args_ts)
    t1 = time.perf_counter()
    Tn_ = hostT[k % 2]
    Tn_.copy_(x[1], non_blocking=True)
    for dst, src in zip(hostF, (u, v, V)):
        dst.copy_(src, non_blocking=True)
    host_dt.copy_(dts[1].reshape(1), non_blocking=True)
    t2 = time.perf_counter()
    torch.cuda.current_stream(dev).synchronize()
    t3 = time.perf_counter()
    acc["call"] += t1 - t0; acc["enqueue_d2h"] += t2 - t1; acc["sync"] += t3 - t2
    return Tn_
Tc = Tp
for k in range(5):
    Tc = step(Tc, k)
for a in acc: acc[a] = 0.0
N = 50
t0 = time.perf_counter()
for k in range(N):
    Tc = step(Tc, k)
tot = time.perf_counter() - t0
print(f"per step {tot / N * 1e3:.3f} ms:", {a: f"{v / N * 1e3:.3f}" for a, v in acc.items()})
# raw copies
a = torch.empty(1, 1, H, W, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    a.copy_(Tp, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D 2 MiB: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(50):
    hostT[0].copy_(a, non_blocking=True)
torch.cuda.synchronize()
print(f"D2H 2 MiB: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
big = torch.empty(4, 1, H, W, dtype=torch.float64, device=dev); hbig = torch.empty(4, 1, H, W, dtype=torch.float64).pin_memory()
t0 = time.perf_counter()
for _ in range(50):
    hbig.copy_(big, non_blocking=True)
torch.cuda.synchronize()
print(f"D2H 8 MiB in one copy: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for k in range(30):
    Tc = step(Tc, k)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
