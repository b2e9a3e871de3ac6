"""Host logic of the slab-decomposed surrogate WITHOUT a GPU: the device operators are
the CPU stand-ins of tests/_emulated_ops.py (float64), the ranks are threads of one process (`ThreadComm`) or a world-2 gloo
group (`DistComm`), and every rank's slab of (u, v, p) and the global max|u|,|v| must equal the single-domain forward of the
same emulated operators -- to rounding of the re-associated GroupNorm / zero-mean sums."""
import os
import socket
import threading

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pbml_mantle_convection_b200 as P
from oracle import ref_numpy as RN
from pbml_mantle_convection_b200 import ops
from pbml_mantle_convection_b200 import slab_surrogate as SS
from tests import _emulated_ops as emu
from tests._util import load_weights

PARAMS = (6.79733173, 475523342.0, 2.58574662)
LEVELS, REPEATS = 3, 2
H, W = 48, 40


def _net():
    torch.manual_seed(3)
    net = P.NewFluidNet(LEVELS, 7, 16, 2, "cpu", act_fn="gelu", r_p="replicate", loss_type="curl", use_symm=True, a_bound=10,
                        repeats=REPEATS, f=3, p_pred=True).double().eval()
    with torch.no_grad():
        for n, p_ in net.named_parameters():
            if "layers.1" in n or n.startswith("gn."):
                p_.add_(0.1 * torch.randn_like(p_))
    return net


def _single_domain(net, T0, xc, yc):
    one = SS.SlabSurrogate(net, H, W, xc[0], yc[:, 0], PARAMS, SS.ThreadComm(SS.ThreadComm.Shared(1), 0), "cpu")
    one.set_T(T0)
    inp, _ = ops.build_input(one.T, one.xc, one.yc, one.yc, one.members)
    u, v, p, uvmax = one.forward(inp)
    return u[0].numpy(), v[0].numpy(), p[0].numpy(), int(uvmax[0])


def _check(got, ref):
    u, v, p, uvmax = ref
    for name, a, b in (("u", got[0], u), ("v", got[1], v), ("p", got[2], p)):
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(b).max()), name
    assert got[3] == uvmax


@pytest.mark.parametrize("world", [2, 3])
def test_slab_surrogate_threads_equal_single_domain(monkeypatch, world):
    emu.install(monkeypatch)
    Hh = {2: 48, 3: 48}[world]
    assert Hh == H
    net = _net()
    xc, yc = RN.synthetic_grid(H, W)
    T0 = RN.synthetic_T0(H, W, seed=5)
    ref = _single_domain(net, T0, xc, yc)
    # the single-domain SlabSurrogate (world 1) is itself the plain network: against NewFluidNet.forward's emulated path
    inp7, _ = RN.build_input(T0[None, None], xc, yc, yc, *PARAMS)
    shared = SS.ThreadComm.Shared(world)
    out, err = [None] * world, []

    def run(r):
        try:
            s = SS.SlabSurrogate(net, H, W, xc[0], yc[:, 0], PARAMS, SS.ThreadComm(shared, r), "cpu")
            s.set_T(T0)
            inp, _ = ops.build_input(s.T, s.xc, s.yc, s.yc, s.members)
            u, v, p, uvmax = s.forward(inp)
            out[r] = (s.gather(u)[0].numpy(), s.gather(v)[0].numpy(), s.gather(p)[0].numpy(), int(uvmax[0]))
        except Exception as e:  # a dead rank would leave the others waiting at the barrier
            err.append(e)
            shared.barrier.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    for r in range(world):
        _check(out[r], ref)
    # and the network itself: the oracle's float64 forward of the same weights
    sd = {k: t.detach().numpy() for k, t in net.state_dict().items()}
    spec = RN.NetSpec(levels=LEVELS, repeats=REPEATS)
    u_o, v_o, p_o = RN.newfluidnet_forward(sd, spec, inp7)
    s = RN.velocity_scaler(*PARAMS)
    # (the slab object stores T and the coordinates in float32, as the device path does: ~1e-7 on the input, amplified by the curl)
    assert np.abs(ref[0] - u_o[0] * s).max() <= 1e-5 * np.abs(u_o * s).max() and np.abs(ref[2] - p_o[0]).max() <= 1e-5 * np.abs(p_o).max()


def test_slab_surrogate_refuses_bad_splits(monkeypatch):
    emu.install(monkeypatch)
    net = _net()
    xc, yc = RN.synthetic_grid(H, W)
    comm = SS.ThreadComm(SS.ThreadComm.Shared(1), 0)
    comm.world, comm.rank = 5, 0
    with pytest.raises(ValueError):
        SS.SlabSurrogate(net, H, W, xc[0], yc[:, 0], PARAMS, comm, "cpu")  # 48 rows / 5 ranks
    comm.world = 6
    with pytest.raises(ValueError):
        SS.SlabSurrogate(net, H, W, xc[0], yc[:, 0], PARAMS, comm, "cpu")  # 8 rows: the coarsest level would own 2 < 3 ghost rows


def _gloo_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mp_ = pytest.MonkeyPatch()
    try:
        emu.install(mp_)
        net = _net()
        xc, yc = RN.synthetic_grid(H, W)
        T0 = RN.synthetic_T0(H, W, seed=5)
        s = SS.SlabSurrogate(net, H, W, xc[0], yc[:, 0], PARAMS, SS.DistComm(), "cpu")
        s.set_T(T0)
        inp, _ = ops.build_input(s.T, s.xc, s.yc, s.yc, s.members)
        u, v, p, uvmax = s.forward(inp)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), u=s.gather(u)[0].numpy(), v=s.gather(v)[0].numpy(), p=s.gather(p)[0].numpy(),
                 uvmax=np.int64(int(uvmax[0])))
    finally:
        mp_.undo()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_slab_surrogate_gloo_world2(monkeypatch, tmp_path):
    mp.spawn(_gloo_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    emu.install(monkeypatch)
    net = _net()
    xc, yc = RN.synthetic_grid(H, W)
    ref = _single_domain(net, RN.synthetic_T0(H, W, seed=5), xc, yc)
    for r in range(2):
        d = np.load(tmp_path / f"r{r}.npz")
        _check((d["u"], d["v"], d["p"], int(d["uvmax"])), ref)
