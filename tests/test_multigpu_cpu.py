"""Host logic of the multi-GPU partitioning on CPU: world-size 2 and 3 `gloo` process
groups drive `SlabStencil` (row slabs + one-row T halo exchange + global dt all-reduce) with the numpy oracle as
the local update, and the gathered field must equal the single-domain oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_numpy as RN
from pbml_mantle_convection_b200 import multigpu as MG

H, W, STEPS, RAQ, CN = 37, 24, 6, 3.0, 0.99


def _fields():
    xc, yc = RN.synthetic_grid(H, W)
    r = np.random.default_rng(11)
    T = RN.synthetic_T0(H, W, seed=3).astype(np.float32)
    psi = np.sin(np.pi * xc / 4 * 2) * np.sin(np.pi * yc)
    u = (np.gradient(psi, axis=0) * 4e3 + r.standard_normal((H, W))).astype(np.float32)
    v = (-np.gradient(psi, axis=1) * 4e3 + r.standard_normal((H, W))).astype(np.float32)
    return xc, yc, T, u, v


def _oracle_callables(st, xc):
    """local_step / local_uvmax for SlabStencil built from the numpy oracle (float64 arithmetic, float32 storage)."""
    s = st.slab
    xc_loc = xc[s.l0:s.l1]
    yc_loc = np.broadcast_to(st.y_loc[:, None], (s.rows, W)).copy()

    def uvmax(u, v):
        ui, vi = u[0, 1:-1, 1:-1].abs().max(), v[0, 1:-1, 1:-1].abs().max()
        return torch.maximum(ui, vi).reshape(1).float().view(torch.int32).clone()

    def step(T, u, v, bits):
        uvm = float(bits.view(torch.float32)[0])
        dt = MG.cfl_dt(uvm, st.dx_min, CN)
        out, _ = RN.adnet_forward(u.numpy().astype(np.float64), v.numpy().astype(np.float64), T.numpy().astype(np.float64), RAQ,
                                  xc_loc, yc_loc, CN, dt=dt, y_walls=(not s.up, not s.down))
        return torch.tensor(out.astype(np.float32)), dt

    return step, uvmax


def _reference():
    xc, yc, T, u, v = _fields()
    T = T[None].astype(np.float64)
    dts = []
    for _ in range(STEPS):
        uvm = float(np.float32(max(np.abs(u[1:-1, 1:-1]).max(), np.abs(v[1:-1, 1:-1]).max())))
        dx = xc.copy()
        dx[:, 0], dx[:, -1] = 0.0, 4.0
        dt = MG.cfl_dt(uvm, float((dx[1:-1, 1:-1] - dx[1:-1, :-2]).min()), CN)
        T, _ = RN.adnet_forward(u[None].astype(np.float64), v[None].astype(np.float64), T, RAQ, xc, yc, CN, dt=dt)
        T = T.astype(np.float32).astype(np.float64)  # float32 storage between steps, like the slabs
        dts.append(dt)
    return T[0].astype(np.float32), dts


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xc, yc, T, u, v = _fields()
        st = MG.SlabStencil(H, W, xc[0], yc[:, 0], rank, world, "cpu", raq=RAQ, cn_max=CN, local_step=lambda *a: None,
                            local_uvmax=lambda *a: None)
        st.local_step, st.local_uvmax = _oracle_callables(st, xc)
        st.scatter(T, u, v)
        dts = []
        for _ in range(STEPS):
            dts.append(st.step())
        full = st.gather().numpy()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), T=full, dts=np.asarray(dts), rows=np.asarray([st.slab.lo, st.slab.hi]))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_slab_stencil_equals_single_domain(world, tmp_path):
    mp.get_context("spawn")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref, ref_dts = _reference()
    covered = []
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(d["T"], ref), f"rank {r}: gathered field differs from the single-domain oracle"
        assert np.array_equal(d["dts"], np.asarray(ref_dts)), "the CFL dt must be the global one on every rank"
        covered += list(range(int(d["rows"][0]), int(d["rows"][1])))
    assert covered == list(range(H))  # slabs tile the rows exactly once


def test_slab_and_member_partition():
    for Hh, world in [(37, 2), (37, 3), (8192, 8), (16, 8)]:
        slabs = [MG.Slab(Hh, world, r) for r in range(world)]
        assert slabs[0].lo == 0 and slabs[-1].hi == Hh and all(a.hi == b.lo for a, b in zip(slabs, slabs[1:]))
        assert not slabs[0].up and not slabs[-1].down and all(s.rows == s.hi - s.lo + int(s.up) + int(s.down) for s in slabs)
    with pytest.raises(ValueError):
        MG.Slab(5, 4, 0)
    got = sum((MG.shard_members(256, 8, r) for r in range(8)), [])
    assert got == list(range(256)) and len(MG.shard_members(10, 4, 3)) in (2, 3)


def test_slab_stencil_refuses_cpu_without_an_update():
    with pytest.raises(RuntimeError):
        MG.SlabStencil(16, 16, np.linspace(0, 4, 16), np.linspace(0, 1, 16), 0, 1, "cpu")
