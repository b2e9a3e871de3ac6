"""Slab decomposition on the GPU.  One GPU is available to the test box, so the
two ranks are emulated in ONE process: two `SlabStencil` objects with the real CUDA local update (ghost rows,
sliced y-coefficients, kernel wall writes landing in ghosts), the all-reduce replaced by a max over the two
local reductions and the send/recv by direct row copies.  The real NCCL path is exercised by
`bench.py --workload slab --gpus N` / tools/slab_check.py on a multi-GPU box and by the gloo tests on CPU."""
import numpy as np
import pytest
import torch

from oracle import ref_numpy as RN
from pbml_mantle_convection_b200 import multigpu as MG
from pbml_mantle_convection_b200 import ops
from pbml_mantle_convection_b200.engine import Grid

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fields(H, W, seed=5):
    xc, yc = RN.synthetic_grid(H, W)
    T = RN.synthetic_T0(H, W, seed=seed).astype(np.float32)
    psi = np.sin(np.pi * xc / 4 * 3) * np.sin(np.pi * yc)
    u = (np.gradient(psi, axis=0) * 3e3).astype(np.float32)
    v = (-np.gradient(psi, axis=1) * 3e3).astype(np.float32)
    return xc, yc, T, u, v


def _single_gpu(xc, yc, T, u, v, steps, raq):
    g = Grid(torch.tensor(xc), torch.tensor(yc), torch.tensor(yc), DEV)
    members = ops.make_members([(raq, 1.0, 1.0)], DEV)
    Tc = torch.tensor(T[None], device=DEV)
    ud, vd = torch.tensor(u[None], device=DEV), torch.tensor(v[None], device=DEV)
    uv = ops.uvmax_reduce(ud, vd, batch_global=True)
    dts = []
    for _ in range(steps):
        Tc, dt, _ = ops.advect_diffuse(Tc, ud, vd, g.xcoef, g.ycoef, members, uv, g.dx_min, 0.99, per_member_dt=False)
        dts.append(float(dt[0]))
    return Tc[0].cpu().numpy(), dts


@pytest.mark.parametrize("H,W,world", [(64, 128, 2), (67, 100, 3), (130, 64, 4)])
def test_emulated_slabs_equal_single_gpu(H, W, world):
    xc, yc, T, u, v = _fields(H, W)
    raq, steps = 2.5, 5
    ref, ref_dts = _single_gpu(xc, yc, T, u, v, steps, raq)
    sts = [MG.SlabStencil(H, W, xc[0], yc[:, 0], r, world, DEV, raq=raq, cn_max=0.99) for r in range(world)]
    for st in sts:
        st.scatter(T, u, v)
    dts = []
    for _ in range(steps):
        bits = torch.stack([st.local_uvmax(st.u, st.v) for st in sts]).max(0).values  # the all-reduce(MAX)
        new = []
        for st in sts:
            Tn, dt = st.local_step(st.T, st.u, st.v, bits)
            new.append(Tn)
        for r, st in enumerate(sts):  # the halo exchange
            if st.slab.up:
                new[r][0, 0].copy_(new[r - 1][0, sts[r - 1].slab.rows - 2])
            if st.slab.down:
                new[r][0, st.slab.rows - 1].copy_(new[r + 1][0, 1])
        for st, Tn in zip(sts, new):
            st.T = Tn
        dts.append(float(dt[0]))
    got = np.concatenate([st.slab.owned(st.T)[0].cpu().numpy() for st in sts], 0)
    assert np.array_equal(got, ref), np.abs(got - ref).max()
    assert dts == ref_dts


def test_world_one_slab_is_the_plain_kernel():
    H, W = 48, 96
    xc, yc, T, u, v = _fields(H, W, seed=9)
    ref, ref_dts = _single_gpu(xc, yc, T, u, v, 3, 1.0)
    st = MG.SlabStencil(H, W, xc[0], yc[:, 0], 0, 1, DEV, raq=1.0)
    st.scatter(T, u, v)
    st.step(3)
    assert np.array_equal(st.gather().cpu().numpy(), ref) and float(st.last_dt[0]) == ref_dts[-1]


@pytest.mark.parametrize("H,W,world", [(64, 128, 2), (71, 256, 3)])
def test_fused_halo_push_emulated(H, W, world):
    """`pbmc_advect_diffuse_slab` with peer ghost-row pointers: the first / last owned row of the output is also
    stored into the neighbour's ghost row by the update kernel itself (no exchange step).  The 'peers' are slabs on
    the same GPU here; on a multi-GPU box the same pointers come from symmetric memory (SlabStencil halo='p2p')."""
    xc, yc, T, u, v = _fields(H, W, seed=13)
    raq, steps = 1.5, 4
    ref, ref_dts = _single_gpu(xc, yc, T, u, v, steps, raq)
    sts = [MG.SlabStencil(H, W, xc[0], yc[:, 0], r, world, DEV, raq=raq, cn_max=0.99) for r in range(world)]
    for st in sts:
        st.scatter(T, u, v)
    bufs = [[st.T.clone(), torch.full_like(st.T, float("nan"))] for st in sts]  # ping-pong per slab; NaN = never written
    dt = torch.zeros(1, dtype=torch.float64, device=DEV)
    for k in range(steps):
        i, o = k % 2, (k + 1) % 2
        bits = torch.stack([st.local_uvmax(st.u, st.v) for st in sts]).max(0).values
        for r, st in enumerate(sts):
            s = st.slab
            up = bufs[r - 1][o][0, sts[r - 1].slab.rows - 1].data_ptr() if s.up else 0
            down = bufs[r + 1][o][0, 0].data_ptr() if s.down else 0
            ops.advect_diffuse_slab(bufs[r][i], st.u, st.v, st.xcoef, st.ycoef, st.members, bits, st.dx_min, 0.99, bufs[r][o], dt,
                                    s.up, s.down, up, down)
    got = np.concatenate([st.slab.owned(bufs[r][steps % 2])[0].cpu().numpy() for r, st in enumerate(sts)], 0)
    assert np.array_equal(got, ref), np.abs(got - ref).max()
    assert float(dt[0]) == ref_dts[-1]


@pytest.mark.parametrize("H,W,world", [(64, 128, 2), (71, 256, 3), (160, 128, 8)])
def test_flag_synchronised_slab_step_emulated(H, W, world):
    """`pbmc_advect_diffuse_slab_sync`: the global dt reduction and the halo exchange both happen inside the update kernel
    (8-byte tag|max slots + ghost-row stores in peer memory; include/pbmc.h).  One GPU here, so the ranks' kernels are
    launched strictly one after the other on one stream -- every kernel finds the tags it waits for already published
    (a kernel must never wait for a kernel that shares its GPU) -- and the 'peer' addresses are other slabs' buffers on
    the same device.  The two-parity slots make that order legal: rank 0 publishes step s + 1 before rank 1 has read
    step s.  Result: bit-identical to the single-domain kernel, dt included; a velocity change restarts the tags."""
    xc, yc, T, u, v = _fields(H, W, seed=17)
    raq, steps = 1.5, 5
    ref, ref_dts = _single_gpu(xc, yc, T, u, v, steps, raq)
    sts = [MG.SlabStencil(H, W, xc[0], yc[:, 0], r, world, DEV, raq=raq, cn_max=0.99) for r in range(world)]
    for st in sts:
        st.scatter(T, u, v)
    bufs = [[st.T.clone(), torch.full_like(st.T, float("nan"))] for st in sts]
    sync = torch.zeros(world, ops.SLAB_SYNC_BYTES // 8, dtype=torch.int64, device=DEV)
    ptrs = [sync[r].data_ptr() for r in range(world)]
    dts = [torch.zeros(1, dtype=torch.float64, device=DEV) for _ in range(world)]

    def run(n_steps, first):
        for r, st in enumerate(sts):
            ops.slab_sync_publish(st.u, st.v, ptrs[r], ptrs, r)
        for k in range(first, first + n_steps):
            i, o = k % 2, (k + 1) % 2
            for r, st in enumerate(sts):
                s = st.slab
                up = bufs[r - 1][o][0, sts[r - 1].slab.rows - 1].data_ptr() if s.up else 0
                down = bufs[r + 1][o][0, 0].data_ptr() if s.down else 0
                ops.advect_diffuse_slab_sync(bufs[r][i], st.u, st.v, st.xcoef, st.ycoef, st.members, st.dx_min, 0.99, bufs[r][o],
                                             dts[r], s.up, s.down, up, down, ptrs[r], ptrs, r)

    run(steps, 0)
    got = np.concatenate([st.slab.owned(bufs[r][steps % 2])[0].cpu().numpy() for r, st in enumerate(sts)], 0)
    assert np.array_equal(got, ref), np.abs(got - ref).max()
    assert all(float(d[0]) == ref_dts[-1] for d in dts)
    st_view = sync.cpu().numpy().view(np.uint32).reshape(world, -1)
    base = 2 * 16 * 2  # uint32 index of steps_done
    assert (st_view[:, base] == steps).all() and (st_view[:, base + 1] == 0).all() and (st_view[:, base + 2] == 0).all()
    assert (st_view[:, base + 3] == 0).all()  # nobody gave up waiting
    # new velocity field: blocks re-zeroed, tags restart, dt follows the new maximum
    for st in sts:
        st.u, st.v = st.u * 0.5, st.v * 0.25
    sync.zero_()
    start = steps % 2
    run(2, start)
    ud, vd = torch.tensor(u[None] * 0.5, device=DEV), torch.tensor(v[None] * 0.25, device=DEV)
    g = Grid(torch.tensor(xc), torch.tensor(yc), torch.tensor(yc), DEV)
    Tc = torch.tensor(ref[None], device=DEV)
    uv = ops.uvmax_reduce(ud, vd, batch_global=True)
    for _ in range(2):
        Tc, dt2, _ = ops.advect_diffuse(Tc, ud, vd, g.xcoef, g.ycoef, ops.make_members([(raq, 1.0, 1.0)], DEV), uv, g.dx_min, 0.99,
                                        per_member_dt=False)
    got2 = np.concatenate([st.slab.owned(bufs[r][(start + 2) % 2])[0].cpu().numpy() for r, st in enumerate(sts)], 0)
    assert np.array_equal(got2, Tc[0].cpu().numpy()) and float(dts[0][0]) == float(dt2[0])
