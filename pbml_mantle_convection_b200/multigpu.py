"""Multi-GPU partitioning of the time-stepping path (BASELINE configs 4 and 5).

The reference has no multi-GPU inference at all (`advect_wi_gaia.py:422-425` picks one device, batch 1); its only
parallel code is DDP training (`multigpu.py:16-34,69`: one process per GPU, NCCL).  The same process model is
kept here -- one process per GPU, `torch.distributed` for the plumbing -- for the two ways the path shards:

* ensemble (config 4): members are independent, so they are dealt out to the ranks and there is NO data-path
  collective (`shard_members`; `bench.py --gpus N` runs it);
* one large grid (config 5): `SlabStencil` splits the rows of the advection-diffusion update (`ADNet.forward`,
  pytorch_networks_convae.py:522-568) into contiguous slabs.  Per step there is one real exchange: the CFL time
  step is ONE scalar for the whole grid (:554-559), so max|u|,|v| is all-reduced (MAX), and each slab needs the
  neighbours' boundary row of T (only T: u and v are used at the cell centre, :547-563) -- one row each way by
  send/recv (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The local update of a slab is the unmodified single-GPU kernel (`pbmc_advect_diffuse`): a slab is stored with
one ghost row where it has a neighbour, the kernel's "wall" writes land in those ghost rows and are overwritten
by the exchange, and the y-coefficients are slices of the global ones (so the spacing across a slab boundary is
the true one).  On CPU there is no kernel: the host logic takes the local update as a callable so that the
world-size-2 gloo tests can drive it with the numpy oracle.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_members(n_members: int, world: int, rank: int):
    """Contiguous block of ensemble members owned by `rank` (sizes differ by at most one)."""
    lo, hi = rank * n_members // world, (rank + 1) * n_members // world
    return list(range(lo, hi))


class Slab:
    """Row range of one rank: owns global rows [lo, hi); stores [l0, l1) = owned rows plus ghosts."""

    def __init__(self, H: int, world: int, rank: int):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad rank/world")
        if H // world < 2:
            raise ValueError(f"{H} rows cannot be split into {world} slabs of >= 2 rows")
        self.H, self.world, self.rank = H, world, rank
        self.lo, self.hi = rank * H // world, (rank + 1) * H // world
        self.up, self.down = rank > 0, rank < world - 1  # neighbour towards row 0 / towards row H-1
        self.l0, self.l1 = self.lo - int(self.up), self.hi + int(self.down)
        self.rows = self.l1 - self.l0

    def take(self, full):
        """Local view (with ghosts) of a [..., H, W] array that holds the whole grid."""
        return full[..., self.l0:self.l1, :]

    def owned(self, local):
        """Owned rows of a local [..., rows, W] array."""
        a = int(self.up)
        return local[..., a:a + (self.hi - self.lo), :]


def cfl_dt(uvmax: float, dx_min: float, cn_max: float) -> float:
    """pytorch_networks_convae.py:557-559 (the diffusive limit reduces to dx_min^2 / 4)."""
    dt_adv = 0.5 * cn_max * dx_min / uvmax
    dt_dif = 0.5 * ((dx_min * dx_min) ** 2) / (dx_min ** 2 + dx_min ** 2)
    return min(dt_adv, dt_dif)


class SlabStencil:
    """Advection-diffusion + CFL time step on one slab of a row-decomposed grid.

    x1d [W], y1d [H]: cell-centre coordinates of the (separable) global grid.  T, u, v are LOCAL float32 arrays
    [1, rows, W] (ghost rows included).  `local_step(T, u, v, uvmax_bits) -> (T_new, dt)` defaults to the CUDA
    kernel; `local_uvmax(u, v) -> int32 tensor [1]` likewise.  `group` is the process group (None = default)."""

    def __init__(self, H, W, x1d, y1d, rank, world, device, raq=0.0, cn_max=0.99, local_step=None, local_uvmax=None,
                 group=None):
        self.slab = Slab(H, world, rank)
        self.H, self.W, self.device = H, W, torch.device(device)
        self.raq, self.cn_max, self.group = float(raq), float(cn_max), group
        x = np.asarray(x1d, dtype=np.float64).copy()
        y = np.asarray(y1d, dtype=np.float64).copy()
        x[0], x[-1] = 0.0, 4.0  # forced wall coordinates, :532-535
        y[0], y[-1] = 0.0, 1.0
        self.x1d, self.y1d = x, y
        self.dx_min = float((x[1:-1] - x[:-2]).min())  # :555 (x spacing only)
        self.y_loc = y[self.slab.l0:self.slab.l1]
        self.local_step, self.local_uvmax = local_step, local_uvmax
        self.T = self.u = self.v = None
        self.n_steps, self.time, self.last_dt = 0, 0.0, None
        if self.device.type == "cuda":
            self._init_cuda()
        elif local_step is None or local_uvmax is None:
            raise RuntimeError("SlabStencil has no CPU implementation of the update: pass local_step/local_uvmax "
                               "(the tests use the numpy oracle) or run on CUDA")

    # ------------------------------------------------------------------ CUDA plumbing (C ABI kernels)
    def _init_cuda(self):
        from . import ops

        dev = self.device
        self._ops = ops
        xg = torch.tensor(self.x1d, dtype=torch.float64, device=dev)
        yg = torch.tensor(self.y1d, dtype=torch.float64, device=dev)
        self.xcoef = ops.stencil_coefs(xg, 0.0, 4.0)
        ycoef_g = ops.stencil_coefs(yg, 0.0, 1.0)  # [3, H], global
        self.ycoef = ycoef_g[:, self.slab.l0:self.slab.l1].contiguous()  # spacing across slab boundaries is the true one
        self.members = ops.make_members([(self.raq, 1.0, 1.0)], dev)
        self._T_out = None
        self._dt = torch.zeros(1, dtype=torch.float64, device=dev)
        if self.local_uvmax is None:
            self.local_uvmax = lambda u, v: ops.uvmax_reduce(u, v, batch_global=True)
        if self.local_step is None:
            def step(T, u, v, uvmax):
                if self._T_out is None or self._T_out.shape != T.shape:
                    self._T_out = torch.empty_like(T)
                out, dt, _ = ops.advect_diffuse(T, u, v, self.xcoef, self.ycoef, self.members, uvmax, self.dx_min,
                                                self.cn_max, per_member_dt=False, T_out=self._T_out, dt_out=self._dt)
                self._T_out = T  # ping-pong: the old T becomes the next output buffer
                return out, dt
            self.local_step = step

    # ------------------------------------------------------------------ state
    def set_local(self, T, u, v):
        f = lambda a: torch.as_tensor(a).to(self.device, torch.float32).reshape(1, self.slab.rows, self.W).contiguous().clone()
        self.T, self.u, self.v = f(T), f(u), f(v)

    def scatter(self, T_full, u_full, v_full):
        """Every rank holds the whole [H, W] fields (tests, small grids): keep the local slab."""
        s = self.slab
        self.set_local(s.take(np.asarray(T_full)), s.take(np.asarray(u_full)), s.take(np.asarray(v_full)))

    def set_velocity(self, u, v):
        self.u = torch.as_tensor(u).to(self.device, torch.float32).reshape(1, self.slab.rows, self.W).contiguous()
        self.v = torch.as_tensor(v).to(self.device, torch.float32).reshape(1, self.slab.rows, self.W).contiguous()

    # ------------------------------------------------------------------ one time step
    def global_uvmax(self):
        """max|u|,|v| over the global interior as float bits (non-negative floats order like their bit patterns,
        so an integer MAX all-reduce is the float MAX); stays on the device."""
        bits = self.local_uvmax(self.u, self.v)
        if self.slab.world > 1:
            dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=self.group)
        return bits

    def exchange_halo(self, T):
        """Send the first / last OWNED row to the neighbours, receive their rows into the ghost rows."""
        s = self.slab
        if s.world == 1:
            return
        ops_, keep = [], []
        if s.up:
            send = T[0, 1].contiguous()
            recv = torch.empty_like(send)
            ops_ += [dist.P2POp(dist.isend, send, s.rank - 1, self.group), dist.P2POp(dist.irecv, recv, s.rank - 1, self.group)]
            keep.append((0, recv))
        if s.down:
            send = T[0, s.rows - 2].contiguous()
            recv = torch.empty_like(send)
            ops_ += [dist.P2POp(dist.isend, send, s.rank + 1, self.group), dist.P2POp(dist.irecv, recv, s.rank + 1, self.group)]
            keep.append((s.rows - 1, recv))
        for w in dist.batch_isend_irecv(ops_):
            w.wait()
        for row, buf in keep:
            T[0, row].copy_(buf)

    def step(self, n=1):
        """n time steps; returns the last dt (device tensor on CUDA, float on CPU)."""
        dt = None
        for _ in range(n):
            bits = self.global_uvmax()
            T_new, dt = self.local_step(self.T, self.u, self.v, bits)
            self.exchange_halo(T_new)
            self.T = T_new
            self.n_steps += 1
        self.last_dt = dt
        return dt

    # ------------------------------------------------------------------ outputs
    def gather(self):
        """Whole-grid T [H, W] on every rank (all_gather of the owned rows; diagnostics / tests)."""
        own = self.slab.owned(self.T)[0].contiguous()
        if self.slab.world == 1:
            return own
        sizes = [Slab(self.H, self.slab.world, r) for r in range(self.slab.world)]
        nmax = max(s.hi - s.lo for s in sizes)  # all_gather needs equal shapes: pad to the largest slab
        mine = torch.zeros(nmax, self.W, dtype=own.dtype, device=own.device)
        mine[:own.shape[0]] = own
        bufs = [torch.empty_like(mine) for _ in sizes]
        dist.all_gather(bufs, mine, group=self.group)
        return torch.cat([b[:s.hi - s.lo] for b, s in zip(bufs, sizes)], 0)
