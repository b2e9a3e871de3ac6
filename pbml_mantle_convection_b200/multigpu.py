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
                 group=None, halo="nccl", dt_sync="flags"):
        self.slab = Slab(H, world, rank)
        self.H, self.W, self.device = H, W, torch.device(device)
        self.raq, self.cn_max, self.group = float(raq), float(cn_max), group
        if halo not in ("nccl", "p2p"):
            raise ValueError("halo must be 'nccl' (send/recv of one row each way) or 'p2p' (rows stored straight into the "
                             "neighbours' ghost rows by the update kernel, peer memory over NVLink)")
        self.halo = halo if world > 1 else "nccl"
        if self.halo == "p2p" and torch.device(device).type != "cuda":
            raise RuntimeError("halo='p2p' needs CUDA peer memory")
        if dt_sync not in ("flags", "nccl"):
            raise ValueError("dt_sync must be 'flags' (global max|u|,|v| exchanged through 8-byte slots in peer memory inside "
                             "the update kernel: no collective call per step) or 'nccl' (all_reduce(MAX) per step)")
        # the in-kernel reduction rides on the same peer-mapped memory as the fused halo push
        self.dt_sync = dt_sync if self.halo == "p2p" else "nccl"
        # second communicator for the halo rows (collective call: every rank constructs its SlabStencil)
        self.halo_group = dist.new_group() if (world > 1 and dist.is_initialized() and group is None and self.halo == "nccl") else group
        self._halo = None
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
        # the update kernel also reduces max|u|,|v| of its slab (same pass); reused for the next step's dt as long as
        # the velocities are unchanged -- a new velocity field (set_velocity / set_local) invalidates it
        self._uv_next = torch.zeros(1, dtype=torch.int32, device=dev)
        self._uv_valid = False
        if self.local_uvmax is None:
            self.local_uvmax = lambda u, v: ops.uvmax_reduce(u, v, batch_global=True)
        if self.halo == "p2p":
            self._init_p2p()
            return
        if self.local_step is None:
            def step(T, u, v, uvmax):
                if self._T_out is None or self._T_out.shape != T.shape:
                    self._T_out = torch.empty_like(T)
                self._uv_next.zero_()
                out, dt, _ = ops.advect_diffuse(T, u, v, self.xcoef, self.ycoef, self.members, uvmax, self.dx_min,
                                                self.cn_max, per_member_dt=False, T_out=self._T_out, dt_out=self._dt,
                                                uv_out=self._uv_next)
                self._uv_valid = True  # max|u|,|v| of these (unchanged) velocities came out of the same pass
                self._T_out = T  # ping-pong: the old T becomes the next output buffer
                return out, dt
            self.local_step = step


    def _init_p2p(self):
        """Fused halo exchange: both T buffers of every rank live in one symmetric-memory allocation, so each rank
        knows the device address of its neighbours' ghost rows and `pbmc_advect_diffuse_slab` stores the slab's first
        / last owned row there while it writes its own output (P2P stores over NVLink / NVSwitch).
        Cross-rank ordering: a rank starts step k+1 only after the dt all-reduce of step k+1, which every rank
        enqueues after its step-k update -- so all pushes of step k have landed before anyone reads them, and with
        two buffers nobody is still reading the buffer a neighbour pushes into."""
        import torch.distributed._symmetric_memory as symm_mem

        s, W, dev, ops = self.slab, self.W, self.device, self._ops
        sizes = [Slab(self.H, s.world, r) for r in range(s.world)]
        rows_max = max(z.rows for z in sizes)
        self._buf = symm_mem.empty((2, rows_max, W), dtype=torch.float32, device=dev)
        self._buf.zero_()
        grp = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(self._buf, group=grp)
        # peers' addresses of THIS tensor: the handle lists the base of the allocation the tensor lives in on every rank
        off = self._buf.data_ptr() - int(hdl.buffer_ptrs[s.rank])
        base = [int(p) + off for p in hdl.buffer_ptrs]
        row_bytes, slot_bytes = W * 4, rows_max * W * 4
        # neighbour ghost rows, per output slot k: up neighbour's LAST local row, down neighbour's local row 0
        self._peer_up = [base[s.rank - 1] + k * slot_bytes + (sizes[s.rank - 1].rows - 1) * row_bytes if s.up else 0 for k in (0, 1)]
        self._peer_down = [base[s.rank + 1] + k * slot_bytes if s.down else 0 for k in (0, 1)]
        self._slot = 0
        self._symm_handle = hdl
        if self.dt_sync == "flags":
            # one pbmc_slab_sync block per rank in a second symmetric allocation; every rank maps all of them
            self._sync = symm_mem.empty((ops.SLAB_SYNC_BYTES // 8,), dtype=torch.int64, device=dev)
            self._sync.zero_()
            self._sync_handle = symm_mem.rendezvous(self._sync, group=grp)
            off_s = self._sync.data_ptr() - int(self._sync_handle.buffer_ptrs[s.rank])
            self._sync_ptrs = [int(p) + off_s for p in self._sync_handle.buffer_ptrs]
            self._sync_arr = ops.peer_array(self._sync_ptrs)  # marshalled once: the step is one C call

            def step_flags(T, u, v, _uvmax_unused):
                k = self._slot
                out = self._buf[1 - k, :s.rows].unsqueeze(0)
                ops.advect_diffuse_slab_sync(T, u, v, self.xcoef, self.ycoef, self.members, self.dx_min, self.cn_max, out,
                                             self._dt, s.up, s.down, self._peer_up[1 - k], self._peer_down[1 - k],
                                             self._sync_ptrs[s.rank], self._sync_arr, s.rank)
                self._slot = 1 - k
                return out, self._dt

            self.local_step = step_flags
            return

        def step(T, u, v, uvmax):
            k = self._slot
            out = self._buf[1 - k, :s.rows].unsqueeze(0)
            self._uv_next.zero_()
            ops.advect_diffuse_slab(T, u, v, self.xcoef, self.ycoef, self.members, uvmax, self.dx_min, self.cn_max, out,
                                    self._dt, s.up, s.down, self._peer_up[1 - k], self._peer_down[1 - k], uv_out=self._uv_next)
            self._uv_valid = True
            self._slot = 1 - k
            return out, self._dt

        self.local_step = step

    # ------------------------------------------------------------------ state
    def set_local(self, T, u, v):
        f = lambda a: torch.as_tensor(a).to(self.device, torch.float32).reshape(1, self.slab.rows, self.W).contiguous().clone()
        self.T = f(T)
        if self.u is not None and self.halo == "p2p" and tuple(self.u.shape) == (1, self.slab.rows, self.W):
            # same buffers: a captured graph (it holds their addresses) stays valid
            self.u.copy_(torch.as_tensor(u).reshape(self.u.shape))
            self.v.copy_(torch.as_tensor(v).reshape(self.v.shape))
        else:
            self.u, self.v = f(u), f(v)
            self._graph = None
        self._uv_valid = False
        if self.halo == "p2p":
            # neighbours may still be pushing rows of an earlier run into this rank's buffers
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            self._slot = 0
            self._buf[0, :self.slab.rows].copy_(self.T[0])
            self.T = self._buf[0, :self.slab.rows].unsqueeze(0)
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            self._republish()  # flag mode: tags restart at 1 with this velocity field's maximum

    def _republish(self):
        """Flag mode: (re)start the tag sequence for the current velocity field.  Host-side barrier on both sides of the
        reset -- nobody may still be polling or publishing -- then every rank publishes tag 1 (pbmc.h protocol)."""
        if self.dt_sync != "flags" or self.halo != "p2p":
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._sync.zero_()  # (a captured graph stays valid: the kernels read tags and step counters from memory)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._ops.slab_sync_publish(self.u, self.v, self._sync_ptrs[self.slab.rank], self._sync_ptrs, self.slab.rank)

    def scatter(self, T_full, u_full, v_full):
        """Every rank holds the whole [H, W] fields (tests, small grids): keep the local slab."""
        s = self.slab
        self.set_local(s.take(np.asarray(T_full)), s.take(np.asarray(u_full)), s.take(np.asarray(v_full)))

    def set_velocity(self, u, v):
        self.u = torch.as_tensor(u).to(self.device, torch.float32).reshape(1, self.slab.rows, self.W).contiguous()
        self.v = torch.as_tensor(v).to(self.device, torch.float32).reshape(1, self.slab.rows, self.W).contiguous()
        self._uv_valid = False
        self._graph = None  # new velocity buffers: a captured graph holds the old addresses
        if self.halo == "p2p" and self.dt_sync == "flags":
            if self._slot != 0:
                raise RuntimeError("set_velocity in flag mode needs an even number of steps since set_local (ping-pong phase)")
            self._republish()

    # ------------------------------------------------------------------ one time step
    def global_uvmax(self):
        """max|u|,|v| over the global interior as float bits (non-negative floats order like their bit patterns,
        so an integer MAX all-reduce is the float MAX); stays on the device."""
        if getattr(self, "_uv_valid", False):
            bits = self._uv_next.clone()
        else:
            bits = self.local_uvmax(self.u, self.v)
        if self.slab.world > 1:
            dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=self.group)
        return bits

    def start_halo(self, T):
        """Post the sends of the first / last OWNED row and the receives of the neighbours' rows (asynchronous).
        The halo traffic runs on its own process group (`halo_group`) so that on NCCL it overlaps with the dt
        all-reduce of the next step instead of queueing behind it on one communicator."""
        s = self.slab
        self._halo = None
        if s.world == 1:
            return
        ops_, keep = [], []
        if s.up:
            send = T[0, 1].contiguous()
            recv = torch.empty_like(send)
            ops_ += [dist.P2POp(dist.isend, send, s.rank - 1, self.halo_group), dist.P2POp(dist.irecv, recv, s.rank - 1, self.halo_group)]
            keep.append((0, recv, send))
        if s.down:
            send = T[0, s.rows - 2].contiguous()
            recv = torch.empty_like(send)
            ops_ += [dist.P2POp(dist.isend, send, s.rank + 1, self.halo_group), dist.P2POp(dist.irecv, recv, s.rank + 1, self.halo_group)]
            keep.append((s.rows - 1, recv, send))
        self._halo = (dist.batch_isend_irecv(ops_), keep, T)

    def finish_halo(self):
        """Wait for the posted exchange and put the received rows into the ghost rows."""
        if getattr(self, "_halo", None) is None:
            return
        works, keep, T = self._halo
        for w in works:
            w.wait()
        for row, buf, _send in keep:
            T[0, row].copy_(buf)
        self._halo = None

    def exchange_halo(self, T):
        self.start_halo(T)
        self.finish_halo()

    def _step_once(self):
        if self.halo == "p2p" and self.dt_sync == "flags":
            # ONE launch: the kernel waits for the ranks' maxima, updates, pushes its boundary rows, publishes
            self.T, dt = self.local_step(self.T, self.u, self.v, None)
            return dt
        bits = self.global_uvmax()
        self.finish_halo()
        T_new, dt = self.local_step(self.T, self.u, self.v, bits)
        if self.halo == "nccl":
            self.start_halo(T_new)
        self.T = T_new
        return dt

    def step(self, n=1, use_graph=True):
        """n time steps; returns the last dt (device tensor on CUDA, float on CPU).
        Order per step: [dt all-reduce of this step || halo exchange posted by the previous step] -> local update ->
        post this step's halo exchange.  In p2p mode there is no exchange step at all, and pairs of steps
        (ping-pong period) are replayed as ONE CUDA graph -- all-reduce included -- so the host cost per step
        (five launches from Python otherwise) disappears; that is what strong scaling of a 0.1 ms step needs."""
        dt = None
        done = 0
        if use_graph and self.halo == "p2p" and n >= 4 and self._slot == 0 and (self._uv_valid or self.dt_sync == "flags"):
            if not getattr(self, "_warm", False):
                # the first launches of a kernel load its module: keep that outside the capture (two steps = one
                # ping-pong period, so the slot parity is unchanged)
                self._step_once()
                self._step_once()
                done += 2
                self._warm = True
            gs = max(2, int(getattr(self, "graph_steps", 2)) // 2 * 2)  # whole ping-pong periods per captured graph
            if getattr(self, "_graph", None) is None or getattr(self, "_graph_len", 0) != gs:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(gs):
                        dt = self._step_once()
                self._graph, self._graph_len = g, gs
            while n - done >= gs:
                self._graph.replay()
                done += gs
            dt = self._dt
            self.T = self._buf[0, :self.slab.rows].unsqueeze(0)
            self.n_steps += done
        for _ in range(n - done):
            dt = self._step_once()
            self.n_steps += 1
        self.finish_halo()  # leave a consistent state (gather / diagnostics / set_velocity may follow)
        self.last_dt = dt
        return dt

    def check_sync(self):
        """Flag mode: raise if a step kernel gave up waiting for a peer's publication (pbmc_slab_sync.failed)."""
        if getattr(self, "_sync", None) is None:
            return
        words = self._sync.view(torch.int32)
        failed = int(words[2 * 2 * 16 + 3].item()) & 0xFFFFFFFF
        if failed:
            missing = [r for r in range(self.slab.world) if failed >> r & 1]
            raise RuntimeError(f"slab step on rank {self.slab.rank}: no publication from rank(s) {missing} within 10 s "
                               f"(steps done {int(words[2 * 2 * 16].item())})")

    def close(self):
        """Drop the captured graph and the peer mappings (collective: every rank calls it) so that the process group can
        be destroyed cleanly afterwards."""
        self._graph = None
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
            self.check_sync()
        if self.slab.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        for name in ("_sync_handle", "_symm_handle", "_sync", "_buf"):
            if hasattr(self, name):
                delattr(self, name)
        self.T = None

    # ------------------------------------------------------------------ outputs
    def gather(self):
        """Whole-grid T [H, W] on every rank (all_gather of the owned rows; diagnostics / tests)."""
        self.check_sync()
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
