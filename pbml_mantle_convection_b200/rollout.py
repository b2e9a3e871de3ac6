"""Device-resident rollouts: B independent members (varying Ra / initial T) advanced together.

This is the GAIA-free form of the reference's hot loop (advect_wi_gaia.py:583-668 around
`ts_net(...)`, i.e. `TS.forward`'s loop pytorch_networks_convae.py:377-473): T, u, v and dt
never leave the device; a whole block of time steps is one CUDA graph.  Each member keeps its
own CFL time step (the reference runs one member per process; ADNet's batch-global amax,
:556, is not inherited here -- use `per_member_dt=False` for that behaviour).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .engine import Grid, RolloutState


def synthetic_grid(H, W):
    """Cell-centre coordinates of the box [0,4]x[0,1]; at 128x506 this is
    the reference's GAIA grid (prepare_gaia_ini.py:24-25).  float64 numpy [H,W] pair."""
    x = np.empty(W)
    y = np.empty(H)
    x[0], x[-1] = 0.0, 4.0
    x[1:-1] = (np.arange(1, W - 1) - 0.5) * 4.0 / (W - 2)
    y[0], y[-1] = 0.0, 1.0
    y[1:-1] = (np.arange(1, H - 1) - 0.5) / (H - 2)
    return np.broadcast_to(x[None, :], (H, W)).copy(), np.broadcast_to(y[:, None], (H, W)).copy()


def synthetic_T0(H, W, seed=1):
    """Conductive profile + 1 % noise with the wall rows/columns enforced."""
    _, yc = synthetic_grid(H, W)
    T = (1.0 - yc) + 0.01 * np.random.default_rng(seed).random((H, W))
    T[0, :], T[-1, :] = 1.0, 0.0
    T[:, 0], T[:, -1] = T[:, 1], T[:, -2]
    return T


class EnsembleRollout:
    """B members on one GPU.  `params`: list of (RaQ, gamma=fkt, beta=fkp), one per member."""

    def __init__(self, net, H, W, params, device, xc=None, yc=None, cn_max=0.99, per_member_dt=True, max_steps=4096):
        self.net, self.device = net, torch.device(device)
        self.B, self.H, self.W = len(params), H, W
        if xc is None:
            xc, yc = synthetic_grid(H, W)
        xc_t, yc_t = torch.as_tensor(xc, dtype=torch.float64), torch.as_tensor(yc, dtype=torch.float64)
        self.grid = Grid(xc_t, yc_t, yc_t, self.device)
        if not self.grid.separable:
            raise NotImplementedError("EnsembleRollout needs a separable (tensor-product) grid; use TS for general fields")
        self.members = ops.make_members(params, self.device)
        self.engine = net._engine(self.device)
        self.state = RolloutState(self.grid, self.members, self.B, 2, max_steps, cn_max, per_member_dt, net.p_pred,
                                  self.device)
        self.max_steps = max_steps
        self.n_done = 0
        self.time = torch.zeros(self.B, dtype=torch.float64, device=self.device)
        self._graphs = {}
        self._stream = torch.cuda.Stream(self.device)

    # ------------------------------------------------------------------ state
    def set_T(self, T0):
        """T0: [B,H,W] (numpy or tensor, any float dtype)."""
        T0 = torch.as_tensor(T0).to(self.device, torch.float32).reshape(self.B, self.H, self.W)
        self.state.T_seq[self.n_done % 2].copy_(T0)

    @property
    def T(self):
        return self.state.T_seq[self.n_done % 2]

    # ------------------------------------------------------------------ stepping
    def step(self, n=1):
        """Enqueue n time steps directly (no graph)."""
        done = 0
        self.engine = self.net._engine(self.device)  # picks up a changed net.conv_impl
        while done < n:
            # dt history is written relative to first_step (rows first-1 .. first-1+k-1 of max_steps + 1); restart the
            # index every call
            first = 1 + (self.n_done % 2)
            k = min(n - done, self.max_steps)
            self.engine.rollout(self.state, first, k)
            self.time += self.state.dt_seq[first - 1:first - 1 + k].sum(0)
            self.n_done += k
            done += k

    def graph(self, steps_per_graph):
        """Capture (once per (k, slot parity)) and return the CUDA graph of k time steps starting
        from the current T slot.  Replaying it advances the state by k steps."""
        k = int(steps_per_graph)
        if k < 1 or k > self.max_steps - 1:
            raise ValueError("steps_per_graph out of range")
        self.engine = self.net._engine(self.device)  # picks up a changed net.conv_impl
        par = self.n_done % 2
        # the graph bakes in packed-weight and workspace addresses and the conv implementation: key on them too
        key = (k, par) + self.engine.graph_key(self.B, self.H, self.W)
        if key not in self._graphs:
            saved = self.state.T_seq.clone()
            self.engine.rollout(self.state, 1 + par, 1)  # eager warm-up: module load + func attributes outside capture
            self.state.T_seq.copy_(saved)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.engine.rollout(self.state, 1 + par, k)
            self._graphs[key] = g
        return self._graphs[key]

    def run(self, n_steps, steps_per_graph=None, track_time=True):
        """Advance n_steps.  With steps_per_graph=k (divides n_steps) the work is k-step CUDA-graph replays.
        track_time=False skips the (tiny, torch-side) accumulation of simulated time."""
        if not steps_per_graph:
            return self.step(n_steps)
        k = int(steps_per_graph)
        if n_steps % k:
            raise ValueError("n_steps must be a multiple of steps_per_graph")
        for _ in range(n_steps // k):
            par = self.n_done % 2
            self.graph(k).replay()
            if track_time:
                self.time += self.state.dt_seq[par:par + k].sum(0)
            self.n_done += k

    # ------------------------------------------------------------------ outputs
    def fields(self):
        """(u, v, p, V) of the last step, float32 [B,H,W]."""
        s = self.state
        return s.u, s.v, s.p, s.V

    def last_dt(self):
        return self.state.dt_seq

    def diagnostics(self):
        """mean-T [B], profile T(y) [B,H], gradient dT/dy [B,H-1] as float64 tensors
        (advect_wi_gaia.py:547,647; load_advection_results notebook :322-323, r -> the run's y)."""
        mean, prof = ops.diagnostics(self.T.contiguous())
        y = self.grid.y1d64.to(self.device)
        dprof = (prof[:, 1:] - prof[:, :-1]) / (y[1:] - y[:-1])[None]
        return mean, prof, dprof
