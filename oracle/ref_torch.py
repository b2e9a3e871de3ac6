"""ATen-CPU restatement of the reference rollout (TEST INFRASTRUCTURE ONLY).

Same algorithm as `oracle/ref_numpy.py`, written with the stock PyTorch CPU operators the
reference itself calls (F.conv2d, F.group_norm, F.gelu, F.avg_pool2d, F.interpolate), so
it is (a) multi-threaded and fast enough to be the CPU baseline `bench.py` times
("kind": "port"), and (b) usable at 512^2 in seconds.  It is functional (weights come in
as a dict with the reference's state_dict keys) and grid-size agnostic -- the reference's
hard-coded 128x506 (`pytorch_networks_convae.py:414-417`, `:1222-1229`) is not inherited.
Parity pin: tests/golden/*.npz (see oracle/__init__.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .ref_numpy import (NetSpec, SC_FKP, SC_FKT, SC_MUL, SC_RAQ, nondim_params,  # noqa: F401
                        sym_h_count)

_PAD = {"zeros": "constant", "constant": "constant", "replicate": "replicate", "reflect": "reflect"}


def _conv_same(x, w, b, mode):
    p = w.shape[-1] // 2
    if mode in ("zeros", "constant"):
        return F.conv2d(x, w, b, padding=p)
    return F.conv2d(F.pad(x, (p, p, p, p), mode=_PAD[mode]), w, b)


def _sym(w, c_out):
    # symmetric_layers_torch.py:118-138 with {'h': h, 'v': 0, 'hv': 0}
    h = sym_h_count(c_out)
    return w if h == 0 else torch.cat([w, torch.flip(w[: h // 2], (3,))], 0)


def prepare_weights(sd, spec: NetSpec, dtype=torch.float64):
    """Expand mirrored filters ONCE (the reference re-builds them every call)."""
    out = {}
    for k, v in sd.items():
        t = torch.as_tensor(v).to(dtype)
        if spec.use_symm and k.endswith("layers.0.weight"):
            t = _sym(t, spec.c_h)
        out[k] = t.contiguous()
    return out


def _fluid_layer(x, W, pre, C, spec):
    # FluidLayer.forward, pytorch_networks_convae.py:790-799
    y = _conv_same(x, W[pre + "layers.0.weight"], W[pre + "layers.0.bias"], spec.r_p)
    y = F.group_norm(y, int(C / min(4, C)), W[pre + "layers.1.weight"], W[pre + "layers.1.bias"], 1e-5)
    return F.gelu(y)


def net_forward(W, spec: NetSpec, inp):
    """NewFluidNet.forward (pytorch_networks_convae.py:1315-1388), loss_type='curl'."""
    H, Wd = inp.shape[-2:]
    C = spec.c_h
    x_in = _fluid_layer(inp, W, "conv.0.", C, spec)
    feats = []
    pooled = x_in
    for l in range(spec.levels):
        if l > 0:
            pooled = F.avg_pool2d(pooled, 2, 2)  # == pool^l(x_in), :1321-1322
        y1 = pooled
        for r in range(spec.repeats):
            y1 = _fluid_layer(y1, W, f"convs.{l}.{r}.", C, spec)
        if l > 0:
            y1 = F.interpolate(y1, size=(H, Wd), mode="bicubic")
        feats.append(y1)
    y = torch.cat(feats + [inp], 1)
    y = _conv_same(y, W["conv.1.weight"], W["conv.1.bias"], spec.r_p)
    y = F.gelu(F.group_norm(y, int(C / 4), W["gn.0.weight"], W["gn.0.bias"], 1e-5))
    y = F.gelu(_conv_same(y, W["conv.2.weight"], W["conv.2.bias"], spec.r_p))
    y = _conv_same(y, W["conv.3.weight"], W["conv.3.bias"], spec.r_p)
    y = y - y.mean(dim=(2, 3), keepdim=True)
    a = y[:, 0] * spec.a_bound
    p = y[:, 1] if spec.p_pred else None
    u = torch.zeros_like(a)
    v = torch.zeros_like(a)
    u[:, 1:-1, 1:-1] = 0.5 * (a[:, 2:, 1:-1] - a[:, :-2, 1:-1])
    v[:, 1:-1, 1:-1] = -0.5 * (a[:, 1:-1, 2:] - a[:, 1:-1, :-2])
    for f in (u, v):
        f[:, 0, 1:-1] = f[:, 1, 1:-1]
        f[:, -1, 1:-1] = f[:, -2, 1:-1]
        f[:, :, 0] = f[:, :, 1]
        f[:, :, -1] = f[:, :, -2]
    u[:, :, 0] = -u[:, :, 1]
    u[:, :, -1] = -u[:, :, -2]
    v[:, 0, :] = -v[:, 1, :]
    v[:, -1, :] = -v[:, -2, :]
    for f in (u, v):
        f[:, 0, 0] = 0
        f[:, 0, -1] = 0
        f[:, -1, 0] = 0
        f[:, -1, -1] = 0
    return u, v, p


def build_input(T, xc, yc, raq, fkt, fkp):
    """pytorch_networks_convae.py:379-407.  T [B,1,H,W]; xc,yc [H,W]; scalars python floats."""
    import math

    raq_nd, fkt_nd, fkp_nd = nondim_params(raq, fkt, fkp)
    V = torch.exp(math.log(fkt) * (0.0 - T) + math.log(fkp) * (1.0 - yc)[None, None])
    V = torch.clip(V, 1e-8, 1.0)
    one = torch.ones_like(T)
    inp = torch.cat(
        [(xc / 4.0)[None, None].expand_as(T), (yc / 4.0)[None, None].expand_as(T), torch.log10(V) / 8,
         one * float(raq_nd), one * float(fkt_nd), one * float(fkp_nd), T], 1)
    return inp, V


def adnet(u, v, T, raq, xc, yc, CN_max, dt=None):
    """ADNet.forward, pytorch_networks_convae.py:522-568, pure slicing.  u,v,T [B,H,W]."""
    xc = xc.clone()
    yc = yc.clone()
    xc[:, 0], xc[:, -1] = 0.0, 4.0
    yc[0, :], yc[-1, :] = 0.0, 1.0
    ui, vi = u[:, 1:-1, 1:-1], v[:, 1:-1, 1:-1]
    dx_l = (xc[1:-1, 1:-1] - xc[1:-1, :-2])[None]
    dx_r = (xc[1:-1, 2:] - xc[1:-1, 1:-1])[None]
    dy_t = (yc[1:-1, 1:-1] - yc[:-2, 1:-1])[None]
    dy_b = (yc[2:, 1:-1] - yc[1:-1, 1:-1])[None]
    Tc = T[:, 1:-1, 1:-1]
    dT_l = Tc - T[:, 1:-1, :-2]
    dT_r = T[:, 1:-1, 2:] - Tc
    dT_t = Tc - T[:, :-2, 1:-1]
    dT_b = T[:, 2:, 1:-1] - Tc
    dT_dx = (dT_l / dx_l) * (ui > 0) + (dT_r / dx_r) * (ui < 0)
    dT_dy = (dT_t / dy_t) * (vi > 0) + (dT_b / dy_b) * (vi < 0)
    lap = (dT_r / dx_r - dT_l / dx_l) / (0.5 * dx_r + 0.5 * dx_l) + (dT_b / dy_b - dT_t / dy_t) / (
        0.5 * dy_b + 0.5 * dy_t)
    if dt is None:
        dx_min = dx_l.min()
        uv = torch.maximum(ui.abs().max(), vi.abs().max())
        dt = torch.minimum(0.5 * CN_max * dx_min / uv, 0.5 * ((dx_min * dx_min) ** 2) / (dx_min**2 + dx_min**2))
    Tn = Tc + dt * (-ui * dT_dx - vi * dT_dy + lap + raq)
    out = F.pad(Tn[:, None], (1, 1, 1, 1), mode="replicate")[:, 0]
    out[:, 0, :] = 1.0
    out[:, -1, :] = 0.0
    return out, dt


def ts_step(W, spec, T, xc, yc, raq, fkt, fkp, CN_max=0.99):
    """One TS iteration (pytorch_networks_convae.py:377-473).  T [B,1,H,W]."""
    import math

    inp, V = build_input(T, xc, yc, raq, fkt, fkp)
    u, v, p = net_forward(W, spec, inp)
    s = math.exp(raq / 10 * SC_RAQ + math.log(fkt) * SC_FKT + math.log(fkp) * SC_FKP) * SC_MUL
    u = u * s
    v = v * s
    Tn, dt = adnet(u, v, T[:, 0], raq, xc, yc, CN_max)
    Tn[:, 0, :] = 1.0
    Tn[:, -1, :] = 0.0
    Tn[:, :, 0] = Tn[:, :, 1]
    Tn[:, :, -1] = Tn[:, :, -2]
    return Tn[:, None], dt, u, v, p, V


@torch.no_grad()
def rollout(W, spec, T0, xc, yc, raq, fkt, fkp, n_steps, CN_max=0.99, keep=()):
    T = T0
    dts, snaps = [], {}
    u = v = p = V = None
    for i in range(1, n_steps + 1):
        T, dt, u, v, p, V = ts_step(W, spec, T, xc, yc, raq, fkt, fkp, CN_max)
        dts.append(float(dt))
        if i in keep:
            snaps[i] = T.clone()
    return T, dts, u, v, p, V, snaps
