"""Index-level numpy restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it restates (paths relative to /root/reference).
No torch operator is used here: the point is to pin the *semantics* (padding, filter
mirroring, pooling floor, bicubic taps, BC ordering) independently of ATen.
Parity pin: tests/golden/*.npz, produced from the real reference by
tests/golden/make_golden.py; checked in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

try:  # scipy is in the image; keep a slow fallback so the oracle never silently changes
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

# velocity de-normalisation constants: pytorch_networks_convae.py:341-352, scaler.py:4-71
SC_RAQ, SC_FKT, SC_FKP, SC_MUL = 1.80167667, 0.4330392, -0.46052953, 5.0
# parameter normalisation: advect_wi_gaia.py:446-450
RAQ_LO, RAQ_HI = 0.12624371, 9.70723344
FKT_LO, FKT_HI = 6.00352841978384, 9.888820429862925
FKP_LO, FKP_HI = 0.005251646002323797, 1.9927988938926755


@dataclass
class NetSpec:
    """Constructor arguments of NewFluidNet that change the arithmetic
    (pytorch_networks_convae.py:1123-1145)."""

    levels: int = 6
    c_i: int = 7
    c_h: int = 16
    c_o: int = 2
    r_p: str = "replicate"
    loss_type: str = "curl"
    use_symm: bool = True
    a_bound: float = 10.0
    repeats: int = 4
    f: int = 3
    p_pred: bool = True


def nondim_params(raq, fkt, fkp):
    """advect_wi_gaia.py:443-450."""
    raq_nd = (raq - RAQ_LO) / (RAQ_HI - RAQ_LO)
    fkt_nd = (np.log10(fkt) - FKT_LO) / (FKT_HI - FKT_LO)
    fkp_nd = (np.log10(fkp) - FKP_LO) / (FKP_HI - FKP_LO)
    return raq_nd, fkt_nd, fkp_nd


def velocity_scaler(raq, fkt, fkp):
    """pytorch_networks_convae.py:343-350 (same constant as scaler.py:6-13)."""
    return np.exp(raq / 10 * SC_RAQ + np.log(fkt) * SC_FKT + np.log(fkp) * SC_FKP) * SC_MUL


def synthetic_grid(H, W, dtype=np.float64):
    """SURVEY.md section 8d synthetic box [0,4]x[0,1]; at 128x506 this is the GAIA grid
    (prepare_gaia_ini.py:24-25).  Row 0 is the hot bottom wall (y=0)."""
    x = np.empty(W, dtype=np.float64)
    y = np.empty(H, dtype=np.float64)
    x[0], x[-1] = 0.0, 4.0
    x[1:-1] = (np.arange(1, W - 1) - 0.5) * 4.0 / (W - 2)
    y[0], y[-1] = 0.0, 1.0
    y[1:-1] = (np.arange(1, H - 1) - 0.5) / (H - 2)
    xc = np.broadcast_to(x[None, :], (H, W)).astype(dtype).copy()
    yc = np.broadcast_to(y[:, None], (H, W)).astype(dtype).copy()
    return xc, yc


def apply_T_bcs(T):
    """pytorch_networks_convae.py:468-471 (in place on [...,H,W])."""
    T[..., 0, :] = 1.0
    T[..., -1, :] = 0.0
    T[..., :, 0] = T[..., :, 1]
    T[..., :, -1] = T[..., :, -2]
    return T


def synthetic_T0(H, W, seed=1, dtype=np.float64):
    """T0 = (1-y) + 0.01 U[0,1) with wall rows/cols enforced.
    Uses numpy's PCG64 so that GPU boxes regenerate identical fields without torch RNG."""
    _, yc = synthetic_grid(H, W)
    rng = np.random.default_rng(seed)
    T = (1.0 - yc) + 0.01 * rng.random((H, W))
    return apply_T_bcs(T).astype(dtype)


# ----------------------------------------------------------------------------- layers
def sym_h_count(c_out):
    """pytorch_networks_convae.py:755 -- number of h-mirrored filters."""
    return int(c_out / 4) if c_out > 4 else int(c_out / 2)


def expand_symmetric(w_unique, c_out):
    """symmetric_layers_torch.py:118-138 with symmetry={'h': h, 'v': 0, 'hv': 0}:
    full filter bank = [unique..., x-mirrored copies of the first h/2 unique filters]."""
    h = sym_h_count(c_out)
    if h == 0:
        return w_unique
    assert w_unique.shape[0] == c_out - h // 2
    return np.concatenate([w_unique, w_unique[: h // 2, :, :, ::-1]], axis=0)


_PAD = {"zeros": "constant", "constant": "constant", "replicate": "edge", "reflect": "reflect"}


def conv2d_same(x, w, b, mode):
    """k x k cross-correlation, stride 1, 'same' output, padding by `mode`
    (nn.Conv2d._conv_forward as used at symmetric_layers_torch.py:138 and
    pytorch_networks_convae.py:1263-1309).  x [B,Ci,H,W], w [Co,Ci,k,k]."""
    k = w.shape[-1]
    p = k // 2
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)), mode=_PAD[mode])
    return conv2d_valid(xp, w, b)


def conv2d_valid(xp, w, b=None):
    B, Ci, Hp, Wp = xp.shape
    Co, _, k, _ = w.shape
    H, W = Hp - k + 1, Wp - k + 1
    out = np.zeros((B, Co, H, W), dtype=xp.dtype)
    for dy in range(k):
        for dx in range(k):
            out += np.einsum("oi,bihw->bohw", w[:, :, dy, dx], xp[:, :, dy : dy + H, dx : dx + W])
    if b is not None:
        out += b.reshape(1, -1, 1, 1)
    return out


def group_norm(x, gamma, beta, groups, eps=1e-5):
    """torch.nn.GroupNorm (pytorch_networks_convae.py:788, :1279): biased variance."""
    B, C, H, W = x.shape
    xg = x.reshape(B, groups, -1)
    m = xg.mean(axis=2, keepdims=True)
    v = ((xg - m) ** 2).mean(axis=2, keepdims=True)
    y = ((xg - m) / np.sqrt(v + eps)).reshape(B, C, H, W)
    return y * gamma.reshape(1, C, 1, 1) + beta.reshape(1, C, 1, 1)


def gelu(x):
    """nn.GELU() exact-erf (pytorch_networks_convae.py:751)."""
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def avg_pool2(x):
    """nn.AvgPool2d((2,2), stride=2), floor on odd sizes (pytorch_networks_convae.py:1225)."""
    H2, W2 = x.shape[-2] // 2, x.shape[-1] // 2
    x = x[..., : 2 * H2, : 2 * W2]
    return 0.25 * (x[..., 0::2, 0::2] + x[..., 0::2, 1::2] + x[..., 1::2, 0::2] + x[..., 1::2, 1::2])


def _cubic_coeffs(t, A=-0.75):
    # taps at offsets -1, 0, +1, +2 (Keys kernel, A = -0.75)
    def c1(s):  # |s| <= 1
        return ((A + 2) * s - (A + 3)) * s * s + 1

    def c2(s):  # 1 < |s| < 2
        return ((A * s - 5 * A) * s + 8 * A) * s - 4 * A

    return np.stack([c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)], axis=0)


def _bicubic_axis(n_in, n_out):
    """Source taps/weights of nn.Upsample(size, mode='bicubic') (align_corners=False),
    pytorch_networks_convae.py:1227-1229: half-pixel source index (not clamped),
    4 taps clamped to the valid range."""
    scale = n_in / n_out
    src = scale * (np.arange(n_out) + 0.5) - 0.5
    i0 = np.floor(src).astype(np.int64)
    t = src - i0
    wts = _cubic_coeffs(t)  # [4, n_out]
    idx = np.stack([np.clip(i0 - 1 + k, 0, n_in - 1) for k in range(4)], axis=0)
    return idx, wts


def bicubic_upsample(x, size):
    H, W = size
    iy, wy = _bicubic_axis(x.shape[-2], H)
    ix, wx = _bicubic_axis(x.shape[-1], W)
    # rows first, then columns (separable; order only affects rounding)
    tmp = sum(wy[k][None, None, :, None] * x[:, :, iy[k], :] for k in range(4))
    out = sum(wx[k][None, None, None, :] * tmp[:, :, :, ix[k]] for k in range(4))
    return out.astype(x.dtype)


def fluid_layer(x, sd, prefix, c_out, spec: NetSpec):
    """FluidLayer.forward (pytorch_networks_convae.py:790-799), non-learned padding."""
    w = sd[prefix + "layers.0.weight"]
    if spec.use_symm:
        w = expand_symmetric(w, c_out)
    y = conv2d_same(x, w, sd[prefix + "layers.0.bias"], spec.r_p)
    y = group_norm(y, sd[prefix + "layers.1.weight"], sd[prefix + "layers.1.bias"], int(c_out / min(4, c_out)))
    return gelu(y)


def curl_head(y, spec: NetSpec):
    """pytorch_networks_convae.py:1343-1388: zero-mean, stream function -> (u, v), wall BCs."""
    y = y - y.mean(axis=(2, 3), keepdims=True)
    a = y[:, 0] * spec.a_bound  # [B,H,W]
    p = y[:, 1] if spec.p_pred else None
    B, H, W = a.shape
    u = np.zeros_like(a)
    v = np.zeros_like(a)
    u[:, 1:-1, 1:-1] = 0.5 * (a[:, 2:, 1:-1] - a[:, :-2, 1:-1])  # d/d(row), :1369
    v[:, 1:-1, 1:-1] = -0.5 * (a[:, 1:-1, 2:] - a[:, 1:-1, :-2])  # -d/d(col), :1370
    for f in (u, v):  # replicate pad, :1372/:1380
        f[:, 0, 1:-1] = f[:, 1, 1:-1]
        f[:, -1, 1:-1] = f[:, -2, 1:-1]
        f[:, :, 0] = f[:, :, 1]
        f[:, :, -1] = f[:, :, -2]
    u[:, :, 0] = -u[:, :, 1]
    u[:, :, -1] = -u[:, :, -2]
    v[:, 0, :] = -v[:, 1, :]
    v[:, -1, :] = -v[:, -2, :]
    for f in (u, v):
        f[:, 0, 0] = f[:, 0, -1] = f[:, -1, 0] = f[:, -1, -1] = 0.0
    return u, v, p


def newfluidnet_forward(sd, spec: NetSpec, inp):
    """NewFluidNet.forward (pytorch_networks_convae.py:1315-1388), r_p != 'learned'."""
    H, W = inp.shape[-2:]
    C = spec.c_h
    x_in = fluid_layer(inp, sd, "conv.0.", C, spec)
    feats = []
    for l in range(spec.levels):
        y1 = x_in
        for _ in range(l):
            y1 = avg_pool2(y1)
        for r in range(spec.repeats):
            y1 = fluid_layer(y1, sd, f"convs.{l}.{r}.", C, spec)
        if l > 0:
            y1 = bicubic_upsample(y1, (H, W))
        feats.append(y1)
    y = np.concatenate(feats + [inp], axis=1)
    y = conv2d_same(y, sd["conv.1.weight"], sd["conv.1.bias"], spec.r_p)
    y = gelu(group_norm(y, sd["gn.0.weight"], sd["gn.0.bias"], int(C / 4)))
    y = gelu(conv2d_same(y, sd["conv.2.weight"], sd["conv.2.bias"], spec.r_p))
    y = conv2d_same(y, sd["conv.3.weight"], sd["conv.3.bias"], spec.r_p)
    if spec.loss_type == "curl":
        return curl_head(y, spec)
    y = y - y.mean(axis=(2, 3), keepdims=True)  # :1343-1354
    return y[:, 0], y[:, 1], (y[:, 2:3] if spec.p_pred else None)


def _curl_uv(a):
    """Central-difference curl of a stream function with the wall BCs of :1356-1388 / :2044-2066 (a: [B,H,W])."""
    u = np.zeros_like(a)
    v = np.zeros_like(a)
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
        f[:, 0, 0] = f[:, 0, -1] = f[:, -1, 0] = f[:, -1, -1] = 0.0
    return u, v


def unet_channels(levels, c_h):
    """Channel bookkeeping of Unet.__init__ (pytorch_networks_convae.py:1866-1935): per-level widths of the
    down path (level l >= 1 has c_h * 2**(l-1) channels, level 0 has c_h) and of each up block's output."""
    down = [c_h] + [c_h * 2 ** (l - 1) for l in range(1, levels)]
    top = down[-1]
    up = []
    for _ in range(levels - 2, 0, -1):
        up.append(top // 2)
        top //= 2
    return down, up, top


def unet_forward(sd, spec: NetSpec, inp):
    """Unet.forward (pytorch_networks_convae.py:1985-2068).
    The time-stepper variant of the surrogate: the input is padded by 3 columns each side -- or, with r_p='learned', the
    first 9-region conv enlarges it by the same 3 columns itself (`bc_x=4, bc_y=1`, :1994-1996) --, goes down `levels`-1
    poolings with the width doubling from level 2 on, comes back up with skip concatenations, and the head removes
    the per-channel mean BEFORE the 3 columns are cropped again.  Returns (u, v, p, T) like the reference."""
    learned = spec.r_p == "learned"
    k, sym = spec.f, spec.use_symm

    def layer(x, prefix, c_out, bc_x=1, bc_y=1):
        if not learned:
            return fluid_layer(x, sd, prefix, c_out, spec)
        y = boundary_learned_conv(x, sd, prefix + "layers.0.", k, c_out, use_symm=sym, bc_x=bc_x, bc_y=bc_y)
        return gelu(group_norm(y, sd[prefix + "layers.1.weight"], sd[prefix + "layers.1.bias"], int(c_out / min(4, c_out))))

    def head_conv(x, idx, c_out):
        if learned:
            return boundary_learned_conv(x, sd, f"conv.{idx}.", k, c_out, use_symm=sym)
        return conv2d_same(x, sd[f"conv.{idx}.weight"], sd[f"conv.{idx}.bias"], spec.r_p)

    x = {0: inp if learned else np.pad(inp, ((0, 0), (0, 0), (0, 0), (3, 3)), mode=_PAD[spec.r_p])}  # :1990-1991
    down, up, top = unet_channels(spec.levels, spec.c_h)
    for r in range(spec.repeats):
        x[0] = layer(x[0], f"conv.{r}.", spec.c_h, bc_x=4 if (learned and r == 0) else 1)
    sizes = {0: x[0].shape[-2:]}
    for l in range(1, spec.levels):
        x[l] = avg_pool2(x[l - 1])
        sizes[l] = x[l].shape[-2:]
        for r in range(spec.repeats):
            x[l] = layer(x[l], f"convs.{l - 1}.{r}.", down[l])
    xu = x[spec.levels - 1]
    for l_i, l in enumerate(range(spec.levels - 2, 0, -1)):
        xu = np.concatenate([x[l], bicubic_upsample(xu, sizes[l])], axis=1)
        for r in range(spec.repeats):
            xu = layer(xu, f"upconvs.{l_i}.{r}.", up[l_i])
    y = np.concatenate([bicubic_upsample(xu, sizes[0]), x[0]], axis=1)
    R = spec.repeats
    y = head_conv(y, R, top)
    y = gelu(group_norm(y, sd["gn.0.weight"], sd["gn.0.bias"], int(top / 4)))
    y = gelu(head_conv(y, R + 1, top))
    y = head_conv(y, R + 2, spec.c_o)
    y = (y - y.mean(axis=(2, 3), keepdims=True))[..., 3:-3]  # :2025
    if spec.loss_type in ("mae", "mass"):  # :2027-2037
        return y[:, 0:1], y[:, 1:2], (y[:, 3:4] if spec.p_pred else None), y[:, 2:3]
    u, v = _curl_uv(y[:, 0] * spec.a_bound)  # :2039-2066
    return u, v, (y[:, 2] if spec.p_pred else None), np.clip(y[:, 1], 0.0, 1.5)


# ----------------------------------------------------------------------------- TS / ADNet
def build_input(T, xc, yc, ycc, raq, fkt, fkp):
    """TS.forward input build, pytorch_networks_convae.py:379-407.  T [B,1,H,W];
    xc,yc,ycc [H,W]; returns ([B,7,H,W], clipped viscosity V [B,1,H,W])."""
    raq_nd, fkt_nd, fkp_nd = nondim_params(raq, fkt, fkp)
    V = np.exp(np.log(fkt) * (0.0 - T) + np.log(fkp) * ((1.0 - ycc)[None, None] - 0.0))
    V = np.clip(V, 1e-8, 1.0)
    B = T.shape[0]
    one = np.ones_like(T)
    inp = np.concatenate(
        [
            np.broadcast_to((xc / 4.0)[None, None], T.shape),
            np.broadcast_to((yc / 4.0)[None, None], T.shape),
            np.log10(V) / 8,
            one * raq_nd,
            one * fkt_nd,
            one * fkp_nd,
            T,
        ],
        axis=1,
    ).astype(T.dtype)
    return inp, V


def adnet_forward(u, v, T, raq, xc, yc, CN_max, dt=None, y_walls=(True, True)):
    """ADNet.forward, pytorch_networks_convae.py:522-568.  u,v,T [B,H,W]; xc,yc [H,W].
    Returns (T_new [B,H,W] incl. ADNet's own wall rows, dt scalar).
    y_walls: which of the first / last rows are real walls.  The reference always has both; a row slab of a
    decomposed grid has a ghost row instead, whose coordinate is NOT forced and whose
    output value is left to the halo exchange."""
    xc = xc.copy()
    yc = yc.copy()
    xc[:, 0], xc[:, -1] = 0.0, 4.0  # :532-535
    if y_walls[0]:
        yc[0, :] = 0.0
    if y_walls[1]:
        yc[-1, :] = 1.0
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
    dT_dx = (dT_l / dx_l) * (ui > 0) + (dT_r / dx_r) * (ui < 0)  # :547
    dT_dy = (dT_t / dy_t) * (vi > 0) + (dT_b / dy_b) * (vi < 0)  # :548
    lap = (dT_r / dx_r - dT_l / dx_l) / (0.5 * dx_r + 0.5 * dx_l) + (dT_b / dy_b - dT_t / dy_t) / (
        0.5 * dy_b + 0.5 * dy_t
    )
    if dt is None:
        dx_min = dx_l.min()
        uv_mag = max(np.abs(ui).max(), np.abs(vi).max())
        dt_adv = 0.5 * CN_max * dx_min / uv_mag
        dt_dif = 0.5 * ((dx_min * dx_min) ** 2) / (dx_min**2 + dx_min**2)
        dt = min(dt_adv, dt_dif)
    Tn = Tc + dt * (-ui * dT_dx - vi * dT_dy + lap + raq)
    out = np.pad(Tn, ((0, 0), (1, 1), (1, 1)), mode="edge")
    if y_walls[0]:
        out[:, 0, :] = 1.0
    if y_walls[1]:
        out[:, -1, :] = 0.0
    return out.astype(T.dtype), dt


def ts_step(sd, spec, T, xc, yc, raq, fkt, fkp, CN_max=0.99):
    """One iteration of TS.forward's loop (pytorch_networks_convae.py:377-473), net='newfluidnet'.
    T [B,1,H,W] -> (T_new [B,1,H,W], dt, u, v, p, V)."""
    inp, V = build_input(T, xc, yc, yc, raq, fkt, fkp)
    u, v, p = newfluidnet_forward(sd, spec, inp)
    s = velocity_scaler(raq, fkt, fkp)
    u = u * s
    v = v * s
    Tn, dt = adnet_forward(u, v, T[:, 0], raq, xc, yc, CN_max)
    Tn = apply_T_bcs(Tn)[:, None]
    return Tn, dt, u, v, p, V


def ts_rollout(sd, spec, T0, xc, yc, raq, fkt, fkp, n_steps, CN_max=0.99, keep=()):
    """TS(ts=n_steps).forward restated; returns final state and the snapshots listed in keep."""
    T = T0
    snaps, dts = {}, []
    u = v = p = V = None
    for i in range(1, n_steps + 1):
        T, dt, u, v, p, V = ts_step(sd, spec, T, xc, yc, raq, fkt, fkp, CN_max)
        dts.append(float(dt))
        if i in keep:
            snaps[i] = T.copy()
    return T, np.asarray(dts), u, v, p, V, snaps


# ----------------------------------------------------------------------------- diagnostics (A11)
def diagnostics(T, y):
    """mean-T (advect_wi_gaia.py:547,647); horizontally averaged profile and its gradient
    (.ipynb_checkpoints/load_advection_results-checkpoint.ipynb:322-323) with r generalised
    to the run's y vector.  T [H,W], y [H]."""
    Tp = T.mean(axis=-1)
    dTp = (Tp[1:] - Tp[:-1]) / (y[1:] - y[:-1])
    return T.mean(), Tp, dTp


# ----------------------------------------------------------------------------- learned boundary conv (A4)
_REGIONS = ("conv", "conv_top_left", "conv_top_right", "conv_bottom_left", "conv_bottom_right",
            "conv_top", "conv_bottom", "conv_left", "conv_right")


def boundary_learned_conv(x, sd, prefix, k, c_out, use_symm=False, bc_x=1, bc_y=1):
    """BoundaryLearnedConvolution2D.forward, pytorch_networks_convae.py:1022-1065.
    NB the row swap at :1060: the strip computed from the LAST input rows lands at output
    row 0 and vice versa; left/right are not swapped."""
    W = {}
    for r in _REGIONS:
        w = sd[f"{prefix}{r}.weight"]
        W[r] = expand_symmetric(w, c_out) if use_symm else w
    pad_x = k + 1 + (bc_x - 1) if k == 5 else k + (bc_x - 1)
    pad_y = k + 1 + (bc_y - 1) if k == 5 else k + (bc_y - 1)
    tl = conv2d_valid(x[:, :, :pad_y, :pad_x], W["conv_top_left"])
    bl = conv2d_valid(x[:, :, -pad_y:, :pad_x], W["conv_bottom_left"])
    tr = conv2d_valid(x[:, :, :pad_y, -pad_x:], W["conv_top_right"])
    br = conv2d_valid(x[:, :, -pad_y:, -pad_x:], W["conv_bottom_right"])
    top = conv2d_valid(x[:, :, :pad_y, :], W["conv_top"])
    left = conv2d_valid(x[:, :, :, :pad_x], W["conv_left"])
    bottom = conv2d_valid(x[:, :, -pad_y:, :], W["conv_bottom"])
    right = conv2d_valid(x[:, :, :, -pad_x:], W["conv_right"])
    mid = conv2d_valid(x, W["conv"])
    mid = np.concatenate([left, mid, right], axis=3)
    top = np.concatenate([tl, top, tr], axis=3)
    bottom = np.concatenate([bl, bottom, br], axis=3)
    out = np.concatenate([bottom, mid, top], axis=2)
    return out + sd[prefix + "learnable_bias"]


def learned_net_forward(sd, spec: NetSpec, inp, fluidnet=False):
    """NewFluidNet.forward (pytorch_networks_convae.py:1315-1388) / FluidNet.forward (:1639-1697) with r_p='learned'
    (SURVEY.md section 8f N1): every conv is the 9-region conv above.  NewFluidNet ends in the wall-BC curl head;
    FluidNet enlarges the head conv's output by one ring (bc_x = bc_y = 2, :1659-1660) and returns the plain central
    differences of the stream function, cropped back to the input size, without wall BCs (:1694-1697)."""
    H, W = inp.shape[-2:]
    C, k, sym = spec.c_h, spec.f, spec.use_symm

    def layer(x, prefix, c_out):  # FluidLayer with the learned conv (:790-799)
        y = boundary_learned_conv(x, sd, prefix + "layers.0.", k, c_out, use_symm=sym)
        return gelu(group_norm(y, sd[prefix + "layers.1.weight"], sd[prefix + "layers.1.bias"], int(c_out / min(4, c_out))))

    x_in = layer(inp, "conv.0.", C)
    feats = []
    for l in range(spec.levels):
        y1 = x_in
        for _ in range(l):
            y1 = avg_pool2(y1)
        for r in range(spec.repeats):
            y1 = layer(y1, f"convs.{l}.{r}.", C)
        feats.append(bicubic_upsample(y1, (H, W)) if l > 0 else y1)
    y = np.concatenate(feats + [inp], axis=1)
    bc = 2 if fluidnet else 1
    y = boundary_learned_conv(y, sd, "conv.1.", k, C, use_symm=sym, bc_x=bc, bc_y=bc)
    y = gelu(group_norm(y, sd["gn.0.weight"], sd["gn.0.bias"], int(C / 4)))
    y = gelu(boundary_learned_conv(y, sd, "conv.2.", k, C, use_symm=sym))
    y = boundary_learned_conv(y, sd, "conv.3.", k, spec.c_o, use_symm=sym)
    if not fluidnet:
        if spec.loss_type == "curl":
            return curl_head(y, spec)
        y = y - y.mean(axis=(2, 3), keepdims=True)
        return y[:, 0], y[:, 1], (y[:, 2:3] if spec.p_pred else None)
    y = y - y.mean(axis=(2, 3), keepdims=True)
    a = y[:, 0] * spec.a_bound
    u = 0.5 * (a[:, 2:, 1:-1] - a[:, :-2, 1:-1])
    v = -0.5 * (a[:, 1:-1, 2:] - a[:, 1:-1, :-2])
    return u, v, (y[:, 1] if spec.p_pred else None)
