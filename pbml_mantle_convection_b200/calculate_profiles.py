"""Steady 1-D temperature profile from the reference's pre-trained numpy MLP -- drop-in for
calculate_profiles.py:57-134 (used once per rollout to initialise T, advect_wi_gaia.py:227).

Host-side numpy like the reference (six small dense layers; not a kernel target).  The MLP
weights are the reference's own asset `mlp_[128, 128, 128, 128, 128].pkl`; like the reference
the file is looked up in the current directory, or pass `mlp_path=` / set PBMC_MLP_PKL.
"""
import os
import pickle

import numpy as np

_RAQ = (0.12624371, 9.70723344)
_FKT = (6.00352841978384, 9.888820429862925)
_FKV = (0.005251646002323797, 1.9927988938926755)
MLP_FILE = "mlp_[128, 128, 128, 128, 128].pkl"


def selu(x):
    alpha, scale = 1.6732632423543772848170429916717, 1.0507009873554804934193349852946
    return scale * (np.maximum(0, x) + np.minimum(alpha * (np.exp(x) - 1), 0))


def _to_unit(rng, log):
    lo, span = rng[0], rng[1] - rng[0]
    return (lambda x: (np.log10(x) - lo) / span) if log else (lambda x: (x - lo) / span)


def _from_unit(rng, log):
    lo, span = rng[0], rng[1] - rng[0]
    return (lambda x: 10 ** (x * span + lo)) if log else (lambda x: x * span + lo)


# the reference's six public helpers (calculate_profiles.py:13-36): RaQ is scaled linearly, the two viscosity
# contrasts in log10; the ranges are those of its 130 training simulations
non_dimensionalize_raq, dimensionalize_raq = _to_unit(_RAQ, False), _from_unit(_RAQ, False)
non_dimensionalize_fkt, dimensionalize_fkt = _to_unit(_FKT, True), _from_unit(_FKT, True)
non_dimensionalize_fkv, dimensionalize_fkv = _to_unit(_FKV, True), _from_unit(_FKV, True)


def get_input(raq_ra, fkt, fkp, y_prof):
    """One row (raq_nd, fkt_nd, fkv_nd, y) per (simulation, profile point)."""
    n, m = len(raq_ra), len(y_prof)
    x = np.zeros((n * m, 4))
    x[:, 0] = np.repeat(non_dimensionalize_raq(np.asarray(raq_ra, dtype=np.float64)), m)
    x[:, 1] = np.repeat(non_dimensionalize_fkt(np.asarray(fkt, dtype=np.float64)), m)
    x[:, 2] = np.repeat(non_dimensionalize_fkv(np.asarray(fkp, dtype=np.float64)), m)
    x[:, 3] = np.tile(np.asarray(y_prof, dtype=np.float64), n)
    return x


def get_profile(inp, mlp, num_sims=1, correction=True, prof_points=128):
    """Residual-SELU MLP forward + wall overwrite + boundary-layer correction (reference :57-99)."""
    last = len(mlp) - 1
    h = inp
    skips = []
    for l, (Wm, b) in enumerate(mlp):
        h = h @ Wm.T + b
        if l == last - 1:
            h = np.concatenate((inp, h), axis=-1)
        if l != last:
            for s in skips:
                h = h + s
            h = selu(h)
            skips.append(h)
    y = h.reshape(num_sims, -1)
    y[:, 0] = 1.0
    y[:, -1] = 0.0
    if correction:
        pts = inp.reshape(num_sims, -1, inp.shape[-1])
        for i in range(num_sims):
            yy = pts[i, :, 3]
            lo = np.where(yy < 0.04)[0]
            slope = (0 - y[i, lo[0]]) / (0 - yy[lo[0]])
            y[i, lo] = slope * yy[lo]
            hi = np.where(yy > 0.985)[0]
            y[i, hi] = np.interp(yy[hi], [yy[hi[-1]], 1], [y[i, hi[-1]], 1])
    return y


def calc_mlp_profile(r_list, t_list, v_list, simulation_dir=None, num_points=128, mlp_path=None):
    path = mlp_path or os.environ.get("PBMC_MLP_PKL") or MLP_FILE
    with open(path, "rb") as fh:
        mlp = pickle.load(fh)
    half = 1 / (num_points * 2)
    y_prof = np.concatenate(([1], np.linspace(half, 1 - half, num_points - 2)[::-1], [0]))
    pred = get_profile(get_input(r_list, t_list, v_list, y_prof), mlp, num_sims=len(r_list))
    if simulation_dir is not None:
        for i in range(len(r_list)):
            with open(simulation_dir + "/ml_prof.txt", "wb") as fh:
                for j in range(len(y_prof)):
                    fh.write(f"{y_prof[j]}   {pred[i, j]}\n".encode("ascii"))
    return pred, y_prof
