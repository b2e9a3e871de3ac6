"""Velocity (de)normalisation -- drop-in for the reference's scaler.py:4-71.

Same signatures and the same in-place behaviour: `x` is modified (`/=`, `*=`) AND returned;
only "uprev"/"vprev" are scaled, every other variable passes through.  Host-side numpy, as
in the reference; the same constant is fused into the head kernel (csrc/head.cu).
"""
import numpy as np

_A_RAQ, _A_FKT, _A_FKP, _MUL = 1.80167667, 0.4330392, -0.46052953, 5


def velocity_scaler(raq, fkt, fkp):
    return np.exp((raq / 10) * _A_RAQ + np.log(fkt) * _A_FKT + np.log(fkp) * _A_FKP) * _MUL


def scale_var(x, raq, fkt, fkp, var):
    if var in ("uprev", "vprev"):
        x /= velocity_scaler(raq, fkt, fkp)
    return x


def unscale_var(x, raq, fkt, fkp, var):
    if var in ("uprev", "vprev"):
        x *= velocity_scaler(raq, fkt, fkp)
    return x
