"""Shared helpers for the tests: golden loading and error norms."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def load_weights(name):
    return load(name + "_weights")


def split_weights(d, prefix="w::"):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def spec_from_variant(d):
    from oracle.ref_numpy import NetSpec

    s = d["spec"]
    return NetSpec(levels=int(s[0]), c_i=int(s[1]), c_h=int(s[2]), c_o=int(s[3]), repeats=int(s[4]), f=int(s[5]),
                   use_symm=bool(s[6]), p_pred=bool(s[7]), r_p=str(d["r_p"]), loss_type=str(d["loss_type"]),
                   a_bound=float(d["a_bound"]))


VARIANTS = ["var_zeros_nosym", "var_reflect_sym", "var_replicate_odd", "var_mae_p", "var_k5"]


UNET_CASES = ["unet_curl_p", "unet_curl_k5", "unet_mae"]


LEARNED_CASES = ["learned_k5", "learned_k3_p", "learned_fluidnet"]


def load_learned_case(tag):
    """One case of tests/golden/learned.npz (whole learned-boundary networks): like load_unet_case."""
    return load_unet_case(tag, "learned")


def load_unet_case(tag, file="unet"):
    """One case of tests/golden/<file>.npz: (spec, input, {name: array for u, v, p, T}, state_dict as numpy)."""
    g = load(file)
    d = {k[len(tag) + 2:]: v for k, v in g.items() if k.startswith(tag + "::")}
    spec = spec_from_variant(d)
    outs = {n: d[n] for n in "uvpT" if n in d}
    return spec, d["inp"], outs, split_weights(d)
