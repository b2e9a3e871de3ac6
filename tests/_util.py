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


def noise(tag):
    """The reference's OWN float32-vs-float64 distance on golden case `tag` (tests/golden/ref_fp32_noise.json, written by
    make_golden.py from the real reference): rel-L2 per output field, max-abs for T, relative for dt."""
    import json

    return json.load(open(os.path.join(GOLDEN, "ref_fp32_noise.json")))[tag]


def field_bound(ref_noise, floor=1e-5):
    """north_star's 1e-5 relative bound, widened to 1.5 x the reference's own fp32 noise where that is larger
    (no fp32 implementation, PyTorch's included, is closer to the fp64 reference than that)."""
    return max(floor, 1.5 * float(ref_noise))


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


UNET_CASES = ["unet_curl_p", "unet_curl_k5", "unet_mae", "unet_learned_k3", "unet_learned_k5"]


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
