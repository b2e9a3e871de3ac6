"""Host-side behaviour of the drop-in modules that needs no GPU: constructor signatures,
state_dict layout (must load the reference's checkpoints), error behaviour, numpy helpers."""
import inspect
import os

import numpy as np
import pytest
import torch

import pbml_mantle_convection_b200 as P
from pbml_mantle_convection_b200 import _lib as L
from pbml_mantle_convection_b200 import calculate_profiles as CP
from pbml_mantle_convection_b200 import ops, scaler
from tests._util import GOLDEN, UNET_CASES, VARIANTS, load, load_unet_case, load_weights, spec_from_variant, split_weights


def make(spec, cls=P.NewFluidNet):
    return cls(spec.levels, spec.c_i, spec.c_h, spec.c_o, "cpu", act_fn="gelu", r_p=spec.r_p, loss_type=spec.loss_type,
               use_symm=spec.use_symm, a_bound=spec.a_bound, repeats=spec.repeats, f=spec.f, p_pred=spec.p_pred)


def test_state_dict_layout_matches_reference_checkpoint():
    from oracle.ref_numpy import NetSpec

    ref = load_weights("roll128")
    net = make(NetSpec()).double()
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())  # same keys in the same order (108 tensors)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == v.shape, k
        assert sd[k].dtype == torch.float64
    net.load_state_dict({k: torch.tensor(v) for k, v in ref.items()})
    assert P.count_parameters(net) == 67956


@pytest.mark.parametrize("tag", VARIANTS)
def test_state_dict_layout_variants(tag):
    g = load(tag)
    spec = spec_from_variant(g)
    net = make(spec).double()
    w = split_weights(g)
    assert list(net.state_dict().keys()) == list(w.keys())
    net.load_state_dict({k: torch.tensor(v) for k, v in w.items()})


def test_learned_boundary_state_dict_and_constraints():
    g = load("ops")
    for tag, k, ci, co, symm in (("blc3", 3, 5, 8, False), ("blc5", 5, 4, 8, False), ("blc3s", 3, 6, 16, True)):
        m = P.BoundaryLearnedConvolution2D(ci, co, k, use_symm=symm).double()
        w = split_weights(g, tag + "_w::")
        assert sorted(m.state_dict().keys()) == sorted(w.keys())
        m.load_state_dict({kk: torch.tensor(v) for kk, v in w.items()})
    net = P.NewFluidNet(5, 7, 16, 1, "cpu", act_fn="gelu", r_p="learned", loss_type="curl", repeats=6, f=5, p_pred=False)
    assert len(net.state_dict()) == 404
    with pytest.raises(ValueError):  # learned + symm + c_o in {2,3}: odd number of mirrored filters
        P.NewFluidNet(2, 7, 16, 2, "cpu", act_fn="gelu", r_p="learned", loss_type="curl", use_symm=True)


def test_same_seed_same_init_as_reference_layout():
    # SymmetricConv2d keeps only the unique filters: 16 outputs -> 14 stored (h = 4 mirrored pairs -> 2 copies)
    c = P.SymmetricConv2d(16, 16, 3, padding="same", padding_mode="replicate", symmetry={"h": 4, "v": 0, "hv": 0})
    assert tuple(c.weight.shape) == (14, 16, 3, 3) and tuple(c.bias.shape) == (16,)
    from pbml_mantle_convection_b200.symmetric_layers_torch import full_weight

    w = full_weight(c)
    assert tuple(w.shape) == (16, 16, 3, 3)
    assert torch.equal(w[14], torch.flip(c.weight[0], (2,))) and torch.equal(w[15], torch.flip(c.weight[1], (2,)))
    with pytest.raises(ValueError):
        P.SymmetricConv2d(4, 4, 3, symmetry={"h": 3})
    with pytest.raises(ValueError):
        P.SymmetricConv2d(4, 4, 3, groups=2, symmetry={"h": 2})


def test_constructor_signatures_match_reference():
    want_net = ["levels", "c_i", "c_h", "c_o", "device", "act_fn", "r_p", "loss_type", "use_symm", "dilation", "a_bound",
                "use_cosine", "repeats", "use_skip", "f", "p_pred", "spectral_conv", "blurr", "drop_rate", "factor"]
    for cls in (P.NewFluidNet, P.FluidNet):
        assert list(inspect.signature(cls.__init__).parameters)[1:] == want_net
    assert list(inspect.signature(P.TS.__init__).parameters)[1:] == ["stokes", "ad", "device", "ts", "advection_scheme",
                                                                    "scale", "p_pred", "net"]
    assert list(inspect.signature(P.TS.forward).parameters)[1:] == ["T_prev", "sdf", "sdf2", "ycc", "raq_nd", "fkt_nd",
                                                                   "fkp_nd", "raq", "fkt", "fkp", "xc", "yc", "u_prev",
                                                                   "v_prev", "dt"]
    assert list(inspect.signature(P.ADNet.__init__).parameters)[1:] == ["device", "r_p", "CN_max"]
    assert list(inspect.signature(P.ADNet.forward).parameters)[1:] == ["inputs", "dt", "T_prev"]
    assert list(inspect.signature(P.FluidLayer.__init__).parameters)[1:] == ["c_i", "c_o", "act_fn", "r_p", "use_symm",
                                                                             "dilation", "f", "drop_rate"]


def test_no_cpu_fallback():
    from oracle.ref_numpy import NetSpec

    net = make(NetSpec(levels=2, repeats=1))
    with pytest.raises(L.PbmcError):
        net(torch.zeros(1, 7, 16, 16))
    with pytest.raises(L.PbmcError):
        P.ADNet("cpu", CN_max=0.99)(torch.zeros(1, 6, 8, 8))
    with pytest.raises(L.PbmcError):
        net.conv[0](torch.zeros(1, 7, 16, 16))


def test_weight_packing_layout():
    w = torch.arange(2 * 7 * 9, dtype=torch.float32).reshape(2, 7, 3, 3)
    wpk = ops.pack_conv_weight(w, [7])
    assert tuple(wpk.shape) == (1, 2, 9, 4, 16)
    for co, ci, t in ((0, 0, 0), (1, 6, 8), (1, 3, 4)):
        assert wpk[0, ci // 4, t, ci % 4, co] == w[co, ci, t // 3, t % 3]
    assert wpk[0, 1, :, 3, :].abs().sum() == 0 and wpk[0, :, :, :, 2:].abs().sum() == 0  # padding lanes are zero
    w2 = torch.randn(16, 23, 3, 3)
    wpk2 = ops.pack_conv_weight(w2, [16, 7])  # concat of two sources: 16 + 7(->8) channels
    assert tuple(wpk2.shape) == (1, 6, 9, 4, 16)
    assert wpk2[0, 4, 2, 1, 5] == w2[5, 17, 0, 2] and wpk2[0, 5, :, 3, :].abs().sum() == 0


def test_member_constants():
    from oracle import ref_numpy as RN

    raq, fkt, fkp = 6.79733173, 475523342.0, 2.58574662
    m = ops.member_values(raq, fkt, fkp)
    nd = RN.nondim_params(raq, fkt, fkp)
    assert np.allclose(m[:3], nd, rtol=1e-15)
    assert np.isclose(m[6], RN.velocity_scaler(raq, fkt, fkp), rtol=1e-15)


def test_scaler_mutates_and_returns():
    x = np.ones(4)
    y = scaler.unscale_var(x, 6.79733173, 475523342.0, 2.58574662, "uprev")
    assert y is x and np.allclose(x, 62866.27, rtol=1e-4)
    scaler.scale_var(x, 6.79733173, 475523342.0, 2.58574662, "vprev")
    assert np.allclose(x, 1.0)
    z = np.ones(3)
    assert scaler.scale_var(z, 1.0, 1e7, 2.0, "Tprev_0") is z and np.all(z == 1)
    assert scaler.unscale_var(z, 1.0, 1e7, 2.0, "pprev") is z and np.all(z == 1)


def test_calc_mlp_profile_known_answer(tmp_path):
    import pickle

    wz = np.load(os.path.join(GOLDEN, "mlp_profile_weights.npz"))
    mlp = [[wz[f"W{i}"], wz[f"b{i}"]] for i in range(6)]
    path = tmp_path / CP.MLP_FILE
    with open(path, "wb") as fh:
        pickle.dump(mlp, fh)
    pred, yprof = CP.calc_mlp_profile([6.79733173], [475523342.0], [2.58574662], simulation_dir=str(tmp_path),
                                      mlp_path=str(path))
    g = load("ops")
    assert np.abs(pred - g["mlp_pred"]).max() < 1e-13 and np.array_equal(yprof, g["mlp_yprof"])
    # SURVEY.md section 8a A13 known answer
    assert np.allclose(pred[0, :3], [1, 0.9988626, 0.9965514], atol=1e-7)
    assert np.allclose(pred[0, -3:], [0.08812508, 0.029065, 0], atol=1e-7) and abs(pred.mean() - 0.919107195) < 1e-8
    lines = open(tmp_path / "ml_prof.txt").read().splitlines()
    assert len(lines) == 128 and lines[0].split() == ["1.0", "1.0"]


def test_synthetic_inputs_match_oracle():
    from oracle import ref_numpy as RN

    for H, W in ((128, 128), (128, 506), (50, 77)):
        xa, ya = P.synthetic_grid(H, W)
        xb, yb = RN.synthetic_grid(H, W)
        assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
        assert np.array_equal(P.synthetic_T0(H, W, seed=3), RN.synthetic_T0(H, W, seed=3))


@pytest.mark.parametrize("tag", UNET_CASES)
def test_unet_state_dict_layout_and_signature(tag):
    """the U-Net drop-in has the reference's module tree (keys, order, shapes), loads its
    weights, keeps its constructor signature (reference :1765-1786) and refuses CPU tensors."""
    spec, inp, _outs, w = load_unet_case(tag)
    net = P.Unet(spec.levels, spec.c_i, spec.c_h, spec.c_o, "cpu", act_fn="gelu", r_p=spec.r_p, loss_type=spec.loss_type,
                 use_symm=False, a_bound=spec.a_bound, repeats=spec.repeats, f=spec.f, p_pred=spec.p_pred).double()
    sd = net.state_dict()
    assert list(sd.keys()) == list(w.keys())
    for k, v in w.items():
        assert tuple(sd[k].shape) == v.shape, k
    net.load_state_dict({k: torch.tensor(v) for k, v in w.items()})
    names = list(inspect.signature(P.Unet.__init__).parameters)[1:]
    assert names == ["levels", "c_i", "c_h", "c_o", "device", "act_fn", "r_p", "loss_type", "use_symm", "dilation", "a_bound",
                     "use_cosine", "repeats", "use_skip", "f", "p_pred", "spectral_conv", "blurr", "drop_rate"]
    d = inspect.signature(P.Unet.__init__).parameters
    assert (d["act_fn"].default, d["r_p"].default, d["loss_type"].default, d["a_bound"].default, d["repeats"].default,
            d["f"].default, d["p_pred"].default) == ("gelu", "replicate", "curl", 10.0, 2, 5, False)
    with pytest.raises(L.PbmcError):
        net(torch.tensor(inp))
