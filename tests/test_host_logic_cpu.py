"""Host logic of the module-level forwards, checked WITHOUT a GPU: the device ops are replaced by CPU stand-ins
(tests/_emulated_ops.py: float64 NCHW tensors play the blocked layout, ATen/numpy play the kernels), so what is tested
is the Python side -- channel bookkeeping, sizes, crops, region stitching, head mapping, filter caches -- against the
vectors produced by the real reference.  The kernels themselves are tested on the GPU (tests/test_gpu_*.py)."""
import numpy as np
import pytest
import torch

import pbml_mantle_convection_b200 as P
from oracle import ref_numpy as RN
from pbml_mantle_convection_b200 import ops
from tests import _emulated_ops as emu
from tests._util import LEARNED_CASES, UNET_CASES, load, load_learned_case, load_unet_case, relerr, split_weights


@pytest.mark.parametrize("tag", UNET_CASES)
def test_unet_forward_host_logic(monkeypatch, tag):
    emu.install(monkeypatch)
    spec, inp, outs, w = load_unet_case(tag)
    net = P.Unet(spec.levels, spec.c_i, spec.c_h, spec.c_o, "cpu", act_fn="gelu", r_p=spec.r_p, loss_type=spec.loss_type,
                 use_symm=False, a_bound=spec.a_bound, repeats=spec.repeats, f=spec.f, p_pred=spec.p_pred).double().eval()
    net.load_state_dict({k: torch.tensor(v) for k, v in w.items()})
    res = dict(zip("uvpT", net(torch.tensor(inp))))
    assert (res["p"] is None) == ("p" not in outs)
    for n, ref in outs.items():
        assert tuple(res[n].shape) == ref.shape and res[n].dtype == torch.float64
        assert relerr(res[n].numpy(), ref) < 1e-6, n  # the forward converts its input to float32 once


@pytest.mark.parametrize("tag,k,co,symm", [("blc3", 3, 8, False), ("blc5", 5, 8, False), ("blc3s", 3, 16, True)])
def test_learned_boundary_conv_host_logic(monkeypatch, tag, k, co, symm):
    """bc_x = bc_y = 1 goes to the two-launch kernel path (interior conv + ring kernel: ONE conv_learned9 call with the
    interior filters, the eight boundary filter sets in the reference's region order and the learnable bias); the
    bc_x = bc_y = 2 enlargement (FluidNet's head) keeps the nine-region composite with the reference's row swap (:1060)."""
    emu.install(monkeypatch)
    calls = []
    inner, inner9 = ops.conv_fwd, ops.conv_learned9
    monkeypatch.setattr(ops, "conv_fwd", lambda srcs, *a, impl="auto", wpk_row=None, **kw: (
        calls.append(("conv_fwd", impl)), inner(srcs, *a, impl=impl, wpk_row=wpk_row, **kw))[1])
    monkeypatch.setattr(ops, "conv_learned9", lambda *a, **kw: (calls.append(("learned9", len(a[3]))), inner9(*a, **kw))[1])
    g = load("ops")
    ci = g[tag + "_x"].shape[1]
    m = P.BoundaryLearnedConvolution2D(ci, co, k, use_symm=symm).double()
    m.load_state_dict({kk: torch.tensor(v) for kk, v in split_weights(g, tag + "_w::").items()})
    y = m(torch.tensor(g[tag + "_x"]))
    assert tuple(y.shape) == g[tag + "_y"].shape and relerr(y.numpy(), g[tag + "_y"]) < 1e-13
    assert calls == [("learned9", 8)]
    if tag == "blc3":
        calls.clear()
        y2 = m(torch.tensor(g[tag + "_x"]), bc_x=2, bc_y=2)
        assert tuple(y2.shape) == g["blc3_y_bc2"].shape and relerr(y2.numpy(), g["blc3_y_bc2"]) < 1e-13
        assert len(calls) == 9 and all(c[0] == "conv_fwd" for c in calls)
    # cached filter images follow in-place updates of ANY of the nine filter sets and of the bias
    x = torch.randn(1, ci, 40, 48, dtype=torch.float64)
    for name in ("conv", "conv_top_right", "conv_left"):
        with torch.no_grad():
            getattr(m, name).weight.mul_(0.5)
        sd = {kk: t.detach().numpy() for kk, t in m.state_dict().items()}
        assert relerr(m(x).numpy(), RN.boundary_learned_conv(x.numpy(), sd, "", k, co, use_symm=symm)) < 1e-13
    with torch.no_grad():
        m.learnable_bias.add_(1.0)
    sd = {kk: t.detach().numpy() for kk, t in m.state_dict().items()}
    assert relerr(m(x).numpy(), RN.boundary_learned_conv(x.numpy(), sd, "", k, co, use_symm=symm)) < 1e-13
    # a strip-sized input (fewer rows than the reference's boundary strips need) is refused by the kernel path
    assert not m.kernel_path_ok(k if k == 5 else 2, 40, [ci])


@pytest.mark.parametrize("tag", LEARNED_CASES)
def test_learned_network_host_logic(monkeypatch, tag):
    """Whole learned-boundary networks through the module-level forward (pyramid, concat order, head enlargement of
    FluidNet, curl heads) against the reference's outputs."""
    emu.install(monkeypatch)
    spec, inp, outs, w = load_learned_case(tag)
    cls = P.FluidNet if tag == "learned_fluidnet" else P.NewFluidNet
    net = cls(spec.levels, spec.c_i, spec.c_h, spec.c_o, "cpu", act_fn="gelu", r_p="learned", loss_type="curl", use_symm=False,
              a_bound=spec.a_bound, repeats=spec.repeats, f=spec.f, p_pred=spec.p_pred).double().eval()
    net.load_state_dict({k: torch.tensor(v) for k, v in w.items()})
    net.use_cuda_graph = False
    res = dict(zip("uvp", net(torch.tensor(inp))))
    assert (res["p"] is None) == ("p" not in outs)
    for n, ref in outs.items():
        assert tuple(res[n].shape) == ref.shape and relerr(res[n].numpy(), ref) < 1e-6, n


def test_TS_unet_branch_host_logic(monkeypatch):
    """`TS(net="unet")` (reference :419-451): channel order of the 10-channel input, u_prev / v_prev / dt passed through,
    wall BCs on the predicted T, empty dts, p None -- against the oracle's U-Net applied to the oracle's input build."""
    emu.install(monkeypatch)
    spec, _inp, _outs, w = load_unet_case("unet_curl_p")
    net = P.Unet(spec.levels, spec.c_i, spec.c_h, spec.c_o, "cpu", act_fn="gelu", r_p=spec.r_p, loss_type=spec.loss_type,
                 use_symm=False, a_bound=spec.a_bound, repeats=spec.repeats, f=spec.f, p_pred=spec.p_pred).double().eval()
    net.load_state_dict({k: torch.tensor(v) for k, v in w.items()})
    H, W = 36, 50
    xc, yc = RN.synthetic_grid(H, W)
    T0 = RN.synthetic_T0(H, W, seed=1)
    raq, fkt, fkp = 6.79733173, 475523342.0, 2.58574662
    nd = RN.nondim_params(raq, fkt, fkp)
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    g = torch.Generator().manual_seed(3)
    up, vp = torch.randn(1, 1, H, W, generator=g, dtype=torch.float64), torch.randn(1, 1, H, W, generator=g, dtype=torch.float64)
    dt = torch.full((1, 1, H, W), 1e-3, dtype=torch.float64)
    ts = P.TS(net, None, "cpu", ts=2, scale=True, p_pred=True, net="unet")
    grid = lambda a: t64(a).view(1, 1, H, W)
    x, dts, u, v, p, V = ts(grid(T0), None, None, grid(yc), t64(nd[0]), t64(nd[1]), t64(nd[2]), t64(raq), t64(fkt), t64(fkp),
                            grid(xc), grid(yc), u_prev=up, v_prev=vp, dt=dt)
    assert sorted(x) == [0, 1, 2] and dts == {} and p is None
    T = T0[None, None]
    for i in (1, 2):
        inp7, _V = RN.build_input(T, xc, yc, yc, raq, fkt, fkp)
        inp = np.concatenate([inp7[:, 0:2], dt.numpy(), inp7[:, 3:6], inp7[:, 2:3], inp7[:, 6:7], up.numpy(), vp.numpy()], 1)
        u_ref, v_ref, _p, Tn = RN.unet_forward(w, spec, inp)
        T = RN.apply_T_bcs(Tn[:, None].copy())
        assert tuple(x[i].shape) == (1, 1, H, W) and np.abs(x[i].numpy() - T).max() < 1e-6
    assert relerr(u.numpy()[0], u_ref) < 1e-5 and relerr(v.numpy()[0], v_ref) < 1e-5
    assert np.abs(V.numpy() - inp7[:, 2:3]).max() < 1e-6
    # a scalar dt is broadcast
    x2 = ts(grid(T0), None, None, grid(yc), t64(nd[0]), t64(nd[1]), t64(nd[2]), t64(raq), t64(fkt), t64(fkp), grid(xc), grid(yc),
            u_prev=up, v_prev=vp, dt=torch.tensor(1e-3, dtype=torch.float64))[0]
    assert np.abs(x2[2].numpy() - x[2].numpy()).max() < 1e-12
    with pytest.raises(ValueError):
        ts(grid(T0), None, None, grid(yc), t64(nd[0]), t64(nd[1]), t64(nd[2]), t64(raq), t64(fkt), t64(fkp), grid(xc), grid(yc))
