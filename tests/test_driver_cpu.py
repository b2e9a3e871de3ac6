"""The GAIA-free rollout driver: loop shape, cadence rules and the on-disk pickle layout
of advect_wi_gaia.py:538-679, exercised on CPU with a stand-in for `ts_net` (any callable with TS.forward's
signature works -- the CUDA drop-in is covered by the -m gpu tests)."""
import pickle

import numpy as np
import torch

from pbml_mantle_convection_b200 import driver as D

H, W = 6, 8


def fake_ts(Tp, sdf, sdf2, ycc, raq_nd, fkt_nd, fkp_nd, raq, fkt, fkp, xcc, ycc2):
    """T decays by 10 % per step, dt = 0.25, u = 1, v = 2, p = 3, V = 4 (shapes like TS.forward, ts=1)."""
    Tn = Tp * 0.9
    f = lambda c: torch.full_like(Tp, c)
    return {0: Tp, 1: Tn}, {1: torch.tensor(0.25, dtype=torch.float64)}, f(1.0), f(2.0), f(3.0), f(4.0)


def _grid():
    x = torch.linspace(0, 4, W, dtype=torch.float64).view(1, 1, 1, W).expand(1, 1, H, W).contiguous()
    y = torch.linspace(0, 1, H, dtype=torch.float64).view(1, 1, H, 1).expand(1, 1, H, W).contiguous()
    return x, y


def test_attempt_loop_and_pickle_layout(tmp_path):
    xcc, ycc = _grid()
    T0 = torch.ones(1, 1, H, W, dtype=torch.float64)
    s = torch.tensor(1.0, dtype=torch.float64)
    t, n_step, (snaps, TS_vec, t_vec, T_vec) = D.attempt(fake_ts, T0, xcc, ycc, s, s, s, s, s, s, t_end=2.0, out_dir=str(tmp_path),
                                                       save_every=0.5, write_every=1.0)
    assert n_step == 8 and abs(t - 2.0) < 1e-12  # while t < t_end, dt = 0.25
    assert np.allclose(t_vec["ML"], np.arange(9) * 0.25) and len(T_vec["ML"]) == 9 and len(TS_vec["ML"]) == 8
    assert np.allclose(T_vec["ML"], 0.9 ** np.arange(9))  # the new T is fed back (module docstring)
    # cadence: initial snapshot, then whenever t > save_t (0 -> 0.75 -> 1.5 ...): steps at t = 0.25, 1.0, 1.75
    assert len(snaps["ML"]["T"]) == 4
    assert np.allclose([a.mean() for a in snaps["ML"]["T"]], [1.0, 0.9, 0.9 ** 4, 0.9 ** 7])
    last_v = snaps["ML"]["v"][-1]
    assert last_v.shape == (H * W, 3) and np.all(last_v[:, 0] == 1.0) and np.all(last_v[:, 1] == 2.0) and np.all(last_v[:, 2] == 0.0)
    assert snaps["ML"]["P"][-1].shape == (H * W,) and np.all(snaps["ML"]["P"][-1] == 3.0)
    for name in ("snapshots_ML.pkl", "TS_vec_ML.pkl", "t_vec_ML.pkl", "T_vec_ML.pkl"):
        assert (tmp_path / name).exists()
    with open(tmp_path / "snapshots_ML.pkl", "rb") as fh:
        disk = pickle.load(fh)
    assert sorted(disk.keys()) == ["P", "T", "v", "xcc", "ycc"] and len(disk["T"]) == 4 and torch.equal(disk["xcc"], xcc)
    with open(tmp_path / "t_vec_ML.pkl", "rb") as fh:
        assert np.allclose(pickle.load(fh), t_vec["ML"])


def test_attempt_respects_max_steps_and_restart_arguments(tmp_path):
    xcc, ycc = _grid()
    T0 = torch.full((1, 1, H, W), 2.0, dtype=torch.float64)
    s = torch.tensor(1.0, dtype=torch.float64)
    t, n_step, logs = D.attempt(fake_ts, T0, xcc, ycc, s, s, s, s, s, s, t_end=100.0, out_dir=str(tmp_path), mode="ML_STOKES",
                                t=5.0, n_step=40, max_steps=43, p_pred=False)
    assert n_step == 43 and abs(t - 5.75) < 1e-12
    assert (tmp_path / "snapshots_ML_STOKES.pkl").exists() and np.all(logs[0]["ML_STOKES"]["P"][-1] == 0.0)


def test_attempt_external_energy_solver_hook(tmp_path):
    """ML_STOKES (advect_wi_gaia.py:486-505, :618-630): the surrogate supplies the velocities, an external solver advances T
    and returns dt every step; in mode "ML" it intervenes every `intervene_TS`-th step.  The driver applies the reference's
    wall rows / side columns / clip(0, 2) to what the solver returns."""
    xcc, ycc = _grid()
    T0 = torch.full((1, 1, H, W), 0.5, dtype=torch.float64)
    s = torch.tensor(1.0, dtype=torch.float64)
    seen = []

    def no_ad_ts(Tp, *a):  # TS(stokes, None, ...): no ADNet -> x has no entry 1, dts is empty
        f = lambda c: torch.full_like(Tp, c)
        return {0: Tp}, {}, f(1.0), f(2.0), f(3.0), f(4.0)

    def solver(state):
        seen.append((state["v"][0, 0], state["v"][0, 1], state["V"][0]))
        state["T"][:] = state["T"] + 1.0  # far above 2: must be clipped; walls must be re-imposed
        return 0.5

    t, n_step, (snaps, TS_vec, t_vec, T_vec) = D.attempt(no_ad_ts, T0, xcc, ycc, s, s, s, s, s, s, t_end=1.4, out_dir=str(tmp_path),
                                                       mode="ML_STOKES", energy_step=solver)
    assert n_step == 3 and abs(t - 1.5) < 1e-12 and seen == [(1.0, 2.0, 4.0)] * 3
    last = snaps["ML_STOKES"]["T"][-1].reshape(H, W)
    assert np.all(last[0] == 1.0) and np.all(last[-1] == 0.0) and np.all(last[1:-1] == 2.0) and (tmp_path / "T_vec_ML_STOKES.pkl").exists()
    # mode "ML": the surrogate's own T / dt except on every 2nd step
    seen.clear()
    t, n_step, (snaps, _, t_vec, _) = D.attempt(fake_ts, T0, xcc, ycc, s, s, s, s, s, s, t_end=100.0, out_dir=str(tmp_path), max_steps=4,
                                                energy_step=solver, intervene_TS=2)
    assert len(seen) == 2 and np.allclose(t_vec["ML"], [0.0, 0.25, 0.75, 1.0, 1.5])
    # without a solver a TS that has no ADNet cannot advance T in a non-ML mode
    class _NoAd:
        ad = None
        __call__ = staticmethod(no_ad_ts)
    import pytest
    with pytest.raises(ValueError):
        D.attempt(_NoAd(), T0, xcc, ycc, s, s, s, s, s, s, t_end=1.0, out_dir=str(tmp_path), mode="ML_STOKES")
