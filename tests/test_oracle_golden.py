"""Pins the CPU oracle (oracle/qmpc_oracle.c) against the reference's own artefacts:
shipped acados run logs and outputs of the reference's numpy code (tests/golden, made by
oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as orc
from conftest import rel_err

N, DT = 10, 0.1     # execute_trajectory.py:122-123 (t_lookahead=1, n_nodes=10)


def _rti_replay(g, quad, gp=None, nmax=None):
    xit, uit = np.zeros((N + 1, 13)), np.zeros((N, 4))     # zero initial iterate
    xref = g["x_ref"]
    n = len(g["x_odom"]) if nmax is None else nmax
    eu, ec, st = [], [], []
    for i in range(min(n, len(xref) - N)):
        yref, yref_e = orc.make_yref(xref[i:i + N])
        alpha = None
        if gp is not None:
            p = np.zeros((3, gp.M)) if i == 0 else g["rgp_mu"][i - 1]   # params pushed one step earlier
            alpha = gp.alpha(p)
        r = orc.rti_step(quad, DT, N, g["x_odom"][i], yref, yref_e, xit, uit, gp=gp, alpha=alpha)
        eu.append(np.abs(uit[0] - g["w_odom"][i]).max())
        ec.append(abs(r["cost"] - g["cost_solution"][i]) / abs(g["cost_solution"][i]))
        st.append(r["status"])
        assert np.allclose(xit[0], g["x_odom"][i], atol=0, rtol=0)
    return np.array(eu), np.array(ec), np.array(st)


def test_rti_gp0_log(golden):
    g = golden("traj2_v10_a10_gp0")
    eu, ec, st = _rti_replay(g, orc.quad_logged_pysim())
    assert len(eu) == 289
    assert (st == 0).all()                      # every QP solved exactly (polished)
    assert np.median(eu) < 5e-8 and eu.max() < 1e-5          # distance to acados+HPIPM log
    assert ec.max() < 2e-5
    assert abs(g["cost_solution"][0] - 45.926709) < 1e-5


@pytest.mark.parametrize("name,nmax", [("traj0_v10_a10_gp2", None), ("traj1_v15_a5_gp2", 46)])
def test_rti_rgp_log(golden, name, nmax):
    g = golden(name)
    gp = orc.GPSpec(g["rgp_X"], g["rgp_theta"])
    eu, ec, st = _rti_replay(g, orc.quad_logged_pysim(), gp, nmax)
    assert (st == 0).all()
    assert np.median(eu) < 5e-8 and eu.max() < 5e-6
    assert ec.max() < 5e-6


@pytest.mark.parametrize("name", ["traj2_v10_a10_gp0", "traj0_v10_a10_gp2", "traj1_v15_a5_gp2", "traj0_v15_a5_gp2"])
def test_nominal_rk4_log(golden, name):
    g = golden(name)
    quad = orc.quad_logged_pysim()
    for i in range(len(g["x_odom"])):
        xp = orc.rk4(quad, g["x_odom"][i], g["w_odom"][i], DT)
        assert np.abs(xp - g["x_pred_odom"][i]).max() < 1e-13


@pytest.mark.parametrize("name,tol", [("traj0_v10_a10_gp2", 1e-11), ("traj1_v15_a5_gp2", 1e-11),
                                      ("traj0_v15_a5_gp2", 1e-11), ("traj2_v10_a10_gp2", 1e-8)])
def test_rgp_log(golden, name, tol):
    g = golden(name)
    gp = orc.GPSpec(g["rgp_X"], g["rgp_theta"])
    mu = np.zeros((3, gp.M))
    Cm = np.stack([orc.rgp_prior(gp.X[d], gp.theta[d])[0] for d in range(3)])
    for i in range(len(g["v_body"])):
        for d in range(3):
            orc.rgp_regress(gp.X[d], gp.theta[d], gp.Kx_inv[d], mu[d], Cm[d], g["v_body"][i, d], g["a_drag"][i, d])
        assert rel_err(mu, g["rgp_mu"][i]) < tol
        assert rel_err(Cm, g["rgp_C"][i]) < tol


@pytest.mark.parametrize("name", ["traj0_v10_a10_gp2", "traj1_v15_a5_gp2"])
def test_drag_residual_log(golden, name):
    g = golden(name)
    for i in range(1, len(g["v_body"])):
        vb, ad = orc.compute_a_drag(g["x_odom"][i], g["x_pred_odom"][i - 1], DT)
        assert np.abs(vb - g["v_body"][i]).max() < 1e-12
        assert np.abs(ad - g["a_drag"][i]).max() < 1e-12


@pytest.mark.parametrize("tag", ["m20", "m50", "m7"])
def test_rgp_reference_code(golden, tag):
    g = golden("reference_code")
    X, th = g[f"rgp_{tag}_X"], g[f"rgp_{tag}_theta"]
    M = X.shape[1]
    for d in range(3):
        Kx, Kinv = orc.rgp_prior(X[d], th[d])
        assert rel_err(Kx, g[f"rgp_{tag}_Kx"][d]) < 1e-14
        assert rel_err(Kinv, g[f"rgp_{tag}_Kx_inv"][d]) < 1e-9
    gp = orc.GPSpec(X, th, Kx_inv=g[f"rgp_{tag}_Kx_inv"])
    mu = np.zeros((3, M))
    Cm = g[f"rgp_{tag}_Kx"].copy()
    for t in range(len(g[f"rgp_{tag}_xt"])):
        for d in range(3):
            orc.rgp_regress(X[d], th[d], gp.Kx_inv[d], mu[d], Cm[d], g[f"rgp_{tag}_xt"][t, d], g[f"rgp_{tag}_yt"][t, d])
        assert rel_err(mu, g[f"rgp_{tag}_mu"][t]) < 1e-9
        assert rel_err(Cm, g[f"rgp_{tag}_C"][t]) < 1e-9
    for d in range(3):
        mean, var = orc.rgp_predict(X[d], th[d], gp.Kx_inv[d], mu[d], Cm[d], g[f"rgp_{tag}_pred_x"])
        assert rel_err(mean, g[f"rgp_{tag}_pred_mean"][d]) < 1e-9
        assert np.abs(var - g[f"rgp_{tag}_pred_var"][d]).max() < 1e-9 * max(1.0, np.abs(g[f"rgp_{tag}_pred_var"][d]).max())
        # predict_using_y == K(x*,X) K_x^-1 y == sum_j k(x*,X_j) alpha_j  (the form the OCP model uses)
        y = g[f"rgp_{tag}_puy_y"][d]
        a = orc.rgp_alpha(gp.Kx_inv[d], y)
        k = th[d, 1] ** 2 * np.exp(-0.5 * (g[f"rgp_{tag}_pred_x"][:, None] - X[d][None, :]) ** 2 / th[d, 0] ** 2)
        assert rel_err(k @ a, g[f"rgp_{tag}_puy_mean"][d]) < 1e-9


def test_drag_reference_code(golden):
    g = golden("reference_code")
    for i in range(len(g["drag_x_now"])):
        vb, ad = orc.compute_a_drag(g["drag_x_now"][i], g["drag_x_pred"][i], float(g["drag_dt"]))
        assert np.abs(vb - g["drag_v_body"][i]).max() < 1e-13
        assert np.abs(ad - g["drag_a_drag"][i]).max() < 1e-11


def test_reference_chunk_reference_code(golden):
    g = golden("reference_code")
    for n, (idx, Nn, skip) in enumerate(g["chunk_cases"]):
        out = orc.reference_chunk(g["chunk_traj"], idx, Nn, skip)
        assert out.shape == g[f"chunk_{n}"].shape, (idx, Nn, skip)
        assert np.array_equal(out, g[f"chunk_{n}"]), (idx, Nn, skip)


@pytest.mark.parametrize("tag,quad", [("hb", orc.quad_hummingbird()), ("log", orc.quad_logged_pysim())])
def test_plant_reference_code(golden, tag, quad):
    g = golden("reference_code")
    for i in range(len(g[f"plant_{tag}_x"])):
        xn, nsub = orc.plant_period(quad, g[f"plant_{tag}_x"][i], g[f"plant_{tag}_u"][i], 0.05)
        assert nsub == 11          # float-accumulation quirk
        assert np.abs(xn - g[f"plant_{tag}_xnext"][i]).max() < 1e-12


@pytest.mark.parametrize("tag", ["m10", "m20", "m7"])
def test_rgp_learn_matches_reference_code(golden, tag):
    """orc_rgp_learn against RGP.learn of the reference's own numpy code (RGP.py:332-482), 12 consecutive calls:
    return value (mu_z, C_z) and the whole state (g, eta, their covariances, the re-computed K_x^-1)"""
    g = golden("rgp_learn")
    X, theta = g[f"learn_{tag}_X"], g[f"learn_{tag}_theta"]
    M = X.shape[0]
    mu_g, C_g = np.zeros(M), np.ascontiguousarray(g[f"learn_{tag}_C_g0"].copy())
    mu_eta, C_eta, C_g_eta = theta.copy(), np.eye(3), np.zeros((M, 3))
    Kxi = np.ascontiguousarray(g[f"learn_{tag}_Kx_inv0"].copy())
    for t, (xt, yt) in enumerate(zip(g[f"learn_{tag}_xt"], g[f"learn_{tag}_yt"])):
        mu_z, C_z = orc.rgp_learn(X, mu_g, C_g, mu_eta, C_eta, C_g_eta, Kxi, xt, yt)
        for name, val in (("mu_z", mu_z), ("C_z", C_z), ("mu_g", mu_g), ("C_g", C_g), ("mu_eta", mu_eta), ("C_eta", C_eta), ("Kx_inv", Kxi)):
            assert rel_err(val, g[f"learn_{tag}_{name}"][t]) < 1e-11, (tag, t, name)


def test_numpy_rgp_restatement_vs_reference_code(golden):
    """oracle/rgp_numpy.py (the 'reference numpy RGP, 1 core' timing leg of bench.py) reproduces the outputs of the
    reference's own RGP class (tests/golden/reference_code.npz, generated from /root/reference by oracle/make_golden.py)"""
    from oracle.rgp_numpy import NumpyRGP
    g = golden("reference_code")
    for tag in ("m20", "m7"):
        X, th = g[f"rgp_{tag}_X"], g[f"rgp_{tag}_theta"]
        xt, yt = g[f"rgp_{tag}_xt"], g[f"rgp_{tag}_yt"]
        models = [NumpyRGP(X[d], th[d]) for d in range(3)]
        assert rel_err(np.stack([m.K_x_inv for m in models]), g[f"rgp_{tag}_Kx_inv"]) < 1e-12
        for t in range(len(xt)):
            for d in range(3):
                models[d].regress(xt[t, d:d + 1], yt[t, d:d + 1])
            assert rel_err(np.stack([m.mu for m in models]), g[f"rgp_{tag}_mu"][t]) < 1e-12
            assert rel_err(np.stack([m.C for m in models]), g[f"rgp_{tag}_C"][t]) < 1e-12
