"""Generate tests/golden/*.npz from the reference (run in the build container only).

    python oracle/make_golden.py

Two sources:
 1. the reference's own shipped run logs under /root/reference/outputs/python_simulation/data
    (pickles written by src/execute_trajectory.py:270-275) -> per-step solver inputs/outputs;
 2. the reference's own numpy code imported unmodified through oracle/ref_shim.py
    (RGP.regress/predict, utils.compute_a_drag, utils.get_reference_chunk, Quadrotor3D.update)
    run on seeded inputs at the BASELINE shape (M=20 basis points).

TEST INFRASTRUCTURE.  The fixtures are committed; /root/reference is not needed to run the tests.
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from gp.GPE import GPEnsemble  # noqa: E402  (reference code)
from gp.RGP import RGP  # noqa: E402
from quad import Quadrotor3D  # noqa: E402
from utils import utils as ref_utils  # noqa: E402

LOGS = "/root/reference/outputs/python_simulation/data"
OUT = os.path.join(HERE, "..", "tests", "golden")


def load(name):
    with open(os.path.join(LOGS, name + ".pkl"), "rb") as f:
        return pickle.load(f)


def log_fixture(name, steps=None, rgp=False):
    d = load(name)
    n = len(d["x_odom"]) if steps is None else min(steps, len(d["x_odom"]))
    out = dict(
        x_odom=np.array(d["x_odom"][:n]), x_pred_odom=np.array(d["x_pred_odom"][:n]),
        x_ref=np.array(d["x_ref"][:len(d["x_ref"]) if steps is None else min(len(d["x_ref"]), n + 16)]),
        w_odom=np.array(d["w_odom"][:n]), cost_solution=np.array(d["cost_solution"][:n], dtype=np.float64),
        t_cpu=np.array([float(np.ravel(t)[0]) for t in d["t_cpu"][:n]]),
    )
    if rgp:
        out["rgp_X"] = np.array(d["rgp_basis_vectors"][0])                       # [3,M]
        out["rgp_theta"] = np.array([[float(np.ravel(v)[0]) for v in ax] for ax in d["rgp_theta"][0]])  # [3,3]
        out["rgp_mu"] = np.array([np.stack(m) for m in d["rgp_mu_g_t"][:n]])     # [n,3,M]
        out["rgp_C"] = np.array([np.stack(c) for c in d["rgp_C_g_t"][:n]])       # [n,3,M,M]
        out["v_body"] = np.array([np.ravel(np.stack(v)) for v in d["v_body"][:n]])   # [n,3]
        out["a_drag"] = np.array([np.ravel(np.stack(v)) for v in d["a_drag"][:n]])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def ref_code_fixtures():
    rng = np.random.default_rng(20261017)
    out = {}
    # --- RGP.regress / predict at the BASELINE shape (M=20, theta of execute_trajectory.py:106)
    for tag, M, vmax, theta, T in [("m20", 20, 10.0, [3.0, 0.1, 0.01], 40), ("m50", 50, 15.0, [3.0, 0.5, 0.01], 10),
                                   ("m7", 7, 5.0, [1.0, 0.1, 0.1], 40)]:
        gpe = GPEnsemble.fromrange([(-vmax, vmax)] * 3, [M] * 3, theta=theta)
        xt = rng.uniform(-vmax * 1.1, vmax * 1.1, size=(T, 3))
        yt = -0.3 * xt - 0.01 * xt * np.abs(xt) + 0.05 * rng.standard_normal((T, 3))
        mus, Cs = [], []
        for t in range(T):
            mu, Cm = gpe.regress([np.array([xt[t, d]]) for d in range(3)], [np.array([yt[t, d]]) for d in range(3)])
            mus.append(np.stack(mu)); Cs.append(np.stack(Cm))
        xs = np.linspace(-vmax * 1.2, vmax * 1.2, 17)
        pm, pv = [], []
        for d in range(3):
            m_, v_ = gpe.gp[d].predict(xs, var=True)
            pm.append(m_); pv.append(v_)
        out[f"rgp_{tag}_X"] = np.stack([g.X for g in gpe.gp])
        out[f"rgp_{tag}_theta"] = np.array(gpe.get_theta(), dtype=np.float64)
        out[f"rgp_{tag}_Kx_inv"] = np.stack([g.K_x_inv for g in gpe.gp])
        out[f"rgp_{tag}_Kx"] = np.stack([g.K_x for g in gpe.gp])
        out[f"rgp_{tag}_xt"], out[f"rgp_{tag}_yt"] = xt, yt
        out[f"rgp_{tag}_mu"], out[f"rgp_{tag}_C"] = np.array(mus), np.array(Cs)
        out[f"rgp_{tag}_pred_x"], out[f"rgp_{tag}_pred_mean"], out[f"rgp_{tag}_pred_var"] = xs, np.stack(pm), np.stack(pv)
        # predict_using_y (numpy branch RGP.py:264-300): mean at xs given y
        y = rng.standard_normal((3, M))
        out[f"rgp_{tag}_puy_y"] = y
        out[f"rgp_{tag}_puy_mean"] = np.stack([gpe.gp[d].predict_using_y(xs, y[d]) for d in range(3)])
    # --- utils.compute_a_drag
    xn = rng.standard_normal((32, 13)); xp = xn + 0.1 * rng.standard_normal((32, 13))
    vb, ad = [], []
    for i in range(32):
        v, a = ref_utils.compute_a_drag(xn[i], xp[i], 0.05)
        vb.append(np.ravel(v)); ad.append(np.ravel(a))
    out["drag_x_now"], out["drag_x_pred"], out["drag_dt"] = xn, xp, np.array(0.05)
    out["drag_v_body"], out["drag_a_drag"] = np.array(vb), np.array(ad)
    # --- utils.get_reference_chunk incl. end padding and skip
    traj = rng.standard_normal((37, 13))
    cases = [(0, 10, 1), (20, 10, 1), (27, 10, 1), (28, 10, 1), (33, 10, 1), (36, 10, 1), (37, 10, 1), (40, 10, 1),
             (0, 5, 3), (20, 5, 3), (23, 5, 3), (30, 5, 3), (35, 5, 3), (36, 5, 3), (5, 20, 2), (0, 20, 1)]
    out["chunk_traj"] = traj
    out["chunk_cases"] = np.array(cases)
    for n, (idx, N, skip) in enumerate(cases):
        out[f"chunk_{n}"] = np.array(ref_utils.get_reference_chunk(traj, idx, N, skip))
    # --- Quadrotor3D.update (plant RK4 with drag) with the hummingbird and the logged constants
    for tag, setter in [("hb", "hummingbird"), ("log", "logged")]:
        quad = Quadrotor3D(payload=False, drag=True)
        if setter == "hummingbird":
            quad.mass = 0.68 + 4 * 0.009
            quad.J = np.array([0.007, 0.007, 0.012]); quad.length = 0.17
            quad.max_thrust = 838.0 ** 2 * 8.54858e-06; quad.c = 0.016
            quad.x_f = np.array([quad.length, 0, -quad.length, 0]); quad.y_f = np.array([0, quad.length, 0, -quad.length])
            quad.z_l_tau = -np.array([-quad.c, quad.c, -quad.c, quad.c])
        else:
            quad.mass = 1.0; quad.length = 0.47 / 2
            quad.x_f = np.array([quad.length, 0, -quad.length, 0]); quad.y_f = np.array([0, quad.length, 0, -quad.length])
        xs_, us_, xo_ = [], [], []
        for i in range(16):
            x = np.concatenate([rng.uniform(-2, 2, 3), [1, 0, 0, 0] + 0.2 * rng.standard_normal(4),
                                rng.uniform(-8, 8, 3), rng.uniform(-2, 2, 3)])
            u = rng.uniform(-0.1, 1.1, 4)
            quad.set_state(x.copy())
            uu = u.copy()
            t = 0.0
            while t < 0.05:                      # execute_trajectory.py:232-243
                quad.update(uu, 5e-3); t += 5e-3
            xs_.append(x); us_.append(u); xo_.append(quad.get_state(quaternion=True, stacked=True))
        out[f"plant_{tag}_x"], out[f"plant_{tag}_u"], out[f"plant_{tag}_xnext"] = np.array(xs_), np.array(us_), np.array(xo_)
    np.savez_compressed(os.path.join(OUT, "reference_code.npz"), **out)
    print("reference_code", len(out), "arrays")


def rgp_learn_fixtures():
    """RGP.learn (RGP.py:332-482) sequences from the reference's own numpy code: inputs and the full state after every call"""
    out = {}
    for tag, M, theta, vmax, seed in [("m10", 10, [1.5, 0.5, 0.1], 5.0, 3), ("m20", 20, [3.0, 0.1, 0.01], 10.0, 4),
                                      ("m7", 7, [1.0, 0.1, 0.1], 5.0, 5)]:
        rng = np.random.default_rng(seed)
        X = np.linspace(-vmax, vmax, M)
        g = RGP(X, np.zeros(M), theta=theta)
        T = 12
        xt = rng.uniform(-vmax, vmax, T)
        yt = -0.3 * xt - 0.01 * xt * np.abs(xt) + 0.05 * rng.standard_normal(T)
        rec = {k: [] for k in ("mu_g", "C_g", "mu_eta", "C_eta", "Kx_inv", "mu_z", "C_z")}
        out[f"learn_{tag}_C_g0"], out[f"learn_{tag}_Kx_inv0"] = g.C_g_t.copy(), g.K_x_inv.copy()
        for t in range(T):
            mu_z, C_z = g.learn(np.array([xt[t]]), np.array([yt[t]]))
            for k, v in (("mu_g", g.mu_g_t), ("C_g", g.C_g_t), ("mu_eta", g.mu_eta_t), ("C_eta", g.C_eta_t),
                         ("Kx_inv", g.K_x_inv), ("mu_z", mu_z), ("C_z", C_z)):
                rec[k].append(np.array(v, dtype=np.float64))
        out[f"learn_{tag}_X"], out[f"learn_{tag}_theta"] = X, np.array(theta, dtype=np.float64)
        out[f"learn_{tag}_xt"], out[f"learn_{tag}_yt"] = xt, yt
        for k, v in rec.items():
            out[f"learn_{tag}_{k}"] = np.array(v)
    np.savez_compressed(os.path.join(OUT, "rgp_learn.npz"), **out)
    print("rgp_learn", len(out), "arrays")


def rgp_predict_cov_fixtures():
    """RGP.predict(cov=True, return_Jt=True) (RGP.py:195-229) of the reference's numpy code after 25 regress calls"""
    out = {}
    rng = np.random.default_rng(77)
    for tag, M, vmax, theta in [("m20", 20, 10.0, [3.0, 0.1, 0.01]), ("m7", 7, 5.0, [1.0, 0.1, 0.1])]:
        g = RGP(np.linspace(-vmax, vmax, M), np.zeros(M), theta=theta)
        xt = rng.uniform(-vmax, vmax, 25)
        yt = -0.3 * xt + 0.05 * rng.standard_normal(25)
        for t in range(25):
            g.regress(np.array([xt[t]]), np.array([yt[t]]))
        xs = np.linspace(-vmax * 1.1, vmax * 1.1, 9)
        mu, C_p, Jt = g.predict(xs, cov=True, return_Jt=True)
        out[f"pc_{tag}_X"], out[f"pc_{tag}_theta"], out[f"pc_{tag}_xt"], out[f"pc_{tag}_yt"] = g.X, np.array(theta), xt, yt
        out[f"pc_{tag}_xs"], out[f"pc_{tag}_mean"], out[f"pc_{tag}_cov"], out[f"pc_{tag}_Jt"] = xs, mu, C_p, Jt
    np.savez_compressed(os.path.join(OUT, "rgp_predict_cov.npz"), **out)
    print("rgp_predict_cov", len(out), "arrays")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    log_fixture("traj2_v10_a10_gp0")                       # 299 steps, circle, gp0
    log_fixture("traj0_v10_a10_gp2", rgp=True)             # 121 steps, RGP M=10
    log_fixture("traj1_v15_a5_gp2", steps=60, rgp=True)    # contractive first segment only
    log_fixture("traj0_v15_a5_gp2", rgp=True)              # RGP + RK4
    log_fixture("traj2_v10_a10_gp2", steps=80, rgp=True)   # RGP stress (diverging covariance)
    ref_code_fixtures()
    rgp_learn_fixtures()
    rgp_predict_cov_fixtures()
