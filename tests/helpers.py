"""Scenario builders shared by the CPU (emulation) and GPU parity tests."""
import numpy as np

from oracle import oracle as orc


def make_gp(M=20, vmax=10.0, theta=(3.0, 0.1, 0.01)):
    return orc.GPSpec(np.tile(np.linspace(-vmax, vmax, M), (3, 1)), np.array(theta))


def random_ocp_batch(B, N, dt, quad, gp=None, seed=0, amp_choices=(2.0, 8.0, 25.0), zero_iterate_frac=0.1):
    """B independent single-step RTI problems: a dynamically plausible iterate (rollout of perturbed hover inputs),
    a perturbed x0, a ramp reference (the larger amplitudes saturate the thrust bounds), random RGP means.
    returns dict(x0, yref, yref_e, xit, uit, alpha, mu)"""
    rng = np.random.default_rng(seed)
    M = gp.M if gp is not None else 0
    xit = np.zeros((B, N + 1, 13)); uit = np.zeros((B, N, 4))
    x0 = np.zeros((B, 13)); yref = np.zeros((B, N, 17)); yref_e = np.zeros((B, 13))
    mu = 0.5 * rng.standard_normal((B, 3, max(M, 1)))
    alpha = np.stack([gp.alpha(mu[b]) for b in range(B)]) if gp is not None else None
    for b in range(B):
        start = np.concatenate([rng.uniform(-1, 1, 2), [rng.uniform(2, 4)], [1, 0, 0, 0], rng.uniform(-1, 1, 3), [0, 0, 0]])
        if rng.uniform() > zero_iterate_frac:
            uit[b] = np.clip(0.3 + 0.05 * rng.standard_normal((N, 4)), 0.0, 1.0)
            xit[b, 0] = start
            for k in range(N):
                xit[b, k + 1] = orc.rk4(quad, xit[b, k], uit[b, k], dt, gp, None if alpha is None else alpha[b])
        x0[b] = start + 0.05 * rng.standard_normal(13)
        amp = amp_choices[b % len(amp_choices)]
        xref = np.zeros((N, 13)); xref[:, 3] = 1
        d = rng.standard_normal(2); d /= np.linalg.norm(d)
        xref[:, 0] = start[0] + d[0] * amp * np.linspace(0.1, 1, N)
        xref[:, 1] = start[1] + d[1] * amp * np.linspace(0.1, 1, N)
        xref[:, 2] = start[2] + 0.3 * amp * np.linspace(0, 1, N)
        yref[b], yref_e[b] = orc.make_yref(xref)
    return dict(x0=x0, yref=yref, yref_e=yref_e, xit=xit, uit=uit, alpha=alpha, mu=mu if gp is not None else None)


def oracle_solve_batch(sc, quad, dt, N, gp=None, idx=None):
    """exact oracle answers for (a subset of) a scenario batch"""
    idx = range(sc["x0"].shape[0]) if idx is None else idx
    xo, uo, cost, iters = [], [], [], []
    for b in idx:
        x, u = sc["xit"][b].copy(), sc["uit"][b].copy()
        r = orc.rti_step(quad, dt, N, sc["x0"][b], sc["yref"][b], sc["yref_e"][b], x, u, gp=gp,
                         alpha=None if gp is None else sc["alpha"][b])
        assert r["status"] == 0, r
        xo.append(x); uo.append(u); cost.append(r["cost"]); iters.append(r["iters"])
    return np.array(xo), np.array(uo), np.array(cost), np.array(iters)


def u_rel(u, u_ref):
    """controls live in [0,1] -> floor 1"""
    return float(np.abs(u - u_ref).max() / max(1.0, np.abs(u_ref).max()))


def x_rel(x, x_ref):
    """states: per-block relative error with floor 1 (quaternion / small velocities)"""
    return float(np.abs(x - x_ref).max() / max(1.0, np.abs(x_ref).max()))
