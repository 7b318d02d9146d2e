"""ctypes front-end of the CPU oracle (oracle/qmpc_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(mpc_quad_ros_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NX, NU, NZ = 13, 4, 17

# cost weights of the reference OCP (reference src/quad_opt.py:122-130)
W_DIAG = np.array([10, 10, 10, 0.1, 0.1, 0.1, 0.1, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.1, 0.1, 0.1, 0.1])
WE_DIAG = W_DIAG[:13].copy()


def quad_vector(mass, max_thrust, J, x_f, y_f, z_l_tau, g=(0.0, 0.0, 9.81)):
    return np.concatenate([[mass, max_thrust], J, x_f, y_f, z_l_tau, g]).astype(np.float64)


def quad_logged_pysim():
    """Constants used by the shipped python-simulation logs (m=1, L=0.235, T=20)."""
    L, c = 0.47 / 2, 0.013
    return quad_vector(1.0, 20.0, [.03, .03, .06], [L, 0, -L, 0], [0, L, 0, -L], [-c, c, -c, c])


def quad_hummingbird():
    """config/hummingbird.xacro through quad.py:395-417 ('+' layout, flipped z_l_tau)."""
    L, c = 0.17, 0.016
    T = 838.0 ** 2 * 8.54858e-06
    return quad_vector(0.68 + 4 * 0.009, T, [0.007, 0.007, 0.012], [L, 0, -L, 0], [0, L, 0, -L], [c, -c, c, -c])


PLANT_DEFAULT = np.array([0.008, 0.3, 0.3, 0.0])  # aero_drag, rotor_drag xyz (quad.py:79-88)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "qmpc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_rti_step.restype = C.c_int
        _LIB.orc_closed_loop.restype = C.c_int
        _LIB.orc_rgp_prior.restype = C.c_int
        _LIB.orc_plant_period.restype = C.c_int
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def _p(a):
    if a is None:
        return C.c_void_p(0)
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(C.c_void_p)


def _pi(a):
    if a is None:
        return C.c_void_p(0)
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _d(x):
    return C.c_double(float(x))


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class GPSpec:
    """Per-axis RGP constants: X [3,M], theta [3,3] = (L, sigma_f, sigma_n); K_x^-1 by numpy (RGP.py:156-157)."""

    def __init__(self, X, theta, Kx_inv=None):
        self.X = _c(X)
        self.M = self.X.shape[1]
        th = np.asarray(theta, dtype=np.float64)
        self.theta = _c(np.tile(th, (3, 1)) if th.ndim == 1 else th)
        if Kx_inv is None:
            Kx_inv = np.stack([rgp_prior(self.X[d], self.theta[d])[1] for d in range(3)])
        self.Kx_inv = _c(Kx_inv)

    def alpha(self, mu):
        mu = np.asarray(mu, dtype=np.float64).reshape(3, self.M)
        return _c(np.stack([rgp_alpha(self.Kx_inv[d], mu[d]) for d in range(3)]))


def f(quad, x, u, gp=None, alpha=None):
    out = np.empty(NX)
    M = gp.M if gp is not None else 0
    lib().orc_f(_p(quad), M, _p(gp.X if gp else None), _p(gp.theta if gp else None), _p(_c(alpha)),
                _p(_c(x)), _p(_c(u)), _p(out))
    return out


def rk4(quad, x, u, dt, gp=None, alpha=None):
    out = np.empty(NX)
    M = gp.M if gp is not None else 0
    lib().orc_rk4(_p(quad), M, _p(gp.X if gp else None), _p(gp.theta if gp else None), _p(_c(alpha)),
                  _p(_c(x)), _p(_c(u)), _d(dt), _p(out))
    return out


def linearize(quad, x, u, dt, gp=None, alpha=None):
    Phi, A, B = np.empty(NX), np.empty((NX, NX)), np.empty((NX, NU))
    M = gp.M if gp is not None else 0
    lib().orc_linearize(_p(quad), M, _p(gp.X if gp else None), _p(gp.theta if gp else None), _p(_c(alpha)),
                        _p(_c(x)), _p(_c(u)), _d(dt), _p(Phi), _p(A), _p(B))
    return Phi, A, B


def rti_step(quad, dt, N, x0, yref, yref_e, xit, uit, gp=None, alpha=None, Wd=W_DIAG, Wed=WE_DIAG,
             lbu=0.0, ubu=1.0, mu_tol=1e-13, max_iter=60, polish=True, return_lin=False):
    """One SQP-RTI iteration.  xit [(N+1),13], uit [N,4] are updated IN PLACE (persistent iterate).
    returns dict(status, cost, iters, kkt[, lin])"""
    assert xit.shape == (N + 1, NX) and uit.shape == (N, NU)
    cost, kkt, iters = C.c_double(), C.c_double(), C.c_int()
    M = gp.M if gp is not None else 0
    lin = np.empty((N, NX + NX * NX + NX * NU)) if return_lin else None
    st = lib().orc_rti_step(_p(quad), _d(dt), N, M, _p(gp.X if gp else None), _p(gp.theta if gp else None),
                            _p(_c(alpha)), _p(_c(Wd)), _p(_c(Wed)), _d(lbu), _d(ubu),
                            _p(_c(x0)), _p(_c(yref)), _p(_c(yref_e)), _p(xit), _p(uit),
                            C.byref(cost), C.byref(iters), C.byref(kkt), _d(mu_tol), int(max_iter), int(polish), _p(lin))
    out = dict(status=st, cost=cost.value, iters=iters.value, kkt=kkt.value)
    if return_lin:
        out["Phi"] = lin[:, :NX].copy()
        out["A"] = lin[:, NX:NX + NX * NX].reshape(N, NX, NX).copy()
        out["B"] = lin[:, NX + NX * NX:].reshape(N, NX, NU).copy()
    return out


def rti_step_batch(quad, dt, N, x0, yref, yref_e, xit, uit, gp=None, alpha=None, Wd=W_DIAG, Wed=WE_DIAG,
                   lbu=0.0, ubu=1.0, mu_tol=1e-13, max_iter=60, polish=True, nthreads=0):
    """B independent RTI steps (OpenMP over vehicles).  x0 [B,13], yref [B,N,17], yref_e [B,13]; xit [B,N+1,13] and
    uit [B,N,4] are updated IN PLACE; alpha [B,3,M] or [3,M] (shared).  returns dict(status [B], cost [B], iters [B], bad)"""
    B = x0.shape[0]
    assert xit.shape == (B, N + 1, NX) and uit.shape == (B, N, NU) and xit.flags.c_contiguous and uit.flags.c_contiguous
    M = gp.M if gp is not None else 0
    cost, iters, status = np.empty(B), np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
    stride = 0
    if alpha is not None:
        alpha = _c(alpha)
        stride = 3 * M if alpha.ndim == 3 else 0
    bad = lib().orc_rti_step_batch(_p(quad), _d(dt), N, M, _p(gp.X if gp else None), _p(gp.theta if gp else None),
                                   _p(alpha), int(stride), _p(_c(Wd)), _p(_c(Wed)), _d(lbu), _d(ubu), int(B),
                                   _p(_c(x0)), _p(_c(yref)), _p(_c(yref_e)), _p(xit), _p(uit), _p(cost), _pi(iters), _pi(status),
                                   _d(mu_tol), int(max_iter), int(polish), int(nthreads))
    return dict(status=status, cost=cost, iters=iters, bad=bad)


def make_yref(x_ref_chunk, u_ref=0.16):
    """quad_opt.py:295-317: yref[j] = [x_ref[j], u_ref*1_4]; yref_N = x_ref[-1]."""
    x_ref_chunk = np.asarray(x_ref_chunk, dtype=np.float64)
    N = x_ref_chunk.shape[0]
    yref = np.concatenate([x_ref_chunk, np.full((N, NU), u_ref)], axis=1)
    return np.ascontiguousarray(yref), x_ref_chunk[-1].copy()


def rgp_prior(X, theta):
    X = _c(X)
    M = X.shape[0]
    Kx, Kinv = np.empty((M, M)), np.empty((M, M))
    lib().orc_rgp_prior(M, _p(X), _p(_c(theta)), _p(Kx), _p(Kinv))
    return Kx, Kinv


def rgp_regress(X, theta, Kx_inv, mu, Cm, xt, yt):
    """in-place single-sample update of mu [M], C [M,M]"""
    lib().orc_rgp_regress(X.shape[0], _p(_c(X)), _p(_c(theta)), _p(_c(Kx_inv)), _p(mu), _p(Cm), _d(xt), _d(yt))


def rgp_learn(X, mu_g, C_g, mu_eta, C_eta, C_g_eta, Kx_inv, xt, yt):
    """RGP.learn (RGP.py:332-482) for one sample; mu_g [M], C_g [M,M], mu_eta [3], C_eta [3,3], Kx_inv [M,M] are updated
    IN PLACE (C-contiguous float64); returns (mu_z [M+3], C_z [M+3,M+3]) like the reference"""
    M = X.shape[0]
    mu_z, C_z = np.empty(M + 3), np.empty((M + 3, M + 3))
    rc = lib().orc_rgp_learn(M, _p(_c(X)), _p(mu_g), _p(C_g), _p(mu_eta), _p(C_eta), _p(_c(C_g_eta)), _p(Kx_inv),
                             _d(xt), _d(yt), _p(mu_z), _p(C_z))
    if rc:
        raise FloatingPointError("singular matrix in orc_rgp_learn")
    return mu_z, C_z


def rgp_alpha(Kx_inv, mu):
    out = np.empty(mu.shape[0])
    lib().orc_rgp_alpha(mu.shape[0], _p(_c(Kx_inv)), _p(_c(mu)), _p(out))
    return out


def rgp_predict(X, theta, Kx_inv, mu, Cm, xs):
    xs = _c(xs)
    mean, var = np.empty(xs.shape[0]), np.empty(xs.shape[0])
    lib().orc_rgp_predict(X.shape[0], _p(_c(X)), _p(_c(theta)), _p(_c(Kx_inv)), _p(_c(mu)), _p(_c(Cm)),
                          xs.shape[0], _p(xs), _p(mean), _p(var))
    return mean, var


def compute_a_drag(x_now, x_pred, dt):
    vb, ad = np.empty(3), np.empty(3)
    lib().orc_compute_a_drag(_p(_c(x_now)), _p(_c(x_pred)), _d(dt), _p(vb), _p(ad))
    return vb, ad


def plant_period(quad, x, u, dt, sim_dt=5e-3, plant=PLANT_DEFAULT):
    x = _c(x).copy()
    n = lib().orc_plant_period(_p(quad), _p(_c(plant)), _p(x), _p(_c(u)), _d(dt), _d(sim_dt))
    return x, n


def reference_chunk(traj, idx, N, skip=1):
    traj = _c(traj)
    out = np.empty((N, NX))
    lib().orc_reference_chunk(_p(traj), traj.shape[0], int(idx), int(N), int(skip), _p(out))
    return out


class ClosedLoop:
    """B independent vehicles, persistent controller state, order of execute_trajectory.py:196-277."""

    def __init__(self, quad, dt, N, traj, x_init, gp=None, plant=PLANT_DEFAULT, sim_dt=5e-3, u_ref=0.16,
                 mu_tol=1e-13, max_iter=60, polish=True, nthreads=0, reset_on_fail=False):
        self.quad, self.dt, self.N, self.gp = _c(quad), float(dt), int(N), gp
        self.traj = _c(traj)
        self.B, self.K = self.traj.shape[0], self.traj.shape[1]
        self.x = _c(x_init).copy()
        self.xit = np.zeros((self.B, N + 1, NX))
        self.uit = np.zeros((self.B, N, NU))
        self.M = gp.M if gp is not None else 0
        if gp is not None:
            self.mu = np.zeros((self.B, 3, self.M))
            self.C = np.ascontiguousarray(np.broadcast_to(
                np.stack([rgp_prior(gp.X[d], gp.theta[d])[0] for d in range(3)]), (self.B, 3, self.M, self.M)))
        else:
            self.mu = self.C = None
        self.xpred_prev = np.zeros((self.B, NX))
        self.have_pred = np.zeros(self.B, dtype=np.int32)
        self.plant, self.sim_dt, self.u_ref = _c(plant), float(sim_dt), float(u_ref)
        self.mu_tol, self.max_iter, self.polish, self.nthreads = mu_tol, max_iter, polish, nthreads
        self.reset_on_fail = bool(reset_on_fail)     # False = reference semantics (solver status ignored, quad_opt.py:333)
        self.step_idx = 0

    def run(self, steps, log=True):
        B, N = self.B, self.N
        u0 = np.empty((steps, B, NU)) if log else None
        xl = np.empty((steps, B, NX)) if log else None
        cl = np.empty((steps, B)) if log else None
        il = np.empty((steps, B), dtype=np.int32) if log else None
        gp = self.gp
        bad = lib().orc_closed_loop(
            _p(self.quad), _p(self.plant), _d(self.dt), _d(self.sim_dt), N, self.M,
            _p(gp.X if gp else None), _p(gp.theta if gp else None), _p(gp.Kx_inv if gp else None), int(gp is not None),
            _p(_c(W_DIAG)), _p(_c(WE_DIAG)), _d(self.u_ref), B, self.K, _p(self.traj), self.step_idx, int(steps),
            _p(self.x), _p(self.xit), _p(self.uit), _p(self.mu), _p(self.C), _p(self.xpred_prev), _pi(self.have_pred),
            _p(u0), _p(xl), _p(cl), _pi(il), _d(self.mu_tol), int(self.max_iter), int(self.polish), int(self.nthreads), int(self.reset_on_fail))
        self.step_idx += steps
        return dict(bad=bad, u0=u0, x=xl, cost=cl, iters=il)


def max_threads():
    return lib().orc_max_threads()
