"""TEST-ONLY: builds tests/emu/libqmpc_emu.so (the product kernels compiled for the host) and calls it."""
import ctypes as C
import os
import subprocess

import numpy as np

from mpc_quad_ros_b200._capi import QmpcConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}
# build flavours of the emulated kernels: name -> (extra compiler defines, tile row stride)
FLAVOURS = {"": ([], 18), "noring": (["-DQMPC_RING=0", "-DQMPC_WR=16"], 16), "trace": (["-DQMPC_EMU_TRACE"], 18),
            # the dense kernel's other Cholesky variants (QMPC_DENSE_FACTOR in mpc_kernels_dense.cuh)
            "factor0": (["-DQMPC_DENSE_FACTOR=0", "-DQMPC_DENSE_SCALED_SOLVE=0", "-DQMPC_DENSE_GPLANES=0"], 18), "factor1": (["-DQMPC_DENSE_FACTOR=1"], 18), "factor2": (["-DQMPC_DENSE_FACTOR=2"], 18)}


def lib(flavour=""):
    if flavour not in _LIBS:
        defs, _ = FLAVOURS[flavour]
        so = os.path.join(_HERE, "libqmpc_emu%s.so" % ("_" + flavour if flavour else ""))
        srcs = [os.path.join(_HERE, f) for f in ("emu_kernels.cpp", "emu_cuda.h")]
        csrc = os.path.join(_HERE, "..", "..", "mpc_quad_ros_b200", "csrc")
        srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-I", _HERE] + defs +
                                  ["-o", so, os.path.join(_HERE, "emu_kernels.cpp")])
        _LIBS[flavour] = C.CDLL(so)
    return _LIBS[flavour]


def make_config(B, N, t_horizon, quad, w_diag, we_diag, gp_X=None, gp_theta=None, mu_tol=0.0, max_iter=0,
                lbu=0.0, ubu=1.0, **policy):
    """policy: any of the qmpc_config solver-policy fields (solver_variant, screen_rounds, warm_start_rounds, ...)"""
    c = QmpcConfig()
    c.batch, c.n_nodes, c.precision, c.device = B, N, 64, 0
    c.n_basis = 0 if gp_X is None else gp_X.shape[1]
    c.ipm_max_iter, c.ipm_mu_tol, c.t_horizon = max_iter, mu_tol, t_horizon
    c.quad[:] = list(quad); c.w_diag[:] = list(w_diag); c.we_diag[:] = list(we_diag)
    c.lbu, c.ubu = lbu, ubu
    for k, v in policy.items():
        assert hasattr(c, k), k
        setattr(c, k, v)
    keep = None
    if gp_X is not None:
        keep = np.ascontiguousarray(gp_X, dtype=np.float64)
        c.gp_theta[:] = list(np.asarray(gp_theta, dtype=np.float64).ravel())
        c.gp_X = keep.ctypes.data_as(C.POINTER(C.c_double))
    return c, keep


def solve(cfg, x0, yref, yref_e, alpha, xit, uit, f32=False, act=None, variant=None, flavour=""):
    """variant: None = cfg.solver_variant as given; 1 = Riccati kernel alone, 2 = screening + dense kernel;
    flavour: build flavour of the emulated kernels (FLAVOURS)"""
    B, N = cfg.batch, cfg.n_nodes
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    x0, yref, yref_e = (np.ascontiguousarray(a, dtype=np.float64) for a in (x0, yref, yref_e))
    alpha = np.zeros((B, 3, max(cfg.n_basis, 1))) if alpha is None else np.ascontiguousarray(alpha, dtype=np.float64)
    u0, cost = np.empty((B, 4)), np.empty(B)
    status, iters, rounds = np.empty(B, dtype=np.int32), np.zeros(B, dtype=np.int32), np.empty(B, dtype=np.int32)   # iters: also an input (previous solve)
    act = np.full((B, 4 * N), 255, dtype=np.uint8) if act is None else act
    W = np.empty((B, N, 13, FLAVOURS[flavour][1]), dtype=np.float32 if f32 else np.float64)
    saved_variant, saved_prec = cfg.solver_variant, cfg.precision
    if variant is not None:
        cfg.solver_variant = variant
    cfg.precision = 32 if f32 else 64
    hard = C.c_int(0)
    fn = getattr(lib(flavour), "emu_solve_%s" % ("f32" if f32 else "f64"))
    fn(C.byref(cfg), p(x0), p(yref), p(yref_e), p(alpha), p(xit), p(uit), p(u0), p(cost), p(status), p(iters), p(rounds), p(act), p(W),
       C.byref(hard))
    cfg.solver_variant, cfg.precision = saved_variant, saved_prec
    return dict(u0=u0, cost=cost, status=status, iters=iters, rounds=rounds, act=act, W=W, hard=hard.value)
