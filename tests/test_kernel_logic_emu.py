"""CPU check of the CUDA kernels' warp-level LOGIC: the product sources (csrc/mpc_kernels.cuh) are compiled for the
host with a lane-by-lane emulation of the CUDA subset they use (tests/emu, test-only) and compared with the oracle.
This is not a fallback: libqmpc.so never contains this build.  The real parity tests are the -m gpu ones."""
import numpy as np
import pytest

from oracle import oracle as orc
from helpers import make_gp, oracle_solve_batch, random_ocp_batch, u_rel, x_rel
from emu import emu


@pytest.mark.parametrize("N,use_gp", [(20, True), (10, False), (7, True)])
def test_emulated_solve_matches_oracle(N, use_gp):
    B, dt = 3, 1.0 / N
    quad = orc.quad_hummingbird()
    gp = make_gp() if use_gp else None
    sc = random_ocp_batch(B, N, dt, quad, gp, seed=N)
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, None if gp is None else gp.X,
                                None if gp is None else gp.theta)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], sc["alpha"], xe, ue)
    xo, uo, cost, iters = oracle_solve_batch(sc, quad, dt, N, gp)
    assert (r["status"] == 0).all()
    assert u_rel(ue, uo) < 1e-7                        # fp64 tolerance of north_star: 1e-6
    assert x_rel(xe, xo) < 1e-7
    assert np.abs(r["cost"] - cost).max() < 1e-8 * max(1.0, np.abs(cost).max())
    assert np.array_equal(r["u0"], ue[:, 0, :])
    assert np.abs(xe[:, 0] - sc["x0"]).max() == 0.0


def test_emulated_solve_fp32_within_1e4():
    B, N = 2, 10
    dt = 1.0 / N
    quad = orc.quad_hummingbird()
    sc = random_ocp_batch(B, N, dt, quad, None, seed=5, amp_choices=(2.0,))
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], None, xe, ue, f32=True)
    xo, uo, cost, iters = oracle_solve_batch(sc, quad, dt, N, None)
    assert (r["status"] != 2).all()
    assert u_rel(ue, uo) < 2e-2     # fp32 IPM without active-set polish: logic check only (see DESIGN.md, fp32 status)
