"""CPU check of the CUDA kernels' warp-level LOGIC: the product sources (csrc/mpc_kernels.cuh) are compiled for the
host with a lane-by-lane emulation of the CUDA subset they use (tests/emu, test-only) and compared with the oracle.
This is not a fallback: libqmpc.so never contains this build.  The real parity tests are the -m gpu ones."""
import numpy as np
import pytest

from oracle import oracle as orc
from helpers import make_gp, oracle_solve_batch, random_ocp_batch, u_rel, x_rel
from emu import emu


# the dense kernel works on the condensed Hessian (cond ~1e6..1e7): ~1e-9 relative instead of ~1e-12; north_star asks 1e-6
TOL = {1: 1e-9, 2: 2e-8}
# qmpc_config.solver_variant: 1 = Riccati kernel alone (one OCP per warp), 2 = Riccati screening launch + dense launch
VARIANTS = pytest.mark.parametrize("variant", [1, 2], ids=["warp_per_ocp", "screen_plus_dense"])


@VARIANTS
@pytest.mark.parametrize("N,use_gp", [(20, True), (10, False), (7, True)])
def test_emulated_solve_matches_oracle(N, use_gp, variant):
    B, dt = 3, 1.0 / N
    quad = orc.quad_hummingbird()
    gp = make_gp() if use_gp else None
    sc = random_ocp_batch(B, N, dt, quad, gp, seed=N)
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, None if gp is None else gp.X,
                                None if gp is None else gp.theta)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], sc["alpha"], xe, ue, variant=variant)
    xo, uo, cost, iters = oracle_solve_batch(sc, quad, dt, N, gp)
    assert (r["status"] == 0).all()
    assert u_rel(ue, uo) < 1e-7                        # fp64 tolerance of north_star: 1e-6
    assert x_rel(xe, xo) < 1e-7
    assert np.abs(r["cost"] - cost).max() < 1e-8 * max(1.0, np.abs(cost).max())
    assert np.array_equal(r["u0"], ue[:, 0, :])
    assert np.abs(xe[:, 0] - sc["x0"]).max() == 0.0


@VARIANTS
def test_emulated_solve_fp32_within_1e4(variant):
    if variant == 2:
        pytest.skip("fp32 handles use the Riccati kernel alone: the condensed Hessian (cond ~1e7) is an fp64-only formulation")
    B, N = 2, 10
    dt = 1.0 / N
    quad = orc.quad_hummingbird()
    sc = random_ocp_batch(B, N, dt, quad, None, seed=5, amp_choices=(2.0,))
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], None, xe, ue, f32=True, variant=variant)
    xo, uo, cost, iters = oracle_solve_batch(sc, quad, dt, N, None)
    assert (r["status"] == 0).all()
    # fp32 build: fp64 linearisation rounded to fp32 tiles, fp32 Riccati/IPM/active-set rounds, one step of iterative
    # refinement with an fp64 residual -> north_star's fp32 tolerance (1e-4 relative) with a wide margin
    assert u_rel(ue, uo) < 1e-4 and x_rel(xe, xo) < 1e-4, (u_rel(ue, uo), x_rel(xe, xo))


@VARIANTS
@pytest.mark.parametrize("N,M", [(1, 0), (2, 3), (3, 20)])
def test_emulated_solve_tiny_horizons(N, M, variant):
    B, dt = 2, 1.0 / N
    quad = orc.quad_hummingbird()
    gp = make_gp(M) if M else None
    sc = random_ocp_batch(B, N, dt, quad, gp, seed=N)
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, None if gp is None else gp.X,
                                None if gp is None else gp.theta)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], sc["alpha"], xe, ue, variant=variant)
    xo, uo, cost, iters = oracle_solve_batch(sc, quad, dt, N, gp)
    assert (r["status"] == 0).all() and u_rel(ue, uo) < TOL[variant] and x_rel(xe, xo) < TOL[variant]


@VARIANTS
def test_emulated_warm_start_from_previous_active_set(variant):
    """second RTI step from the first step's iterate and active set: the warm-started active-set rounds (or the IPM
    fall-back) must land on the exact minimiser again, and the returned active set is consistent with the controls"""
    B, N = 3, 20      # odd batch: the last warp of the two-OCP kernel carries an idle half
    dt = 1.0 / N
    quad = orc.quad_hummingbird()
    gp = make_gp(20)
    sc = random_ocp_batch(B, N, dt, quad, gp, seed=3, amp_choices=(8.0, 2.0))
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, gp.X, gp.theta)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r1 = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], sc["alpha"], xe, ue, variant=variant)
    act = r1["act"]
    assert (act <= 2).all()
    assert ((ue.reshape(B, -1) == 0.0) == (act == 1)).all() and ((ue.reshape(B, -1) == 1.0) == (act == 2)).all()
    sc2 = dict(sc)
    sc2["x0"] = sc["x0"] + 0.002 * np.random.default_rng(1).standard_normal(sc["x0"].shape)
    sc2["xit"], sc2["uit"] = xe.copy(), ue.copy()
    xo, uo, _, _ = oracle_solve_batch(sc2, quad, dt, N, gp)
    x2, u2 = xe.copy(), ue.copy()
    r2 = emu.solve(cfg, sc2["x0"], sc["yref"], sc["yref_e"], sc["alpha"], x2, u2, act=act.copy(), variant=variant)
    assert (r2["status"] == 0).all() and (r2["rounds"] >= 1).all()
    assert u_rel(u2, uo) < TOL[variant] and x_rel(x2, xo) < TOL[variant]


def test_emulated_dense_kernel_continues_the_screening_rounds():
    """screening limited to one round, the dense kernel continues with active-set rounds from the handed-over guess
    (qmpc_config screen_rounds=1, dense_warm_rounds=8) and falls back to its IPM: same exact minimiser"""
    B, N = 4, 20
    dt = 1.0 / N
    quad = orc.quad_hummingbird()
    gp = make_gp(20)
    sc = random_ocp_batch(B, N, dt, quad, gp, seed=3, amp_choices=(8.0, 2.0))
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, gp.X, gp.theta, screen_rounds=1, dense_warm_rounds=8)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r1 = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], sc["alpha"], xe, ue, variant=2)
    sc2 = dict(sc)
    sc2["x0"] = sc["x0"] + 0.05 * np.random.default_rng(1).standard_normal(sc["x0"].shape)   # moves the active set
    sc2["xit"], sc2["uit"] = xe.copy(), ue.copy()
    xo, uo, _, _ = oracle_solve_batch(sc2, quad, dt, N, gp)
    x2, u2 = xe.copy(), ue.copy()
    r2 = emu.solve(cfg, sc2["x0"], sc["yref"], sc["yref_e"], sc["alpha"], x2, u2, act=r1["act"].copy(), variant=2)
    assert (r2["status"] == 0).all()
    assert u_rel(u2, uo) < TOL[2] and x_rel(x2, xo) < TOL[2]
    print("dense continuation: ipm iters", r2["iters"], "rounds", r2["rounds"])


@VARIANTS
def test_emulated_breakdown_is_contained(variant):
    """a non-finite iterate in ONE vehicle: that vehicle reports status 2, keeps its iterate, holds its (clipped) previous
    first control and forgets its active set; its neighbours (same warp / same CTA loop) are solved as if it were not there"""
    B, N = 3, 20
    dt = 1.0 / N
    quad = orc.quad_hummingbird()
    sc = random_ocp_batch(B, N, dt, quad, None, seed=21, amp_choices=(2.0,))
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG)
    xo, uo, _, _ = oracle_solve_batch(sc, quad, dt, N, None)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    xe[1, 4, 8] = np.nan
    ue[1, 0] = [0.3, 1.7, -0.2, 0.5]
    r = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], None, xe, ue, variant=variant)
    assert r["status"].tolist() == [0, 2, 0]
    assert np.allclose(r["u0"][1], [0.3, 1.0, 0.0, 0.5]) and (r["act"][1] == 255).all()
    assert np.isnan(xe[1, 4, 8]) and np.array_equal(ue[1, 0], [0.3, 1.7, -0.2, 0.5])       # iterate untouched
    for b in (0, 2):
        assert np.abs(ue[b] - uo[b]).max() < TOL[variant] and np.abs(xe[b] - xo[b]).max() < 10 * TOL[variant]


@VARIANTS
def test_emulated_cold_handle_never_uses_the_remembered_active_set(variant):
    """warm_start_rounds < 0: neither the screening nor the dense kernel may look at the active set of the previous solve
    (the dense kernel used to run up to 8 warm rounds anyway).  Every OCP takes the cold IPM: iters > 0, and
    the answer is the same exact minimiser."""
    B, N = 3, 20
    dt = 1.0 / N
    quad = orc.quad_hummingbird()
    gp = make_gp(20)
    sc = random_ocp_batch(B, N, dt, quad, gp, seed=3, amp_choices=(8.0, 2.0))
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, gp.X, gp.theta, warm_start_rounds=-1)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r1 = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], sc["alpha"], xe, ue, variant=variant)
    assert (r1["iters"] > 0).all()
    sc2 = dict(sc)
    sc2["x0"] = sc["x0"] + 0.002 * np.random.default_rng(1).standard_normal(sc["x0"].shape)
    sc2["xit"], sc2["uit"] = xe.copy(), ue.copy()
    xo, uo, _, _ = oracle_solve_batch(sc2, quad, dt, N, gp)
    x2, u2 = xe.copy(), ue.copy()
    r2 = emu.solve(cfg, sc2["x0"], sc["yref"], sc["yref_e"], sc["alpha"], x2, u2, act=r1["act"].copy(), variant=variant)
    assert (r2["status"] == 0).all() and (r2["iters"] > 0).all(), (r2["iters"], r2["rounds"])
    assert u_rel(u2, uo) < TOL[variant] and x_rel(x2, xo) < TOL[variant]
    if variant == 2:
        assert r2["hard"] == B          # nothing was screened


@VARIANTS
def test_emulated_build_without_tile_ring_matches_oracle(variant):
    """the default build streams the stage tiles through a shared-memory ring (QMPC_RING=2, padded tile rows QMPC_WR=18);
    the A/B flavour without the ring (tiles read in place, unpadded rows) lands on the same minimiser from a cold solve
    and from the warm-started second solve"""
    B, N = 3, 20
    dt = 1.0 / N
    quad = orc.quad_hummingbird()
    gp = make_gp(20)
    sc = random_ocp_batch(B, N, dt, quad, gp, seed=3, amp_choices=(8.0, 2.0))
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, gp.X, gp.theta)
    xo, uo, _, _ = oracle_solve_batch(sc, quad, dt, N, gp)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r1 = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], sc["alpha"], xe, ue, variant=variant, flavour="noring")
    assert (r1["status"] == 0).all() and u_rel(ue, uo) < TOL[variant] and x_rel(xe, xo) < TOL[variant]
    sc2 = dict(sc)
    sc2["x0"] = sc["x0"] + 0.01 * np.random.default_rng(1).standard_normal(sc["x0"].shape)
    sc2["xit"], sc2["uit"] = xe.copy(), ue.copy()
    xo2, uo2, _, _ = oracle_solve_batch(sc2, quad, dt, N, gp)
    x2, u2 = xe.copy(), ue.copy()
    r2 = emu.solve(cfg, sc2["x0"], sc["yref"], sc["yref_e"], sc["alpha"], x2, u2, act=r1["act"].copy(), variant=variant, flavour="noring")
    assert (r2["status"] == 0).all() and u_rel(u2, uo2) < TOL[variant] and x_rel(x2, xo2) < TOL[variant]


@pytest.mark.parametrize("flavour", ["factor0", "factor1", "factor2"])
def test_emulated_dense_kernel_cholesky_variants(flavour):
    """the dense kernel's three blocked Cholesky variants (two CTA barriers per block column / one warp per block column with
    completion flags between the warps / one barrier per block column with a redundant diagonal tile) land on the oracle's
    minimiser: cold solve (IPM + rounds, factorisations with dR) and warm-started second solve (pinned rounds), N = 20 and an
    odd horizon whose last block column sits alone in its warp"""
    for N in (20, 7):
        B = 3
        dt = 1.0 / N
        quad = orc.quad_hummingbird()
        gp = make_gp(20)
        sc = random_ocp_batch(B, N, dt, quad, gp, seed=5, amp_choices=(8.0, 2.0))
        cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, gp.X, gp.theta)
        xo, uo, _, _ = oracle_solve_batch(sc, quad, dt, N, gp)
        xe, ue = sc["xit"].copy(), sc["uit"].copy()
        r1 = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], sc["alpha"], xe, ue, variant=2, flavour=flavour)
        assert (r1["status"] == 0).all() and u_rel(ue, uo) < TOL[2] and x_rel(xe, xo) < TOL[2]
        sc2 = dict(sc)
        sc2["x0"] = sc["x0"] + 0.05 * np.random.default_rng(2).standard_normal(sc["x0"].shape)
        sc2["xit"], sc2["uit"] = xe.copy(), ue.copy()
        xo2, uo2, _, _ = oracle_solve_batch(sc2, quad, dt, N, gp)
        x2, u2 = xe.copy(), ue.copy()
        r2 = emu.solve(cfg, sc2["x0"], sc["yref"], sc["yref_e"], sc["alpha"], x2, u2, act=r1["act"].copy(), variant=2, flavour=flavour)
        assert (r2["status"] == 0).all() and u_rel(u2, uo2) < TOL[2] and x_rel(x2, xo2) < TOL[2]


@pytest.mark.parametrize("layout", ["equispaced", "equispaced_forced_general", "scattered", "narrow_kernel", "velocities_off_grid"])
def test_emulated_gp_basis_point_layouts(layout, monkeypatch):
    """K1's GP term: an equispaced axis (linspace) is evaluated with three exps and a recurrence that starts at the basis
    point nearest to the velocity; any other layout (or EMU_GP_DIRECT, the test hook that forces it) takes one exp per kernel
    value.  Both must land on the oracle's minimiser, also with a length-scale far below the spacing and with velocities
    outside the grid."""
    B, N, M = 3, 10, 20
    dt = 1.0 / N
    rng = np.random.default_rng(9)
    quad = orc.quad_hummingbird()
    if layout == "scattered":
        gp = orc.GPSpec(np.sort(rng.uniform(-10, 10, (3, M)), axis=1), np.array((3.0, 0.1, 0.01)))
    elif layout == "narrow_kernel":
        gp = orc.GPSpec(np.tile(np.linspace(-10, 10, M), (3, 1)), np.array((0.3, 0.5, 0.01)))
    elif layout == "velocities_off_grid":
        gp = orc.GPSpec(np.tile(np.linspace(-0.5, 0.5, M), (3, 1)), np.array((0.2, 0.3, 0.01)))
    else:
        gp = make_gp(M)
    if layout == "equispaced_forced_general":
        monkeypatch.setenv("EMU_GP_DIRECT", "1")
    sc = random_ocp_batch(B, N, dt, quad, gp, seed=13)
    cfg, keep = emu.make_config(B, N, 1.0, quad, orc.W_DIAG, orc.WE_DIAG, gp.X, gp.theta)
    xe, ue = sc["xit"].copy(), sc["uit"].copy()
    r = emu.solve(cfg, sc["x0"], sc["yref"], sc["yref_e"], sc["alpha"], xe, ue, variant=1)
    xo, uo, cost, iters = oracle_solve_batch(sc, quad, dt, N, gp)
    assert (r["status"] == 0).all()
    assert u_rel(ue, uo) < 1e-8 and x_rel(xe, xo) < 1e-8, (layout, u_rel(ue, uo), x_rel(xe, xo))
