"""Per-vehicle timeline of the solver kernel (qmpc_timeline_*): when does each OCP start and end inside one launch,
how long do typical / straggler OCPs run, and how many are in flight over time.  Usage: python scripts/diag_timeline.py [steps]"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_quad_ros_b200 import _capi
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.trajectory import random_smooth_trajectories
B, N, M = int(os.environ.get("BATCH", 4096)), 20, 20
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B)
opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe)
traj = random_smooth_trajectories(B, steps + 10 + N + 2, 1.0 / N)
loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()))
lib = _capi.lib()
for s in range(steps):
    loop.step()
torch.cuda.synchronize()
_capi.check(lib.qmpc_timeline_enable(opt._h, 1))
for rep in range(3):
    loop.step()
    tl = np.zeros((B, 2), dtype=np.int64)
    _capi.check(lib.qmpc_timeline_read(opt._h, tl.ctypes.data_as(C.c_void_p)))
    st, it = opt.solver_status(); rd = opt.solver_rounds()
    it, rd = it.cpu().numpy(), rd.cpu().numpy()
    t0 = tl[:, 0].min()
    s_us, e_us = (tl[:, 0] - t0) / 1e3, (tl[:, 1] - t0) / 1e3
    dur = e_us - s_us
    print(f"--- step {steps + rep}: kernel span {e_us.max():.0f} us; start p50 {np.median(s_us):.0f} max {s_us.max():.0f}; "
          f"dur p10 {np.quantile(dur, .1):.0f} p50 {np.median(dur):.0f} p90 {np.quantile(dur, .9):.0f} p99 {np.quantile(dur, .99):.0f} max {dur.max():.0f}")
    grid = np.linspace(0, e_us.max(), 21)
    infl = [(int(((s_us <= t) & (e_us > t)).sum())) for t in grid]
    print("in flight @5% steps:", infl)
    for lo, hi in ((0, 0), (1, 5), (6, 12), (13, 25), (26, 99)):
        m = (it >= lo) & (it <= hi)
        if m.any():
            print(f"  ipm iters {lo:2d}-{hi:2d}: n={int(m.sum()):5d} rounds mean {rd[m].mean():.2f} dur mean {dur[m].mean():7.0f} us max {dur[m].max():7.0f}")
    print("  rounds histogram (it==0):", np.bincount(rd[it == 0], minlength=8).tolist(), " (it>0):", np.bincount(rd[it > 0], minlength=12).tolist())
    for r in range(1, 7):
        m = (it == 0) & (rd == r)
        if m.any(): print(f"    it==0 rounds=={r}: dur p50 {np.median(dur[m]):.0f} us")
    top = np.argsort(-dur)[:6]
    print("  slowest:", [(int(b), int(it[b]), int(rd[b]), int(dur[b]), int(s_us[b])) for b in top], "(b, it, rd, dur_us, start_us)")
