"""Per-step solver statistics of the start-up transient (the driver's bench window is steps 5..25 from the zero iterate):
hard-list size, OCPs that ran the IPM, rounds, and the device time of K1 / screening / dense launches on a single stream.
Usage: python scripts/diag_transient.py [steps]      env: BATCH, WORKLOAD=random_smooth|lemniscate, VPEAK"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_quad_ros_b200 import _capi
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.trajectory import lemniscate_trajectories, random_smooth_trajectories
B, N, M = int(os.environ.get("BATCH", 4096)), 20, 20
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
opts = sys.argv[2] if len(sys.argv) > 2 else ""
policy = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in opts.split(",") if kv}
every = int(os.environ.get("EVERY", 1))
quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B)
opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe, **policy)
if os.environ.get("WORKLOAD", "random_smooth") == "lemniscate":
    traj = lemniscate_trajectories(B, steps + N + 2, 1.0 / N, v_peak=float(os.environ.get("VPEAK", 20.0)), seed=1234)
else:
    traj = random_smooth_trajectories(B, steps + N + 2, 1.0 / N, seed=1234)
loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()))
lib = _capi.lib()
print(f"# policy [{opts}]")
print("step  hard  n_ipm  it_mean(ipm)  it_max rounds_mean  r1 r2 r3 r>3 rmax | ms_lin ms_screen ms_dense | sat_u0_frac  bad")
for s in range(steps):
    _capi.check(lib.qmpc_timing_enable(opt._h, 1))
    loop.step()
    ms_lin, ms_ipm, cnt, ms_dense = C.c_double(), C.c_double(), C.c_int(), C.c_double()
    _capi.check(lib.qmpc_timing_read(opt._h, C.byref(ms_lin), C.byref(ms_ipm), C.byref(cnt)))
    _capi.check(lib.qmpc_timing_read_dense(opt._h, C.byref(ms_dense)))
    _capi.check(lib.qmpc_timing_enable(opt._h, 0))
    hc = C.c_int()
    _capi.check(lib.qmpc_get_hard_count(opt._h, C.byref(hc), _capi.stream_ptr()))
    st, it = opt.solver_status(); rd = opt.solver_rounds()
    st, it, rd = st.cpu().numpy(), it.cpu().numpy(), rd.cpu().numpy()
    u0 = loop.u0.cpu().numpy()
    sat = float(((u0 <= 1e-9) | (u0 >= 1 - 1e-9)).mean())
    ipm = it > 0
    if s % every:
        continue
    print(f"{s:4d} {hc.value:5d} {int(ipm.sum()):6d} {it[ipm].mean() if ipm.any() else 0:8.2f} {int(it.max()):4d} {rd.mean():10.2f}   "
          f"{int((rd == 1).sum())} {int((rd == 2).sum())} {int((rd == 3).sum())} {int((rd > 3).sum())} {int(rd.max())} | "
          f"{ms_lin.value:.3f} {ms_ipm.value - ms_dense.value:.3f} {ms_dense.value:.3f} | {sat:.3f} {int((st != 0).sum())}", flush=True)
