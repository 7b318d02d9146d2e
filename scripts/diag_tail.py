"""Per-step distribution of solver work (IPM iterations, refinement rounds) in the closed loop: who is the straggler?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.trajectory import random_smooth_trajectories
B, N, M = 4096, 20, 20
quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B)
opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe)
traj = random_smooth_trajectories(B, 70 + N + 2, 1.0 / N)
loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()))
for s in range(66):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loop.step(); e1.record(); torch.cuda.synchronize()
    st, it = opt.solver_status(); rd = opt.solver_rounds()
    w = it.double() + 0.85 * rd.double()
    top = torch.topk(w, 3)
    if s >= 8:
        print(f"step {s:2d} {e0.elapsed_time(e1):5.2f} ms | work mean {w.mean():.2f} p99 {torch.quantile(w, 0.99):.1f} max {w.max():.1f} "
              f"| top (b,it,rd): {[(int(b), int(it[b]), int(rd[b])) for b in top.indices]} | it>12: {int((it > 12).sum())} bad {int((st != 0).sum())}")
