"""A few closed-loop control steps at the BASELINE shape, for ncu (no CPU legs, no timing)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.trajectory import random_smooth_trajectories

B = int(os.environ.get("B", 4096)); N = int(os.environ.get("N", 20)); M = int(os.environ.get("M", 20))
steps = int(os.environ.get("STEPS", 6)); prec = int(os.environ.get("PREC", 64))
quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B) if M else None
opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe, precision=prec)
traj = random_smooth_trajectories(B, steps + N + 2, 1.0 / N)
loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()))
for s in range(steps):
    loop.step()
torch.cuda.synchronize()
st, it = opt.solver_status()
print("status counts", torch.bincount(st).tolist(), "ipm iters mean", it.double().mean().item(), "max", it.max().item())
