import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.swarm import SharedSwarmRGP
from mpc_quad_ros_b200.trajectory import random_smooth_trajectories
B, N, M = 1024, 20, 20
traj = random_smooth_trajectories(B, 60 + N + 2, 1.0 / N)
for shared in (False, True):
    quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
    gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=1 if shared else B)
    opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe)
    sw = SharedSwarmRGP(gpe, opt) if shared else None
    loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()), shared_swarm=sw)
    for s in range(50):
        loop.step()
        if s % 10 == 9:
            st, it = opt.solver_status(); rd = opt.solver_rounds()
            err = (loop.x[:, :3] - torch.as_tensor(traj[:, s + 1, :3]).cuda()).norm(dim=1)
            sat = ((loop.u0 <= 0) | (loop.u0 >= 1)).double().mean()
            mu = gpe.mu_tensor()
            print(f"shared={shared} step {s}: track err mean {err.mean():.3f} max {err.max():.2f} | sat {sat:.3f} | it {it.double().mean():.2f} rd {rd.double().mean():.2f} | |mu| max {mu.abs().max():.3e} | |v| mean {loop.x[:,7:10].norm(dim=1).mean():.2f}")
