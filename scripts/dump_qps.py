"""Dump the single-step inputs of every vehicle at a few closed-loop steps (solver robustness studies on the CPU)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.trajectory import random_smooth_trajectories, lemniscate_trajectories
B, N, M = 1024, 20, 20
out = {}
for wl in ("random_smooth", "lemniscate20"):
    quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
    gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B)
    opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe)
    traj = random_smooth_trajectories(B, 80 + N + 2, 1.0 / N) if wl == "random_smooth" else lemniscate_trajectories(B, 80 + N + 2, 1.0 / N, v_peak=20.0)
    loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()))
    for s in range(75):
        if s in (3, 20, 45, 70):
            xit, uit = opt.get_iterate()
            alpha = gpe.alpha_tensor().clone()
            x_now = loop.x.clone()
            loop.step()
            st, it = opt.solver_status()
            out[f"{wl}_{s}"] = dict(x0=x_now.cpu().numpy(), chunk=loop.chunk.cpu().numpy(), alpha=alpha.cpu().numpy(),
                                    xit=xit.cpu().numpy(), uit=uit.cpu().numpy(), status=st.cpu().numpy(), iters=it.cpu().numpy(),
                                    rounds=opt.solver_rounds().cpu().numpy(), u_out=opt.get_iterate()[1].cpu().numpy())
        else:
            loop.step()
np.save(os.path.join(ROOT, "gpurun_out", "qps.npy"), np.array([out], dtype=object), allow_pickle=True)
print({k: (int((v["status"] != 0).sum()), float(v["iters"].mean())) for k, v in out.items()})
