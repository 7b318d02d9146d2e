"""How good are candidate warm-start guesses of the active set?  Records the final active set of every OCP over a closed loop
and scores, per step: the previous set as it is (what the kernels use), the previous set shifted by one node (the receding
horizon moved on), a per-OCP adaptive pick (whichever of the two was closer last step) and the better of the two in
hindsight.  score = share of OCPs whose guess is exactly right (such an OCP settles in one round), and the mean number of
wrong entries.  Usage: python scripts/diag_predictors.py [steps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_quad_ros_b200.execute_trajectory import ClosedLoop
from mpc_quad_ros_b200.gp.GPE import GPEnsemble
from mpc_quad_ros_b200.quad import Quadrotor3D
from mpc_quad_ros_b200.quad_opt import quad_optimizer
from mpc_quad_ros_b200.trajectory import random_smooth_trajectories
B, N, M = int(os.environ.get("BATCH", 4096)), 20, 20
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=B)
opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe)
traj = random_smooth_trajectories(B, steps + N + 2, 1.0 / N, seed=1234)
loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()))
acts = []
for s in range(steps):
    loop.step()
    acts.append(opt.get_active_set().cpu().numpy().reshape(B, N, 4).copy())
shift = lambda a: np.concatenate([a[:, 1:], a[:, -1:]], 1)
pick = np.zeros(B, bool)          # adaptive: True = use the shifted guess
print("step | exact-hit share: unshifted shifted adaptive best-of-two | mean wrong entries: unshifted shifted adaptive | empty sets")
for s in range(1, steps):
    prev, cur = acts[s - 1], acts[s]
    ok = (prev <= 2).all((1, 2)) & (cur <= 2).all((1, 2))
    d0 = (prev != cur).sum((1, 2)); d1 = (shift(prev) != cur).sum((1, 2))
    da = np.where(pick, d1, d0)
    if s % int(os.environ.get("EVERY", 4)) == 0 or s < 6:
        print(f"{s:4d} | {np.mean(d0[ok] == 0):.3f} {np.mean(d1[ok] == 0):.3f} {np.mean(da[ok] == 0):.3f} {np.mean(np.minimum(d0, d1)[ok] == 0):.3f} | "
              f"{d0[ok].mean():.2f} {d1[ok].mean():.2f} {da[ok].mean():.2f} | {np.mean((cur == 0).all((1, 2))):.3f}")
    pick = d1 < d0
