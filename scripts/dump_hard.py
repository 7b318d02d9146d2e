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
cases = []
for s in range(62):
    xit, uit = opt.get_iterate(); alpha = gpe.alpha_tensor().clone(); x_now = loop.x.clone()
    loop.step()
    st, it = opt.solver_status(); rd = opt.solver_rounds()
    if s in (10, 20, 30, 36, 42, 50, 58, 61):
        w = it.double() + 0.85 * rd.double()
        for b in torch.topk(w, 4).indices.tolist():
            cases.append(dict(step=s, b=b, x0=x_now[b].cpu().numpy(), chunk=loop.chunk[b].cpu().numpy(), alpha=alpha[b].cpu().numpy(),
                              xit=xit[b].cpu().numpy(), uit=uit[b].cpu().numpy(), status=int(st[b]), iters=int(it[b]), rounds=int(rd[b])))
np.save(os.path.join(ROOT, "gpurun_out", "hard_cases.npy"), np.array(cases, dtype=object), allow_pickle=True)
print("saved", len(cases))
