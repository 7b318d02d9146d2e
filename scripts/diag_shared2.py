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
from mpc_quad_ros_b200 import _capi
import ctypes as C
B, N, M = 1024, 20, 20
traj = random_smooth_trajectories(B, 30 + N + 2, 1.0 / N)
quad = Quadrotor3D(drag=True, batch=B).set_hummingbird_params()
gpe = GPEnsemble.fromrange([(-10, 10)] * 3, [M] * 3, theta=[3.0, 0.1, 0.01], batch=1)
opt = quad_optimizer(quad, t_horizon=1.0, n_nodes=N, gpe=gpe)
sw = SharedSwarmRGP(gpe, opt)
loop = ClosedLoop(quad, opt, torch.as_tensor(traj), torch.as_tensor(traj[:, 0, :].copy()), shared_swarm=sw)
lib = _capi.lib()
lib.qmpc_residual_x_device.restype = C.c_void_p; lib.qmpc_residual_y_device.restype = C.c_void_p
for s in range(14):
    loop.step()
    torch.cuda.synchronize()
    mu = gpe.mu_tensor(); Cm = gpe.C_tensor(); al = gpe.alpha_tensor()
    import ctypes
    yt = torch.empty((B,3),dtype=torch.float64,device="cuda"); xt=torch.empty_like(yt)
    ctypes.cdll.LoadLibrary("libcudart.so").cudaMemcpy(ctypes.c_void_p(yt.data_ptr()), ctypes.c_void_p(lib.qmpc_residual_y_device(opt._h)), B*24, 3)
    ctypes.cdll.LoadLibrary("libcudart.so").cudaMemcpy(ctypes.c_void_p(xt.data_ptr()), ctypes.c_void_p(lib.qmpc_residual_x_device(opt._h)), B*24, 3)
    st,it=opt.solver_status()
    print(f"   |yt| max {yt.abs().max().item():.3e} |xt| max {xt.abs().max().item():.3e} status bad {(st!=0).sum().item()} nan-x vehicles {(~torch.isfinite(loop.x).all(dim=1)).sum().item()}")
    print(f"step {s}: x finite {torch.isfinite(loop.x).all().item()} u0 finite {torch.isfinite(loop.u0).all().item()} info finite {torch.isfinite(sw.info).all().item()} "
          f"|info| {sw.info.abs().max().item():.3e} mu finite {torch.isfinite(mu).all().item()} |mu| {mu.abs().max().item():.3e} C finite {torch.isfinite(Cm).all().item()} "
          f"|C| {Cm.abs().max().item():.3e} Cmin diag {torch.diagonal(Cm[0,0]).min().item():.3e} alpha {al.abs().max().item():.3e}")
