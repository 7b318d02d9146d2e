"""mpc_quad_ros_b200 — B200-native batched RTI-MPC + recursive-GP quadrotor controller.

Host-side mirror of the reference's call surface (smidmatej/mpc_quad_ros):
    quad_opt.quad_optimizer, gp.GPE.GPEnsemble, gp.RGP.RGP, quad.Quadrotor3D, utils.utils
over the C-ABI library csrc/libqmpc.so (include/qmpc.h).  PyTorch is used for device memory and streams only.
There is no CPU path: every compute call needs the CUDA library and a GPU.
"""
from . import _capi  # noqa: F401

__all__ = ["quad_opt", "quad", "gp", "utils", "execute_trajectory", "trajectory", "swarm"]
