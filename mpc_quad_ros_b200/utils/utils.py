"""Hot-path helpers of reference src/utils/utils.py, same names and argument meaning.
compute_a_drag and get_reference_chunk run on the GPU through the C-ABI (csrc/aux_kernels.cuh); the quaternion
helpers are tensor plumbing (torch ops on whatever device the input lives on)."""
import ctypes as C

import numpy as np
import torch

from .. import _capi


def _t(a, device=None):
    t = a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float64)
    t = t.to(torch.float64)
    if device is not None:
        t = t.to(device)
    return t.contiguous()


def q_to_rot_mat(q):
    """utils.py:325-340 (un-normalised polynomial form); q [...,4] -> [...,3,3]"""
    q = _t(q)
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        torch.stack([1 - 2 * (y ** 2 + z ** 2), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x ** 2 + z ** 2), 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x ** 2 + y ** 2)], -1)], -2)


def v_dot_q(v, q):
    """utils.py:317-322"""
    return (q_to_rot_mat(q) @ _t(v).unsqueeze(-1)).squeeze(-1)


def quaternion_inverse(q):
    """utils.py:434-440"""
    q = _t(q)
    return q * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=q.dtype, device=q.device)


def skew_symmetric(v):
    """utils.py:394-412 (4x4 PAMPC form)"""
    v = _t(v)
    z = torch.zeros_like(v[..., 0])
    a, b, c = v.unbind(-1)
    return torch.stack([torch.stack([z, -a, -b, -c], -1), torch.stack([a, z, c, -b], -1),
                        torch.stack([b, -c, z, a], -1), torch.stack([c, b, -a, z], -1)], -2)


def body_velocity(x):
    """v_dot_q(x[7:10], quaternion_inverse(x[3:7])) for x [...,13]"""
    x = _t(x)
    return v_dot_q(x[..., 7:10], quaternion_inverse(x[..., 3:7]))


def get_reference_chunk(reference_trajectory, current_idx, control_nodes, skip=1, device="cuda:0"):
    """utils.py:897-931.  reference_trajectory [K,13] (numpy, batch 1) or [B,K,13] CUDA tensor -> [N,13] / [B,N,13]"""
    assert skip % 1 == 0, "Skip must be an integer"
    single = not torch.is_tensor(reference_trajectory) and np.asarray(reference_trajectory).ndim == 2
    traj = _t(reference_trajectory, device)
    if traj.ndim == 2:
        traj = traj.unsqueeze(0)
    B, K, _ = traj.shape
    out = torch.empty((B, control_nodes, 13), dtype=torch.float64, device=traj.device)
    _capi.check(_capi.lib().qmpc_reference_chunk(B, K, _capi.ptr(traj), int(current_idx), int(control_nodes), int(skip),
                                                 _capi.ptr(out), _capi.stream_ptr()))
    return out[0].cpu().numpy() if single else out


def compute_a_drag(x_now, x_pred_minus_1, dt, device="cuda:0"):
    """utils.py:934-950.  (13,) numpy inputs -> two lists of three (1,) arrays (reference format);
    [B,13] CUDA tensors -> (v_body [B,3], a_drag [B,3]) tensors."""
    single = not torch.is_tensor(x_now)
    xn, xp = _t(x_now, device).reshape(-1, 13), _t(x_pred_minus_1, device).reshape(-1, 13)
    B = xn.shape[0]
    vb = torch.empty((B, 3), dtype=torch.float64, device=xn.device)
    ad = torch.empty_like(vb)
    _capi.check(_capi.lib().qmpc_compute_a_drag(B, _capi.ptr(xn), _capi.ptr(xp), C.c_double(dt), _capi.ptr(vb),
                                                _capi.ptr(ad), _capi.stream_ptr()))
    if single:
        vb, ad = vb[0].cpu().numpy(), ad[0].cpu().numpy()
        return [np.array([vb[i]]) for i in range(3)], [np.array([ad[i]]) for i in range(3)]
    return vb, ad
