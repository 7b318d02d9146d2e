"""Multi-GPU plumbing: vehicle sharding and the shared-swarm RGP exchange (BASELINE configs 2/3).

Vehicles are independent, so the N-GPU path is pure sharding: rank r owns a contiguous block of vehicles and runs the
same kernels on it; there is no data-path collective.  The only exchange step exists in shared-swarm mode, where ONE
RGP drag model serves every vehicle: each rank reduces the information-form contributions of its vehicles on the GPU
(qrgp_shared_accumulate), the [3, M*M+M] fp64 block is all-reduced (NCCL over NVLink; gloo in the CPU tests), and every
rank applies the identical posterior update (qrgp_shared_apply), so no broadcast is needed.
This is new semantics (the reference never regresses more than one sample per call); it is defined as
"apply the single-sample RGP.regress for every vehicle with gain and innovation evaluated at the pre-update model"."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _capi


def shard_range(total, rank, world):
    """contiguous block of ceil(total/world) vehicles for `rank` (last rank may be short): returns (first, count)"""
    per = -(-total // world)
    first = min(rank * per, total)
    return first, max(0, min(per, total - first))


def allreduce_info(info, group=None):
    """sum the information-form block over ranks in place (NCCL for CUDA tensors, gloo for CPU tensors)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(info, op=dist.ReduceOp.SUM, group=group)
    return info


class SharedSwarmRGP:
    """gpe: GPEnsemble created with batch=1 (the shared model, replicated on every rank);
    quad_opt: quad_optimizer over this rank's B vehicles, constructed with that same gpe."""

    def __init__(self, gpe, quad_opt, group=None):
        assert gpe.batch == 1, "the shared model is a batch-1 ensemble"
        assert quad_opt.gpe is gpe
        self.gpe, self.opt, self.group = gpe, quad_opt, group
        self.M = gpe.M
        self.info = torch.zeros((3, self.M * self.M + self.M), dtype=torch.float64, device=gpe.device)

    def accumulate(self, v_body=None, a_drag=None):
        """information-form sums over this rank's vehicles; default inputs = residuals left by quad_optimizer.step"""
        lib, h = _capi.lib(), self.opt._h
        B = self.opt.batch
        if v_body is None:
            lib.qmpc_residual_x_device.restype = C.c_void_p
            lib.qmpc_residual_y_device.restype = C.c_void_p
            xt, yt = C.c_void_p(lib.qmpc_residual_x_device(h)), C.c_void_p(lib.qmpc_residual_y_device(h))
        else:
            v_body, a_drag = v_body.contiguous(), a_drag.contiguous()
            B = v_body.shape[0]
            xt, yt = _capi.ptr(v_body), _capi.ptr(a_drag)
        _capi.check(lib.qrgp_shared_accumulate(self.gpe._h, B, xt, yt, _capi.ptr(self.info), _capi.stream_ptr()))
        return self.info

    def exchange_and_apply(self):
        allreduce_info(self.info, self.group)
        _capi.check(_capi.lib().qrgp_shared_apply(self.gpe._h, _capi.ptr(self.info), _capi.stream_ptr()))

    def update(self, v_body=None, a_drag=None):
        """one shared-model update per control step: accumulate -> all-reduce -> apply"""
        self.accumulate(v_body, a_drag)
        self.exchange_and_apply()
