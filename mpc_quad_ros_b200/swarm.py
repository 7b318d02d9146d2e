"""Multi-GPU plumbing: vehicle sharding and the shared-swarm RGP exchange (BASELINE configs 2/3).

Vehicles are independent, so the N-GPU path is pure sharding: rank r owns a contiguous block of vehicles and runs the
same kernels on it; there is no data-path collective.  The only exchange step exists in shared-swarm mode, where ONE
RGP drag model serves every vehicle: each rank reduces the information-form contributions of its vehicles on the GPU
(qrgp_shared_accumulate), the [3, M*M+M] fp64 block is all-reduced (NCCL over NVLink; gloo in the CPU tests), and every
rank applies the identical posterior update (qrgp_shared_apply), so no broadcast is needed.
This is new semantics (the reference never regresses more than one sample per call); it is defined as
"apply the single-sample RGP.regress (reference src/gp/RGP.py:303-330) for every vehicle of every rank, one after the
other".  The gain row J and the prior variance of a sample depend only on its input and on the fixed basis, so this
sequence of Kalman updates equals ONE information-form update with the summed J'J/r and J'y/r - exact, and independent
of the order of the vehicles.

Overlap: the residual of control step t needs only x_now and the PREVIOUS step's prediction, both known before the
solve of step t, and the solve of step t uses the model pushed after step t-1 (the reference pushes the means after its
solve, quad_opt.py:402-404).  `begin()` therefore runs residual -> accumulate -> all-reduce -> apply on a side stream
while the main stream linearises and solves; `end()` joins the two and swaps the double-buffered alpha."""
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
        dev = gpe.device
        self.info = torch.zeros((3, self.M * self.M + self.M), dtype=torch.float64, device=dev)
        # overlapped path: side stream, residual buffers, double-buffered alpha [2][3][M]
        self.side = torch.cuda.Stream(device=dev)
        self.xt = torch.zeros((quad_opt.batch, 3), dtype=torch.float64, device=dev)
        self.yt = torch.zeros((quad_opt.batch, 3), dtype=torch.float64, device=dev)
        self.alpha_buf = torch.zeros((2, 3, self.M), dtype=torch.float64, device=dev)
        self.cur = 0
        self._res_done = torch.cuda.Event()
        self._done = torch.cuda.Event()
        self._t0 = self._t1 = None          # optional timing events around the all-reduce (bench)
        self.time_allreduce = False
        self.allreduce_ms = []
        self._pending = False

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
        """one shared-model update per control step on the current stream: accumulate -> all-reduce -> apply"""
        self.accumulate(v_body, a_drag)
        self.exchange_and_apply()

    # ---- overlapped update (see module docstring) -------------------------------------------------------------
    def begin(self, x_now, x_pred_prev, first_step):
        """queue residual -> accumulate -> all-reduce -> apply of THIS control step on the side stream.  The caller then
        queues the solve on the current stream with `quad_optimizer.step(..., rgp=False)` and calls end()."""
        lib = _capi.lib()
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)                       # x_now / x_pred_prev of this step are ready
        with torch.cuda.stream(self.side):
            s = _capi.stream_ptr()
            xp = x_now if first_step else x_pred_prev     # reference: the first residual is taken against the state itself
            _capi.check(lib.qmpc_compute_a_drag(self.opt.batch, _capi.ptr(x_now), _capi.ptr(xp),
                                                C.c_double(self.opt.optimization_dt), _capi.ptr(self.xt), _capi.ptr(self.yt), s))
            self._res_done.record(self.side)
            _capi.check(lib.qrgp_shared_accumulate(self.gpe._h, self.opt.batch, _capi.ptr(self.xt), _capi.ptr(self.yt),
                                                   _capi.ptr(self.info), s))
            if self.time_allreduce:
                self._t0 = torch.cuda.Event(enable_timing=True); self._t1 = torch.cuda.Event(enable_timing=True)
                self._t0.record(self.side)
            allreduce_info(self.info, self.group)
            if self.time_allreduce:
                self._t1.record(self.side)
            _capi.check(lib.qrgp_shared_apply(self.gpe._h, _capi.ptr(self.info), s))
            nxt = self.alpha_buf[self.cur ^ 1]
            _capi.check(lib.qrgp_get_alpha(self.gpe._h, _capi.ptr(nxt), s))
            self._done.record(self.side)
        main.wait_event(self._res_done)                   # the solve's epilogue overwrites x_pred_prev
        self._pending = True

    def end(self):
        """join: later work on the current stream sees the updated model; the next solve reads the new alpha buffer"""
        assert self._pending
        torch.cuda.current_stream().wait_event(self._done)
        self.cur ^= 1
        _capi.check(_capi.lib().qmpc_bind_alpha(self.opt._h, _capi.ptr(self.alpha_buf[self.cur]), 0))
        if self.time_allreduce and self._t0 is not None:
            self.allreduce_ms.append((self._t0, self._t1))
        self._pending = False
