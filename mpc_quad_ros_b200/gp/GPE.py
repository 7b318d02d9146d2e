"""GPEnsemble — three RGPs (body axes x, y, z) per vehicle, B vehicles (reference src/gp/GPE.py:34-268).
State (mu [B,3,M], C [B,3,M,M]) is owned by a qrgp handle of libqmpc.so; regress / predict are CUDA kernels."""
import ctypes as C

import numpy as np
import torch

from .. import _capi
from .RGP import RGP, _Clone


class GPEnsemble:
    def __init__(self, gp_list, type, batch=1, device="cuda:0"):
        if type != "RGP":
            raise NotImplementedError("only the RGP ensemble is part of the control-step path (GPE.py type 'GP' is offline)")
        assert len(gp_list) == 3, "one model per body axis"
        self.gp, self.type = gp_list, type
        self.batch, self.device = batch, torch.device(device)
        M = gp_list[0].X.shape[0]
        assert all(g.X.shape[0] == M for g in gp_list), "all axes must share the number of basis points"
        self.M = M
        self.X = np.ascontiguousarray(np.stack([g.X for g in gp_list]), dtype=np.float64)
        self.theta = np.ascontiguousarray(np.array([g.get_theta() for g in gp_list]), dtype=np.float64)
        self.K_x = np.ascontiguousarray(np.stack([g.K_x for g in gp_list]))
        self.K_x_inv = np.ascontiguousarray(np.stack([g.K_x_inv for g in gp_list]))
        for d, g in enumerate(gp_list):
            if isinstance(g, RGP):
                g._ens, g._axis = self, d
        self._h = C.c_void_p()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _capi.check(_capi.lib().qrgp_create(batch, M, p(self.X), p(self.theta), p(self.K_x), p(self.K_x_inv),
                                            self.device.index or 0, C.byref(self._h)))
        y0 = np.stack([g.y_ for g in gp_list])
        C0 = [getattr(g, "_C0", None) for g in gp_list]
        if np.any(y0 != 0) or any(c is not None for c in C0):
            mu = torch.as_tensor(np.broadcast_to(y0, (batch, 3, M)).copy(), device=self.device)
            Cm = np.stack([C0[d] if C0[d] is not None else self.K_x[d] for d in range(3)])
            Cm = torch.as_tensor(np.broadcast_to(Cm, (batch, 3, M, M)).copy(), device=self.device)
            self.set_state(mu, Cm)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _capi.lib().qrgp_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- constructors (GPE.py:46-150) -----------------------------------------------------------------------
    @classmethod
    def fromlist(cls, gp_list=[], batch=1, device="cuda:0"):
        if not all(isinstance(g, RGP) for g in gp_list):
            raise ValueError("All GP objects in the list must be of the same type")
        return cls(gp_list, "RGP", batch, device)

    @classmethod
    def frombasisvectors(cls, X, y, C, theta, batch=1, device="cuda:0"):
        assert len(X) == 3 and len(y) == 3 and len(C) == 3 and len(theta) == 3
        return cls([_Spec(X[i], y[i], C[i], theta[i]) for i in range(3)], "RGP", batch, device)._wrap()

    @classmethod
    def fromemptybasisvectors(cls, X, batch=1, device="cuda:0"):
        assert len(X) == 3, "X must have length 3"
        return cls([_Spec(X[i], np.zeros(X[i].shape[0]), None, [1.0, 0.1, 0.1]) for i in range(3)], "RGP", batch, device)._wrap()

    @classmethod
    def fromrange(cls, x_min_max, n_basis, theta=None, batch=1, device="cuda:0"):
        assert len(x_min_max) == 3 and len(n_basis) == 3
        th = [1.0, 0.1, 0.1] if theta is None else theta
        specs = []
        for i in range(3):
            Xb = np.linspace(x_min_max[i][0], x_min_max[i][1], n_basis[i])
            specs.append(_Spec(Xb, np.zeros(Xb.shape[0]), None, th))
        return cls(specs, "RGP", batch, device)._wrap()

    def _wrap(self):
        """replace the light-weight specs by RGP views bound to this ensemble"""
        views = []
        for d, s in enumerate(self.gp):
            v = RGP.__new__(RGP)
            v.X, v.y_, v.theta, v.sigma_n = s.X, s.y_, list(s.theta), s.theta[2]
            v.K_x, v.K_x_inv, v._C0 = s.K_x, s.K_x_inv, s._C0
            from .RGP import RBF
            v.K = RBF(L=np.eye(1) * s.theta[0], sigma_f=s.theta[1])
            v._ens, v._axis = self, d
            views.append(v)
        self.gp = views
        return self

    def get_theta(self):
        return [self.gp[n].get_theta() for n in range(len(self.gp))]

    # ---- device state ------------------------------------------------------------------------------------------
    def mu_tensor(self):
        out = torch.empty((self.batch, 3, self.M), dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib().qrgp_get_mu(self._h, _capi.ptr(out), _capi.stream_ptr()))
        return out

    def C_tensor(self):
        out = torch.empty((self.batch, 3, self.M, self.M), dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib().qrgp_get_C(self._h, _capi.ptr(out), _capi.stream_ptr()))
        return out

    def alpha_tensor(self):
        out = torch.empty((self.batch, 3, self.M), dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib().qrgp_get_alpha(self._h, _capi.ptr(out), _capi.stream_ptr()))
        return out

    def set_state(self, mu=None, Cm=None):
        mu = None if mu is None else mu.to(self.device, torch.float64).contiguous()
        Cm = None if Cm is None else Cm.to(self.device, torch.float64).contiguous()
        _capi.check(_capi.lib().qrgp_set_state(self._h, _capi.ptr(mu), _capi.ptr(Cm), _capi.stream_ptr()))

    def _mu_axis(self, d):
        mu = self.mu_tensor()[:, d]
        return mu[0].cpu().numpy() if self.batch == 1 else mu

    def _C_axis(self, d):
        Cm = self.C_tensor()[:, d]
        return Cm[0].cpu().numpy() if self.batch == 1 else Cm

    # ---- regress (GPE.py:244-268) ------------------------------------------------------------------------------------
    def _to_b3(self, v):
        if torch.is_tensor(v):
            t = v.to(self.device, torch.float64).reshape(self.batch, 3)
        else:
            t = torch.as_tensor(np.array([np.ravel(a) for a in v], dtype=np.float64).T.copy(), device=self.device)
            t = t.reshape(self.batch, 3)
        return t.contiguous()

    def regress(self, X_t, y_t):
        """list-in/list-out like the reference for batch 1 (X_t, y_t lists of three (1,) arrays -> (list mu, list C));
        tensors [B,3] in -> (mu [B,3,M], C [B,3,M,M]) out."""
        batched = torch.is_tensor(X_t)
        if not batched:
            assert len(X_t) == len(self.gp) and len(y_t) == len(self.gp)
            assert all(isinstance(a, np.ndarray) for a in X_t) and all(isinstance(a, np.ndarray) for a in y_t)
        xt, yt = self._to_b3(X_t), self._to_b3(y_t)
        _capi.check(_capi.lib().qrgp_regress(self._h, _capi.ptr(xt), _capi.ptr(yt), _capi.stream_ptr()))
        mu, Cm = self.mu_tensor(), self.C_tensor()
        if batched or self.batch > 1:
            return mu, Cm
        mu, Cm = mu[0].cpu().numpy(), Cm[0].cpu().numpy()
        return [mu[d] for d in range(3)], [Cm[d] for d in range(3)]

    def regress_from_states(self, x_now, x_pred_prev, dt):
        """fused utils.compute_a_drag + regress for [B,13] CUDA tensors; returns (v_body, a_drag) [B,3]"""
        vb = torch.empty((self.batch, 3), dtype=torch.float64, device=self.device)
        ad = torch.empty_like(vb)
        _capi.check(_capi.lib().qrgp_regress_from_states(self._h, _capi.ptr(x_now), _capi.ptr(x_pred_prev), C.c_double(dt),
                                                         _capi.ptr(vb), _capi.ptr(ad), _capi.stream_ptr()))
        return vb, ad

    def _regress_axis(self, d, Xt, yt):
        xt = torch.full((self.batch, 3), float("nan"), dtype=torch.float64, device=self.device)
        y = torch.zeros_like(xt)
        xt[:, d] = torch.as_tensor(np.ravel(Xt), dtype=torch.float64).to(self.device)
        y[:, d] = torch.as_tensor(np.ravel(yt), dtype=torch.float64).to(self.device)
        _capi.check(_capi.lib().qrgp_regress(self._h, _capi.ptr(xt), _capi.ptr(y), _capi.stream_ptr()))

    # ---- predict (GPE.py:165-241) --------------------------------------------------------------------------------------
    def _queries(self, X_t):
        if torch.is_tensor(X_t):
            return X_t.to(self.device, torch.float64).reshape(self.batch, 3, -1).contiguous()
        q = np.stack([np.ravel(a) for a in X_t]).astype(np.float64)
        return torch.as_tensor(np.broadcast_to(q, (self.batch,) + q.shape).copy(), device=self.device)

    def predict_tensor(self, xs, want_var=False):
        xs = xs.contiguous()
        mean = torch.empty_like(xs)
        var = torch.empty_like(xs) if want_var else None
        _capi.check(_capi.lib().qrgp_predict(self._h, xs.shape[2], _capi.ptr(xs), _capi.ptr(mean), _capi.ptr(var),
                                             _capi.stream_ptr()))
        return mean, var

    def predict(self, X_t, std=False):
        assert torch.is_tensor(X_t) or len(X_t) == len(self.gp)
        mean, var = self.predict_tensor(self._queries(X_t), want_var=std)
        if torch.is_tensor(X_t) or self.batch > 1:
            return (mean, var.sqrt()) if std else mean
        mu = [mean[0, d].cpu().numpy() for d in range(3)]
        if std:
            return mu, [var[0, d].sqrt().cpu().numpy() for d in range(3)]
        return mu

    def predict_using_y(self, X_t, y, std=False):
        if std:
            raise NotImplementedError("predict_using_y returns the mean on the GPU path")
        xs = self._queries(X_t)
        if torch.is_tensor(y):
            yy = y.to(self.device, torch.float64).reshape(self.batch, 3, self.M).contiguous()
        else:
            yy = np.stack([np.ravel(a) for a in y]).astype(np.float64)
            yy = torch.as_tensor(np.broadcast_to(yy, (self.batch, 3, self.M)).copy(), device=self.device)
        mean = torch.empty_like(xs)
        _capi.check(_capi.lib().qrgp_predict_using_y(self._h, xs.shape[2], _capi.ptr(xs), _capi.ptr(yy), _capi.ptr(mean),
                                                     _capi.stream_ptr()))
        if torch.is_tensor(X_t) or self.batch > 1:
            return mean
        return np.stack([mean[0, d].cpu().numpy() for d in range(3)], axis=1)    # (m,3) like np.concatenate(axis=1)

    def _predict_axis(self, d, xs, want_var):
        q = np.zeros((3, xs.shape[0]))
        q[d] = xs
        mean, var = self.predict_tensor(self._queries(list(q)), want_var)
        return mean[0, d].cpu().numpy(), (var[0, d].cpu().numpy() if want_var else None)

    def predict_cov_tensor(self, xs):
        """gains Jt [B,3,m,M] and full posterior covariance [B,3,m,m] at the queries xs [B,3,m]"""
        xs = xs.contiguous()
        B, _, m = xs.shape
        Jt = torch.empty((B, 3, m, self.M), dtype=torch.float64, device=self.device)
        cov = torch.empty((B, 3, m, m), dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib().qrgp_predict_cov(self._h, m, _capi.ptr(xs), _capi.ptr(Jt), _capi.ptr(cov), _capi.stream_ptr()))
        return Jt, cov

    def _predict_cov_axis(self, d, xs):
        q = np.zeros((3, xs.shape[0]))
        q[d] = xs
        Jt, cov = self.predict_cov_tensor(self._queries(list(q)))
        return Jt[0, d].cpu().numpy(), cov[0, d].cpu().numpy()

    def _predict_using_y_axis(self, d, xs, y):
        q = np.zeros((3, xs.shape[0]))
        q[d] = xs
        yy = np.zeros((3, self.M))
        yy[d] = y
        out = self.predict_using_y(list(q), list(yy))
        return out[:, d]

    def fit(self):
        raise NotImplementedError("RGP is not fitted with fit() method, use regress() instead")

    def jacobian(self, z):
        raise NotImplementedError("RGP does not have jacobian() method")


class _Spec:
    """constants of one axis before the ensemble exists"""

    def __init__(self, X, y_, C, theta):
        from .RGP import prior
        self.X, self.y_, self._C0 = np.asarray(X, dtype=np.float64), np.asarray(y_, dtype=np.float64), C
        self.theta = [float(t) for t in theta]
        self.K_x, self.K_x_inv = prior(self.X, self.theta)

    def get_theta(self):
        return list(self.theta)
