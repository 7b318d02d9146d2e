"""RGP — recursive Gaussian process on fixed basis points (reference src/gp/RGP.py:104-330), one 1-D model per
body axis.  Device state lives in a GPEnsemble (csrc/rgp_kernels.cuh through the C-ABI); an RGP object is a view
of one axis of an ensemble, with the reference's attribute names (X, mu_g_t, C_g_t, K_x, K_x_inv, sigma_n)."""
import numpy as np


def rbf_matrix(x1, x2, L, sigma_f):
    """RBF.calculate_covariance_matrix (RGP.py:88-99), vectorised: sigma_f^2 exp(-1/2 (a-b) L^-2 (a-b))"""
    e = np.asarray(x1, dtype=np.float64)[:, None] - np.asarray(x2, dtype=np.float64)[None, :]
    return sigma_f ** 2 * np.exp(-0.5 * e * (1.0 / (L * L)) * e)


def prior(X, theta):
    """K_x = K(X,X) + sigma_n^2 I and its inverse, exactly as RGP.__init__ does (RGP.py:156-157)"""
    L, sf, sn = theta
    Kx = rbf_matrix(X, X, L, sf) + sn ** 2 * np.eye(len(X))
    return Kx, np.linalg.inv(Kx)


class RBF:
    """RGP.py:24-102 (numpy branch)"""

    def __init__(self, L=np.eye(1), sigma_f=1):
        self.L, self.sigma_f = L, sigma_f

    def __call__(self, x1, x2):
        L = float(np.asarray(self.L).ravel()[0])
        return float(self.sigma_f ** 2 * np.exp(-0.5 * (x1 - x2) * (1.0 / (L * L)) * (x1 - x2)))

    def calculate_covariance_matrix(self, x1, x2):
        assert x1.ndim == 1 and x2.ndim == 1
        return rbf_matrix(x1, x2, float(np.asarray(self.L).ravel()[0]), self.sigma_f)


class RGP:
    def __init__(self, X, y_, C=None, theta=[1.0, 0.1, 0.1], _ensemble=None, _axis=0):
        assert X.ndim == 1, "X must be a 1D array"
        assert y_.ndim == 1, "y_ must be a 1D array"
        assert X.shape[0] == y_.shape[0], "X and y_ must have the same number of rows"
        assert len(theta) == 3, "theta must be a list of 3 hyperparameters [L, sigma_f, sigma_n]"
        if C is not None:
            assert C.shape[0] == C.shape[1] and C.shape[0] == X.shape[0]
        self.X, self.y_ = X, y_
        self.theta = [float(t) for t in theta]
        self.sigma_n = self.theta[2]
        self.K = RBF(L=np.eye(1) * self.theta[0], sigma_f=self.theta[1])
        self.K_x, self.K_x_inv = prior(X, self.theta)
        self._C0 = C
        self._ens, self._axis = _ensemble, _axis
        if _ensemble is None:          # stand-alone RGP: a private ensemble whose axis 0 is this model
            from .GPE import GPEnsemble
            self._ens = GPEnsemble([self, _Clone(self), _Clone(self)], "RGP")
            self._axis = 0

    def get_theta(self):
        return list(self.theta)

    # device state views (batch 1 -> numpy like the reference; batched -> tensors [B,M], [B,M,M])
    @property
    def mu_g_t(self):
        return self._ens._mu_axis(self._axis)

    @property
    def C_g_t(self):
        return self._ens._C_axis(self._axis)

    def regress(self, Xt, yt):
        """RGP.py:303-330 (k = 1 sample)"""
        assert Xt.ndim == 1 and yt.ndim == 1 and Xt.shape == yt.shape
        assert Xt.shape[0] == 1, "the control loop regresses one sample per call (utils.compute_a_drag)"
        self._ens._regress_axis(self._axis, Xt, yt)
        return self.mu_g_t, self.C_g_t

    def predict(self, X_t_star, cov=False, var=False, std=False, return_Jt=False):
        """RGP.py:168-229 numpy branch: posterior mean (and var / std) at the query points"""
        assert isinstance(X_t_star, np.ndarray) and X_t_star.ndim == 1
        if cov or return_Jt:           # same precedence and return shapes as the reference (:212-229)
            mean, _ = self._ens._predict_axis(self._axis, X_t_star, want_var=False)
            Jt, C_p = self._ens._predict_cov_axis(self._axis, X_t_star)
            out = (mean, C_p) if cov else ((mean, np.diag(C_p)) if var else ((mean, np.sqrt(np.diag(C_p))) if std else (mean,)))
            if return_Jt:
                out = out + (Jt,)
            return out if len(out) > 1 else out[0]
        mean, v = self._ens._predict_axis(self._axis, X_t_star, want_var=(var or std))
        if var:
            return mean, v
        if std:
            return mean, np.sqrt(v)
        return mean

    def learn(self, Xt, yt):
        """RGP.py:332-482.  Hyper-parameter learning changes K_x / K_x_inv of this one model, which the control loop's
        ensemble shares between vehicles; it lives on its own device object (never called from the loop, as in the
        reference): use RGPLearner(X, y_, C, theta).learn(Xt, yt)."""
        raise NotImplementedError("use mpc_quad_ros_b200.gp.RGP.RGPLearner for RGP.learn (RGP*)")

    def predict_using_y(self, X_t_star, y, cov=False, var=False, std=False, return_Jt=False):
        """RGP.py:235-300 numpy branch (mean only)"""
        assert isinstance(X_t_star, np.ndarray) and isinstance(y, np.ndarray)
        if cov or var or std or return_Jt:
            raise NotImplementedError("predict_using_y returns the mean on the GPU path")
        return self._ens._predict_using_y_axis(self._axis, X_t_star, y)


class _Clone:
    """same constants as another RGP (placeholder axes of a stand-alone model)"""

    def __init__(self, other):
        self.X, self.y_, self.theta, self._C0 = other.X, other.y_, other.theta, other._C0
        self.K_x, self.K_x_inv, self.sigma_n = other.K_x, other.K_x_inv, other.sigma_n

    def get_theta(self):
        return list(self.theta)


class RGPLearner:
    """RGP* — RGP.__init__ + RGP.learn of the reference (src/gp/RGP.py:106-157, 332-482, sigma points :485-505) for a batch
    of independent 1-D models that share one basis grid X: joint recursive estimate of the function values at X and of
    the hyper-parameters eta = (L, sigma_f, sigma_n) by the reference's unscented transform.  batch == 1 speaks numpy
    like the reference ((1,) arrays in, (mu_z, C_z) out); batch > 1 takes / returns CUDA tensors with a leading model
    dimension.  Attribute names follow the reference: mu_g_t, C_g_t, mu_eta_t, C_eta_t, K_x_inv, get_theta()."""

    def __init__(self, X, y_=None, C=None, theta=[1.0, 0.1, 0.1], batch=1, device=None):
        import ctypes as Ct
        import torch
        from .. import _capi
        X = np.ascontiguousarray(X, dtype=np.float64)
        assert X.ndim == 1, "X must be a 1D array"
        assert len(theta) == 3, "theta must be a list of 3 hyperparameters [L, sigma_f, sigma_n]"
        self.X, self.batch, self.M = X, int(batch), X.shape[0]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._capi, self._torch = _capi, torch
        self._h = Ct.c_void_p()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        p = lambda a: a.ctypes.data_as(Ct.c_void_p)
        _capi.check(_capi.lib().qrgpl_create(self.batch, self.M, p(X), p(th), self.device.index or 0, Ct.byref(self._h)))
        if (y_ is not None and np.any(np.asarray(y_) != 0)) or C is not None:
            mu = None if y_ is None else torch.as_tensor(np.broadcast_to(np.asarray(y_, dtype=np.float64), (self.batch, self.M)).copy(), device=self.device)
            Cm = None if C is None else torch.as_tensor(np.broadcast_to(np.asarray(C, dtype=np.float64), (self.batch, self.M, self.M)).copy(), device=self.device)
            _capi.check(_capi.lib().qrgpl_set_state(self._h, _capi.ptr(mu), _capi.ptr(Cm), None, None, None, None, _capi.stream_ptr()))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._capi.lib().qrgpl_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _state(self, which):
        torch, B, M = self._torch, self.batch, self.M
        shapes = {"mu_g": (B, M), "C_g": (B, M, M), "mu_eta": (B, 3), "C_eta": (B, 3, 3), "Kx_inv": (B, M, M)}
        t = torch.empty(shapes[which], dtype=torch.float64, device=self.device)
        args = [self._capi.ptr(t) if k == which else None for k in ("mu_g", "C_g", "mu_eta", "C_eta", "Kx_inv")]
        self._capi.check(self._capi.lib().qrgpl_get_state(self._h, *args, self._capi.stream_ptr()))
        return t[0].cpu().numpy() if B == 1 else t

    mu_g_t = property(lambda self: self._state("mu_g"))
    C_g_t = property(lambda self: self._state("C_g"))
    mu_eta_t = property(lambda self: self._state("mu_eta"))
    C_eta_t = property(lambda self: self._state("C_eta"))
    K_x_inv = property(lambda self: self._state("Kx_inv"))

    def get_theta(self):
        eta = self.mu_eta_t
        return [float(v) for v in eta] if self.batch == 1 else eta

    def status(self):
        st = self._torch.empty(self.batch, dtype=self._torch.int32, device=self.device)
        self._capi.check(self._capi.lib().qrgpl_get_status(self._h, self._capi.ptr(st), self._capi.stream_ptr()))
        return st

    def learn(self, Xt, yt):
        """RGP.learn: one sample per model; returns (mu_z [M+3], C_z [M+3, M+3]) (leading model dimension if batched)"""
        torch, B, M = self._torch, self.batch, self.M
        if B == 1 and isinstance(Xt, np.ndarray):
            assert Xt.shape[0] == 1 and yt.shape[0] == 1, "Only one-dimensional regression is supported"
        xt = torch.as_tensor(np.asarray(Xt, dtype=np.float64).reshape(B) if not torch.is_tensor(Xt) else Xt.reshape(B), device=self.device).contiguous()
        y = torch.as_tensor(np.asarray(yt, dtype=np.float64).reshape(B) if not torch.is_tensor(yt) else yt.reshape(B), device=self.device).contiguous()
        mu_z = torch.empty((B, M + 3), dtype=torch.float64, device=self.device)
        C_z = torch.empty((B, M + 3, M + 3), dtype=torch.float64, device=self.device)
        self._capi.check(self._capi.lib().qrgpl_learn(self._h, self._capi.ptr(xt), self._capi.ptr(y), self._capi.ptr(mu_z),
                                                       self._capi.ptr(C_z), self._capi.stream_ptr()))
        return (mu_z[0].cpu().numpy(), C_z[0].cpu().numpy()) if B == 1 else (mu_z, C_z)
