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
        if cov or return_Jt:
            raise NotImplementedError("full covariance / Jt outputs are internal to regress on the GPU path")
        mean, v = self._ens._predict_axis(self._axis, X_t_star, want_var=(var or std))
        if var:
            return mean, v
        if std:
            return mean, np.sqrt(v)
        return mean

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
