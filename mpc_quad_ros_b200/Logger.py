"""Logger — append-per-key dict logger with the reference's pickle schema (reference src/Logger.py:24-62,
src/utils/save_dataset.py:6-15), so that logs of GPU runs can be read by the reference's Visualiser / DataLoaderGP.
Keys written by the loop: x_odom, x_pred_odom, x_ref, t_odom, w_odom, t_cpu, cost_solution, rgp_basis_vectors,
rgp_mu_g_t, rgp_C_g_t, rgp_theta, v_body, a_drag (reference src/execute_trajectory.py:270-275)."""
import os
import pickle

import numpy as np


def _to_host(v):
    """CUDA tensors -> numpy (lists are converted element-wise); everything else is kept"""
    if hasattr(v, "detach"):
        return v.detach().cpu().numpy()
    if isinstance(v, (list, tuple)):
        return [_to_host(e) for e in v]
    return v


class Logger:
    def __init__(self, filename):
        """reference Logger.py:26-35: relative names land under outputs/gazebo_simulation/data/ (quirk kept)"""
        self.dictionary = dict()
        if os.path.isabs(filename):
            self.filepath = filename
        else:
            self.filepath = os.path.join(os.getcwd(), "outputs", "gazebo_simulation", "data", filename)

    def log(self, dict_to_log):
        """Logger.py:37-45: append every value to the list stored under its key"""
        for key, value in dict_to_log.items():
            self.dictionary.setdefault(key, []).append(_to_host(value))

    def save_log(self, filepath=None):
        """Logger.py:47-62 -> save_dataset.save_dict (pickle)"""
        path = filepath or self.filepath
        if not path.endswith(".pkl"):
            path = path + ".pkl"
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "wb") as f:
            pickle.dump(self.dictionary, f)
        return path


def load_log(path):
    """save_dataset.load_dict"""
    with open(path, "rb") as f:
        return pickle.load(f)


def save_rgp_state(gpe, path):
    """checkpoint of the ensemble: constants + learned mean/covariance of every vehicle.  (The reference's RGP.save
    stores only X, the prior y_ and theta, RGP.py:507-521 — kept as the 'X','y','theta' entries — and so loses the
    learned model; mu/C are added here.)"""
    d = {"X": gpe.X, "theta": gpe.theta, "y": np.zeros_like(gpe.X), "mu": gpe.mu_tensor().cpu().numpy(),
         "C": gpe.C_tensor().cpu().numpy()}
    np.savez_compressed(path, **d)


def load_rgp_state(gpe, path):
    import torch
    d = np.load(path)
    assert np.array_equal(d["X"], gpe.X) and np.allclose(d["theta"], gpe.theta), "checkpoint belongs to another ensemble"
    gpe.set_state(torch.as_tensor(d["mu"]), torch.as_tensor(d["C"]))
