"""Quadrotor3D — parameter holder for the OCP model and the simulated plant
(reference src/quad.py:22-436).  The plant integration (`update`) runs on the GPU through
qmpc_plant_period (csrc/aux_kernels.cuh); state lives in a CUDA tensor [B,13]."""
import ctypes as C
import xml.etree.ElementTree as XMLtree

import numpy as np
import torch

from . import _capi


class Quadrotor3D:
    def __init__(self, noisy=False, drag=False, payload=False, motor_noise=False, batch=1, device="cuda:0"):
        # defaults = HEAD of reference src/quad.py:41-94
        self.max_thrust = 20
        self.drag = drag
        self.max_input_value, self.min_input_value = 1, 0
        self.J = np.array([.03, .03, .06])
        self.mass = 0.03
        self.length = 0.08 / 2
        self.x_f = np.array([self.length, 0, -self.length, 0])
        self.y_f = np.array([0, self.length, 0, -self.length])
        self.c = 0.013
        self.z_l_tau = np.array([-self.c, self.c, -self.c, self.c])
        self.g = np.array([0, 0, 9.81])
        self.rotor_drag_xy, self.rotor_drag_z = 0.3, 0.0
        self.rotor_drag = np.array([self.rotor_drag_xy, self.rotor_drag_xy, self.rotor_drag_z])
        self.aero_drag = 0.008
        self.payload_mass = 0.3 * payload
        if payload or noisy or motor_noise:
            raise NotImplementedError("payload / noise options of the reference plant are not part of the hot path")
        self.batch, self.device = batch, torch.device(device)
        self._x = None          # CUDA [B,13], created on first set_state
        self.u = np.zeros(4)

    # ---- parameter sets -------------------------------------------------------------------------------
    def set_logged_pysim_params(self):
        """constants the shipped python-simulation logs were produced with (quad.py:57,60 commented originals)"""
        self.mass, self.length = 1.0, 0.47 / 2
        self.x_f = np.array([self.length, 0, -self.length, 0])
        self.y_f = np.array([0, self.length, 0, -self.length])
        return self

    def set_hummingbird_params(self):
        """config/hummingbird.xacro:29-50 through set_parameters_from_file (quad.py:385-417)"""
        return self._set_from_attrib(dict(mass=0.68, mass_rotor=0.009, ixx=0.007, iyy=0.007, izz=0.012, arm_length=0.17,
                                          max_rot_velocity=838, motor_constant=8.54858e-06, moment_constant=0.016),
                                     "hummingbird")

    def _set_from_attrib(self, a, quad_name):
        self.mass = float(a["mass"]) + float(a["mass_rotor"]) * 4
        self.J = np.array([float(a["ixx"]), float(a["iyy"]), float(a["izz"])])
        self.length = float(a["arm_length"])
        self.max_thrust = float(a["max_rot_velocity"]) ** 2 * float(a["motor_constant"])
        self.c = float(a["moment_constant"])
        if quad_name != "hummingbird":
            h = np.cos(np.pi / 4) * self.length
            self.x_f, self.y_f = np.array([h, -h, -h, h]), np.array([-h, -h, h, h])
            self.z_l_tau = np.array([-self.c, self.c, -self.c, self.c])
        else:
            self.x_f = np.array([self.length, 0, -self.length, 0])
            self.y_f = np.array([0, self.length, 0, -self.length])
            self.z_l_tau = -np.array([-self.c, self.c, -self.c, self.c])
        return self

    def set_parameters_from_file(self, params_filepath, quad_name):
        """quad.py:385-417; xacro parsing of utils.parse_xacro_file (utils.py:748-772) without getchildren()"""
        root = XMLtree.parse(params_filepath).getroot()
        a = {}
        for el in root.iter():
            name = el.attrib.get("name")
            if el.tag.endswith("property") and name in ("mass", "mass_rotor", "arm_length", "max_rot_velocity",
                                                        "motor_constant", "moment_constant"):
                a[name] = el.attrib["value"]
            if el.tag.endswith("property") and name == "body_inertia":
                for ch in el.iter():
                    if "ixx" in ch.attrib:
                        a.update(ixx=ch.attrib["ixx"], iyy=ch.attrib["iyy"], izz=ch.attrib["izz"])
        return self._set_from_attrib(a, quad_name)

    def set_cf_params(self):
        """quad.py:419-435"""
        self.mass, self.J, self.length = 0.027, np.array([1.8e-5, 1.8e-5, 3.3e-5]), 0.04
        self.max_thrust, self.c = 0.3, 0.016
        h = np.cos(np.pi / 4) * self.length
        self.x_f, self.y_f = np.array([h, -h, -h, h]), np.array([-h, -h, h, h])
        self.z_l_tau = np.array([-self.c, self.c, -self.c, self.c])
        return self

    def quad_vector(self):
        """flat [20] parameter block of include/qmpc.h"""
        return np.concatenate([[self.mass, self.max_thrust], self.J, self.x_f, self.y_f, self.z_l_tau, self.g]).astype(np.float64)

    def plant_vector(self):
        if not self.drag:
            return np.zeros(4)
        return np.array([self.aero_drag, *self.rotor_drag], dtype=np.float64)

    # ---- state ------------------------------------------------------------------------------------------
    def set_state(self, x):
        """x: (13,) numpy (batch 1) or [B,13] tensor/array"""
        t = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=torch.float64)
        t = t.reshape(-1, 13)
        assert t.shape[0] == self.batch, f"expected {self.batch} states"
        self._x = t.to(self.device).contiguous().clone()

    def state_tensor(self):
        return self._x

    def get_state(self, quaternion=False, stacked=False, body_frame=False):
        if not (quaternion and stacked):
            raise NotImplementedError("only get_state(quaternion=True, stacked=True[, body_frame]) is on the hot path")
        x = self._x
        if body_frame:
            from .utils.utils import body_velocity
            x = x.clone()
            x[:, 7:10] = body_velocity(x)
        return x[0].cpu().numpy() if self.batch == 1 else x

    def get_control(self):
        return self.u

    def update(self, u, dt, n_sub=1):
        """Quadrotor3D.update (quad.py:234-253): clip u to [0,1], one RK4 step of the plant (n_sub steps of dt)."""
        ut = torch.as_tensor(np.asarray(u) if not torch.is_tensor(u) else u, dtype=torch.float64).reshape(-1, 4)
        ut = ut.to(self.device).contiguous()
        assert ut.shape[0] == self.batch
        q, p = self.quad_vector(), self.plant_vector()
        _capi.check(_capi.lib().qmpc_plant_period(
            q.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), self.batch, _capi.ptr(self._x), _capi.ptr(ut),
            C.c_double(dt), int(n_sub), _capi.stream_ptr()))
        self.u = np.clip(np.asarray(u), 0, 1) if not torch.is_tensor(u) else u.clamp(0, 1)
