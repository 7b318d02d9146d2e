"""Import shim that lets the reference's own numpy code (src/utils/utils.py, src/gp/RGP.py,
src/gp/GPE.py, src/quad.py) run unmodified in this container.

TEST INFRASTRUCTURE: used only by oracle/make_golden.py (fixture generation, in the build
container where /root/reference exists).  Nothing at test/bench run time imports it.
"""
import sys
import types

import numpy as np

REFERENCE_SRC = "/root/reference/src"


def install(reference_src=REFERENCE_SRC):
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    mod("casadi", MX=type("MX", (), {}), Function=type("Function", (), {}))
    mod("rospy", logwarn=lambda *a, **k: None, loginfo=lambda *a, **k: None)
    mod("matplotlib")
    mod("matplotlib.pyplot")
    mod("seaborn")
    mod("pyquaternion", Quaternion=type("Quaternion", (), {}))
    mod("config")
    mod("config.configuration_parameters", DirectoryConfig=type("DirectoryConfig", (), {}))
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    if reference_src not in sys.path:
        sys.path.insert(0, reference_src)
