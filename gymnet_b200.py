"""Import shim: the package lives in the directory `gym.net_b200/` (the name the build contract
fixes), which Python cannot import by name.  This module turns itself into that package:
`import gymnet_b200`, `from gymnet_b200.vector import CartPoleVecEnv`, ...
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "gym.net_b200")]
__package__ = "gymnet_b200"
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
