"""Import shim: the package lives in the directory `gym.net_b200/` (the name the build contract
fixes), which Python cannot import by name.  Importing this module loads that directory as the
package `gymnet_b200` and puts it in sys.modules in its own place:
`import gymnet_b200`, `from gymnet_b200.vector import CartPoleVecEnv`, ...
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "gym.net_b200")
_spec = _ilu.spec_from_file_location("gymnet_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_pkg = _ilu.module_from_spec(_spec)
_sys.modules["gymnet_b200"] = _pkg      # the import system returns this entry
_spec.loader.exec_module(_pkg)
