"""Import helper: the product package lives in `adjoint-ode-adaptivity_b200/` (a directory
name that is not a Python identifier); this registers it as `adjoint_ode_adaptivity_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "adjoint-ode-adaptivity_b200")
PKG_NAME = "adjoint_ode_adaptivity_b200"


def load_package():
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(PKG_NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules[PKG_NAME]
        raise
    return mod
