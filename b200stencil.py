"""Loader: makes the package directory `finalprojectrepo.jl_b200/` (whose name is not a valid Python identifier)
importable as the module `b200stencil`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "finalprojectrepo.jl_b200")


def _load():
    name = "b200stencil"
    spec = importlib.util.spec_from_file_location(name, os.path.join(_PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_load()
