"""Import alias for the package directory `fft-ocean-waves_b200/` (a hyphen is not importable).

    import fft_ocean_waves_b200 as fow

loads `fft-ocean-waves_b200/__init__.py` as the package `fft_ocean_waves_b200` (sub-modules included).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fft-ocean-waves_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
