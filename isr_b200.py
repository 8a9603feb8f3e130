"""Import alias: ``import isr_b200`` loads the package in ``image-super-resolution_b200/``.

The package directory is named after the reference repo (hyphenated, so not a valid
Python identifier); this shim registers it in ``sys.modules`` under ``isr_b200``.
"""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "image-super-resolution_b200")
_spec = importlib.util.spec_from_file_location(
    "isr_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["isr_b200"] = _mod
_spec.loader.exec_module(_mod)
