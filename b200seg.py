"""Import shim: exposes the package directory `thyroid-nodule-image-segmentation-unet-ddti_b200/` (whose name is
not a valid Python identifier) as the importable package `b200seg`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "thyroid-nodule-image-segmentation-unet-ddti_b200")
_spec = importlib.util.spec_from_file_location("b200seg", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200seg"] = _mod
_spec.loader.exec_module(_mod)
PACKAGE_DIR = _DIR
_mod.PACKAGE_DIR = _DIR
