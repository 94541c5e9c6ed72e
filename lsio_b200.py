"""Import alias: the package directory is named `light-speed-io_b200` (not a Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("light-speed-io_b200")
sys.modules[__name__] = _pkg
