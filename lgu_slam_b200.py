"""Import shim: `import lgu_slam_b200` -> the package directory `lgu-slam_b200/`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("lgu-slam_b200")
