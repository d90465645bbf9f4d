"""lgu-slam_b200 -- B200-native (sm_100a) correlation hot path of LGU-SLAM.

The directory name carries a hyphen (it mirrors the project name), so import it with
`importlib.import_module("lgu-slam_b200")` or through the root-level shim `import lgu_slam_b200`.
Contents: csrc/ (CUDA kernels + C ABI, built into liblgu_corr.so), ops (the reference's operator
API on torch CUDA tensors), dropin/defCorrSample.py (module-name-compatible replacement).
"""
from . import _lib, ops  # noqa: F401

__all__ = ["_lib", "ops"]
