"""Drop-in replacement for the reference's torch extension module `defCorrSample`
(/root/reference/offersample_LGS/setup.py:8-9, droid.cpp:138-147).

Put this directory on PYTHONPATH ahead of the reference build and
`droid_slam/modules/corr.py:8` / `droid_slam/gaussianMask_cuda.py:5` import it unchanged:
the same 7 functions, positional arguments, list-of-tensor returns and in-place side effects,
executed by the sm_100a kernels behind the C ABI in include/lgu_corr.h.
"""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
_ops = importlib.import_module("lgu-slam_b200").ops

gaussianMask = _ops.gaussianMask
gaussianMask_backward = _ops.gaussianMask_backward
lowMem_defSample = _ops.lowMem_defSample
corr_index_forward = _ops.corr_index_forward
corr_index_backward = _ops.corr_index_backward
defCorr_index_forward = _ops.defCorr_index_forward
defCorr_index_backward = _ops.defCorr_index_backward

__all__ = ["gaussianMask", "gaussianMask_backward", "lowMem_defSample", "corr_index_forward",
           "corr_index_backward", "defCorr_index_forward", "defCorr_index_backward"]
