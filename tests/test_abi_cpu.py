"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/lgu_corr.h declares (no compute calls -- there is no GPU here), the operator layer refuses CPU
tensors loudly (no fallback), and the product package never touches oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lgu-slam_b200")


@pytest.fixture(scope="module")
def libpath():
    import lgu_slam_b200
    if not os.path.exists(lgu_slam_b200._lib.LIB_PATH):
        lgu_slam_b200._lib.build()
    return lgu_slam_b200._lib.LIB_PATH


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lgu_corr.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lgu_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_reference_operators():
    syms = _declared_symbols()
    for need in ("lgu_corr_index_forward", "lgu_corr_index_backward", "lgu_defcorr_index_forward",
                 "lgu_defcorr_index_backward", "lgu_gaussian_mask_forward", "lgu_gaussian_mask_backward",
                 "lgu_lowmem_defsample_forward", "lgu_altcorr_forward"):
        assert need in syms


def test_library_exports_every_declared_symbol(libpath):
    L = ctypes.CDLL(libpath)
    for s in _declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/lgu_corr.h but not exported by liblgu_corr.so"
    L.lgu_build_info.restype = ctypes.c_char_p
    assert b"sm_100a" in L.lgu_build_info()


def test_library_is_sm100a_only(libpath):
    out = subprocess.run(["cuobjdump", "--list-elf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_bad_arguments_return_error_codes(libpath):
    L = ctypes.CDLL(libpath)
    L.lgu_last_error_string.restype = ctypes.c_char_p
    rc = L.lgu_corr_index_forward(None, None, None, 1, 4, 4, 4, 4, 1, None)
    assert rc == 1 and b"null" in L.lgu_last_error_string()
    buf = ctypes.c_void_p(16)
    rc = L.lgu_defcorr_index_forward(buf, buf, buf, buf, 1, 0, 4, 4, 4, 3, None)
    assert rc == 1 and b"bad sizes" in L.lgu_last_error_string()
    assert L.lgu_corr_index_forward(buf, buf, buf, 0, 4, 4, 4, 4, 1, None) == 0      # empty edge set is a no-op


def test_ops_refuse_cpu_tensors_loudly():
    import lgu_slam_b200
    v = torch.zeros(1, 2, 2, 2, 2)
    c = torch.zeros(1, 2, 2, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        lgu_slam_b200.ops.corr_index_forward(v, c, 1)


def test_dropin_module_has_the_reference_names():
    sys.path.insert(0, os.path.join(PKG, "dropin"))
    try:
        import defCorrSample
    finally:
        sys.path.pop(0)
    # /root/reference/offersample_LGS/droid.cpp:138-147
    for name in ("gaussianMask", "gaussianMask_backward", "lowMem_defSample", "corr_index_forward",
                 "corr_index_backward", "defCorr_index_forward", "defCorr_index_backward"):
        assert callable(getattr(defCorrSample, name))


def test_product_package_never_references_the_oracle():
    bad = []
    for dp, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle|liblgu_oracle|oracle\.oracle", txt, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
