"""ctypes loader for the C-ABI product library (include/lgu_corr.h -> liblgu_corr.so).

There is deliberately NO fallback: if the sm_100a library has not been built, importing the
ops fails loudly with the build command.  Nothing here imports oracle/.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LGU_CORR_LIB") or os.path.join(_HERE, "liblgu_corr.so")   # override: kernel A/B experiments
ABI_VERSION = 1

_lib = None


class LguError(RuntimeError):
    """A C-ABI call returned a non-zero status (the reference raises RuntimeError via TORCH_CHECK)."""


def build(verbose=False):
    """Compile liblgu_corr.so in-tree with nvcc for sm_100a (seconds; no GPU needed)."""
    import subprocess
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building liblgu_corr.so failed (see output above)")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (or `make -C lgu-slam_b200/csrc`). "
                "There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.lgu_abi_version.restype = ctypes.c_int
        L.lgu_build_info.restype = ctypes.c_char_p
        L.lgu_last_error_string.restype = ctypes.c_char_p
        got = L.lgu_abi_version()
        if got != ABI_VERSION:
            raise ImportError(f"liblgu_corr.so ABI version {got} != expected {ABI_VERSION}; rebuild it")
        _lib = L
    return _lib


LAUNCHES = 0          # successful C-ABI calls so far (every one launches at least one of this library's kernels)


def check(status, what):
    global LAUNCHES
    LAUNCHES += 1
    if status != 0:
        msg = lib().lgu_last_error_string().decode("utf-8", "replace")
        raise LguError(f"{what} failed (status {status}): {msg}")
