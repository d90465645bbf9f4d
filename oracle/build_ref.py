"""TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference CUDA extension for sm_100.

The reference hot path is CUDA-only (there is no CPU implementation in
/root/reference), so the strongest oracle available is the reference's own
kernels recompiled for B200:

  oracle/_ref/defCorrSample_ref*.so   <- /root/reference/offersample_LGS/{droid.cpp,*.cu}
  oracle/_ref/altcorr_ref*.so         <- /root/reference/src/altcorr_kernel.cu + oracle/altcorr_ref_binding.cpp
  oracle/_ref/py/{corr,gaussianMask_cuda}.py  <- byte-for-byte copies of the reference's Python callers
                                         (droid_slam/modules/corr.py, droid_slam/gaussianMask_cuda.py), staged
                                         so that the GPU box -- where /root/reference does not exist -- can run
                                         the reference's OWN CorrBlock / AltCorrBlock / GaussianMask on either
                                         extension (tests/test_dropin_gpu.py)

Sources are compiled from where they lie (never copied); outputs go only into
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  The
reference's own setup.py is NOT used: its arch list stops at sm_89
(/root/reference/offersample_LGS/setup.py:21-27).  Nothing in the product
path imports these modules; tests/ and bench.py's reference-CUDA side-by-side
do, and only as the checker.

Run:  python oracle/build_ref.py          (takes ~6 min on 8 cores, cached afterwards)
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("LGU_REFERENCE_ROOT", "/root/reference")


def _built(name):
    return glob.glob(os.path.join(OUT, name + "*.so"))


def build(verbose=False, force=False):
    """Build both reference modules if /root/reference is present; return dict name->path."""
    os.makedirs(OUT, exist_ok=True)
    have_ref = os.path.isdir(os.path.join(REF, "offersample_LGS"))
    if have_ref and (force or not _built("defCorrSample_ref") or not _built("altcorr_ref")):
        os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
        os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 8))
        from torch.utils.cpp_extension import load
        if force or not _built("defCorrSample_ref"):
            d = os.path.join(OUT, "build_defcorr")
            os.makedirs(d, exist_ok=True)
            srcs = [os.path.join(REF, "offersample_LGS", f) for f in (
                "droid.cpp", "defCorrSample_kernel.cu", "lowMem_defSample.cu",
                "corrSample_kernel.cu", "gaussianAttn.cu")]
            load(name="defCorrSample_ref", sources=srcs, build_directory=d,
                 extra_cuda_cflags=["-O3"], extra_cflags=["-O3"],
                 verbose=verbose, is_python_module=False)
            _promote(d, "defCorrSample_ref")
        if force or not _built("altcorr_ref"):
            d = os.path.join(OUT, "build_altcorr")
            os.makedirs(d, exist_ok=True)
            srcs = [os.path.join(HERE, "altcorr_ref_binding.cpp"),
                    os.path.join(REF, "src", "altcorr_kernel.cu")]
            load(name="altcorr_ref", sources=srcs, build_directory=d,
                 extra_cuda_cflags=["-O3"], extra_cflags=["-O3"],
                 verbose=verbose, is_python_module=False)
            _promote(d, "altcorr_ref")
    if have_ref:
        _stage_python()
    return {n: (_built(n) or [None])[0] for n in ("defCorrSample_ref", "altcorr_ref")}


PY_FILES = {"corr.py": ("droid_slam", "modules", "corr.py"), "gaussianMask_cuda.py": ("droid_slam", "gaussianMask_cuda.py")}


def _stage_python():
    """Stage the reference's Python callers, unmodified, under oracle/_ref/py (git-ignored like the .so files)."""
    import shutil
    d = os.path.join(OUT, "py")
    os.makedirs(d, exist_ok=True)
    for name, rel in PY_FILES.items():
        src = os.path.join(REF, *rel)
        dst = os.path.join(d, name)
        if os.path.isfile(src) and (not os.path.isfile(dst) or open(src, "rb").read() != open(dst, "rb").read()):
            shutil.copyfile(src, dst)


def staged_python():
    d = os.path.join(OUT, "py")
    return d if all(os.path.isfile(os.path.join(d, n)) for n in PY_FILES) else None


def load_ref_python(def_corr_sample, droid_backends, tag):
    """Execute the staged, unmodified reference modules with `import defCorrSample` / `import droid_backends`
    resolved to the given objects (the compiled reference extension or the drop-in).  Returns (corr_module,
    gaussianMask_cuda_module) under private names, so that several bindings can live in one process; None if the
    files were not staged."""
    import importlib.util
    d = staged_python()
    if d is None:
        return None
    saved = {k: sys.modules.get(k) for k in ("defCorrSample", "droid_backends")}
    sys.modules["defCorrSample"] = def_corr_sample
    sys.modules["droid_backends"] = droid_backends
    try:
        mods = []
        for name in ("corr.py", "gaussianMask_cuda.py"):
            spec = importlib.util.spec_from_file_location(f"ref_{name[:-3]}_{tag}", os.path.join(d, name))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            mods.append(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return tuple(mods)


def _promote(build_dir, name):
    import shutil
    for so in glob.glob(os.path.join(build_dir, name + "*.so")):
        shutil.copy2(so, os.path.join(OUT, os.path.basename(so)))


def load_ref(name):
    """Import a prebuilt reference module (GPU box: prebuilt files only, no /root/reference)."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    paths = _built(name)
    if not paths:
        return None
    spec = importlib.util.spec_from_file_location(name, paths[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
