"""
build_ref.py -- compiles the REFERENCE's own CUDA extension ``ntt_cuda`` from the sources where they lie
(/root/reference/src/liberate/ntt/ntt.cpp + ntt_cuda_kernel.cu, unmodified) into oracle/_ref/ntt_cuda_ref.so
for sm_100, so that the GPU parity tests can run the reference kernels side by side with ours on the B200
(tests/test_gpu_vs_reference_kernels.py).  TEST INFRASTRUCTURE ONLY; outputs stay in oracle/_ref/ (git-ignored,
shipped to the GPU box with the snapshot).  No reference source is copied into the repository.

The only accommodation is oracle/ref_dispatch_shim.h, force-included to make the 15
AT_DISPATCH_INTEGRAL_TYPES(a.type(), ...) sites compile against torch 2.11.

    python oracle/build_ref.py            # build container only (/root/reference is absent on the GPU box)
"""
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/src/liberate/ntt")
OUT = HERE / "_ref"


def build(verbose=False):
    if not REF.exists():
        print("reference sources not present; nothing to build")
        return None
    so = OUT / "ntt_cuda_ref.so"
    if so.exists():
        return so
    OUT.mkdir(exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    shim = str(HERE / "ref_dispatch_shim.h")
    load(name="ntt_cuda_ref", sources=[str(REF / "ntt.cpp"), str(REF / "ntt_cuda_kernel.cu")],
         extra_cflags=["-O2", "-include", shim], extra_cuda_cflags=["-O3", "-include", shim],
         build_directory=str(OUT), is_python_module=False, verbose=verbose)
    return so


CSPRNG_EXTS = {"chacha20_cuda": ["chacha20.cpp", "chacha20_cuda_kernel.cu"],
               "randint_cuda": ["randint.cpp", "randint_cuda_kernel.cu"],
               "discrete_gaussian_cuda": ["discrete_gaussian.cpp", "discrete_gaussian_cuda_kernel.cu"],
               "randround_cuda": ["randround.cpp", "randround_cuda_kernel.cu"]}


def build_one(name, verbose=False):
    """one extension of the reference, from its sources in place, into oracle/_ref/"""
    if name == "ntt_cuda_ref":
        return build(verbose)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "2")
    from torch.utils.cpp_extension import load
    cs = REF.parent / "csprng"
    bdir = OUT / f"build_{name}"
    bdir.mkdir(parents=True, exist_ok=True)
    if not (bdir / f"{name}.so").exists():
        load(name=name, sources=[str(cs / s) for s in CSPRNG_EXTS[name]], extra_cuda_cflags=["-O3"],
             build_directory=str(bdir), is_python_module=False, verbose=verbose)
    return bdir / f"{name}.so"


def build_engine(verbose=False):
    """"installs" the whole reference package for the GPU-side comparison rows (bench.py --impl reference_gpu,
    tests/test_gpu_vs_reference_engine.py): the four csprng extensions are compiled UNMODIFIED from
    /root/reference/src/liberate/csprng, and the reference's Python files + prime tables are copied into
    oracle/_ref/site/liberate -- the same thing ``pip install --target`` would do.  Everything stays under the
    git-ignored oracle/_ref/; nothing of it enters the repository."""
    import shutil
    src_pkg = REF.parent
    if not src_pkg.exists():
        print("reference sources not present; nothing to install")
        return None
    site = OUT / "site"
    pkg = site / "liberate"
    if (pkg / ".complete").exists():
        return site
    # the five extensions are independent: compile them side by side, one child process each (about 4 minutes
    # instead of 11 on 8 cores)
    import subprocess
    OUT.mkdir(exist_ok=True)
    jobs = [subprocess.Popen([sys.executable, __file__, "--one", name]) for name in ["ntt_cuda_ref"] + list(CSPRNG_EXTS)]
    failed = [j.args[-1] for j in jobs if j.wait() != 0]
    if failed:
        raise RuntimeError(f"reference extensions failed to build: {failed}")
    exts = CSPRNG_EXTS
    if pkg.exists():
        shutil.rmtree(pkg)
    shutil.copytree(src_pkg, pkg, ignore=shutil.ignore_patterns("*.cu", "*.cpp", "*.h", "__pycache__", "tests"))
    for name in exts:
        shutil.copy(OUT / f"build_{name}" / f"{name}.so", pkg / "csprng" / f"{name}.so")
    shutil.copy(OUT / "ntt_cuda_ref.so", pkg / "ntt" / "ntt_cuda_ref.so")
    # liberate/ntt/__init__.py does ``from . import ntt_cuda``: forward that name to the compiled module
    (pkg / "ntt" / "ntt_cuda.py").write_text(
        "import importlib.util, pathlib, sys\n"
        "_p = pathlib.Path(__file__).with_name('ntt_cuda_ref.so')\n"
        "_s = importlib.util.spec_from_file_location('ntt_cuda_ref', _p)\n"
        "_m = importlib.util.module_from_spec(_s); _s.loader.exec_module(_m)\n"
        "globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith('_')})\n")
    (pkg / ".complete").write_text("ok")
    return site


if __name__ == "__main__":
    if "--one" in sys.argv:
        print(build_one(sys.argv[sys.argv.index("--one") + 1], verbose="-v" in sys.argv))
    elif "--engine" in sys.argv:
        print(build_engine(verbose="-v" in sys.argv))
    else:
        print(build(verbose="-v" in sys.argv))
