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


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
