// Force-included (-include) in front of the UNMODIFIED reference sources when oracle/build_ref.py compiles
// /root/reference/src/liberate/ntt/{ntt.cpp,ntt_cuda_kernel.cu} for the parity tests.
// The shipped sources call AT_DISPATCH_INTEGRAL_TYPES(a.type(), ...) at 15 sites; torch >= 2.x no longer converts
// at::DeprecatedTypeProperties to c10::ScalarType there.  Instead of patching a copy of the
// sources, the macro is re-defined to accept either type.  Nothing else is changed.
#pragma once
#include <torch/extension.h>
#include <ATen/Dispatch.h>

namespace ckks_ref_shim {
inline c10::ScalarType st(c10::ScalarType t) { return t; }
inline c10::ScalarType st(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace ckks_ref_shim

#undef AT_DISPATCH_INTEGRAL_TYPES
#define AT_DISPATCH_INTEGRAL_TYPES(TYPE, NAME, ...) \
    AT_DISPATCH_SWITCH(ckks_ref_shim::st(TYPE), NAME, AT_DISPATCH_CASE_INTEGRAL_TYPES(__VA_ARGS__))
