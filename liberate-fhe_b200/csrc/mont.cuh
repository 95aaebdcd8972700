// mont.cuh -- 62-bit-buffer Montgomery arithmetic for sm_100a, bit-identical to the reference.
//
// The reference's mont_mult_scalar_cuda_kernel (src/liberate/ntt/ntt_cuda_kernel.cu:12-59) evaluates,
// with R = 2^62 and k = -q^-1 mod R,
//        s = (a*b*k) mod R ,   u = (a*b + s*q) / R        (an exact division, no final subtraction)
// through 31-bit halves on signed int64 (13 IMUL64).  Its result is a *mathematically unique* integer,
// so any exact evaluation reproduces it bit for bit.  We use the 64x64->128 multiplier path:
//        x   = a*b                     (lo, hi)
//        s   = (lo * k) mod 2^62
//        u   = floor(x / 2^62) + floor(s*q / 2^62) + (s != 0)
// because x mod 2^62 + (s*q) mod 2^62 is 0 or exactly 2^62 (and 0 iff s == 0, k being odd).
// floor(s*q/2^62) = umulhi(s, 4q) since q < 2^60.  ~10-11 IMAD instead of ~40.
//
// Domain of bit-exactness: |a|, |b| < 2^62 (the reference's own formula overflows beyond that).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ckks {

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be
// scheduled while its predecessor in the stream is still draining.  Every hot-path kernel starts with pdl_enter():
// "my dependents may be scheduled as soon as all my CTAs have started", then "wait until everything before me in the
// stream has completed and is visible" -- before ANY global access, so there is no RAW or WAR hazard; what is gained is the
// launch latency and the ramp-up of the next grid under the tail of this one.  No-ops for ordinary launches.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}


constexpr uint64_t MASK62 = (1ull << 62) - 1;

struct LimbConst {
    uint64_t q2;  // 2q      (the reference's _2q)
    uint64_t q4;  // 4q      (for umulhi(s, 4q) == floor(s*q / 2^62))
    uint64_t k;   // -q^-1 mod 2^62  (kl | kh << 31)
    uint64_t q;   // q       (ql | qh << 31)
};

__device__ __forceinline__ LimbConst load_limb_const(const int64_t* __restrict__ _2q, const int64_t* __restrict__ ql,
                                                     const int64_t* __restrict__ qh, const int64_t* __restrict__ kl,
                                                     const int64_t* __restrict__ kh, int i) {
    LimbConst c;
    c.q = (uint64_t)ql[i] | ((uint64_t)qh[i] << 31);
    c.q2 = _2q ? (uint64_t)_2q[i] : (c.q << 1);
    c.q4 = c.q << 2;
    c.k = (uint64_t)kl[i] | ((uint64_t)kh[i] << 31);
    return c;
}

// a: signed (|a| < 2^62) if SIGNED_A, else a in [0, 2^62).  w in [0, 2^62) (a twiddle or per-limb scalar).
template <bool SIGNED_A>
__device__ __forceinline__ int64_t mont_mul_w(int64_t a, uint64_t w, uint64_t q4, uint64_t k) {
    const uint64_t lo = (uint64_t)a * w;
    uint64_t hi = __umul64hi((uint64_t)a, w);
    if (SIGNED_A) hi -= (a < 0) ? w : 0ull;  // signed high word of a*w
    const uint64_t xh = (hi << 2) | (lo >> 62);  // floor(a*w / 2^62), two's complement
    const uint64_t s = (lo * k) & MASK62;
    const uint64_t t = __umul64hi(s, q4);
    return (int64_t)(xh + t + (s != 0 ? 1ull : 0ull));
}

// both operands signed (|a|,|b| < 2^62): the general mont_mult(a, b)
__device__ __forceinline__ int64_t mont_mul_ss(int64_t a, int64_t b, uint64_t q4, uint64_t k) {
    const uint64_t lo = (uint64_t)a * (uint64_t)b;
    const uint64_t hi = (uint64_t)__mul64hi(a, b);
    const uint64_t xh = (hi << 2) | (lo >> 62);
    const uint64_t s = (lo * k) & MASK62;
    const uint64_t t = __umul64hi(s, q4);
    return (int64_t)(xh + t + (s != 0 ? 1ull : 0ull));
}

// reference mont_redc (kern.cu:559-607): s = (x*k) mod R ; returns floor((x_low62 + s*q)/R) where the
// reference adds the full x (not x mod R) into the low partial sum: carry = (x + sl*ql) >> 31.  For
// x in [0, 2^62) that is exactly (x + s*q)/R.  For other x the reference's value is
//   floor((x + s*q) / 2^62)  evaluated with x taken as a signed 64-bit number in the low-word sum,
// which equals (x >> 62) + floor(s*q/2^62) + carry(x mod 2^62 != 0); identical to mont_mul_ss(x, 1).
__device__ __forceinline__ int64_t mont_redc(int64_t x, uint64_t q4, uint64_t k) {
    const uint64_t lo = (uint64_t)x;
    const uint64_t xh = (uint64_t)(x >> 62);
    const uint64_t s = (lo * k) & MASK62;
    const uint64_t t = __umul64hi(s, q4);
    return (int64_t)(xh + t + (s != 0 ? 1ull : 0ull));
}

// lazy add / sub mod 2q, signed compare exactly like kern.cu:1016-1058
__device__ __forceinline__ int64_t lazy_add(int64_t a, int64_t b, int64_t q2) {
    const int64_t s = a + b;
    return (s < q2) ? s : s - q2;
}
__device__ __forceinline__ int64_t lazy_sub(int64_t a, int64_t b, int64_t q2) {
    const int64_t s = a + q2 - b;
    return (s < q2) ? s : s - q2;
}
// kern.cu:664-680
__device__ __forceinline__ int64_t reduce_q(int64_t a, int64_t q) { return (a < q) ? a : a - q; }
// kern.cu:682-699
__device__ __forceinline__ int64_t make_signed(int64_t a, int64_t q) { return (a <= (q >> 1)) ? a : a - q; }

// Cooley-Tukey butterfly, kern.cu:257-274
template <bool SG>
__device__ __forceinline__ void ct_bfly(int64_t& U, int64_t& O, uint64_t w, const LimbConst& c) {
    const int64_t V = mont_mul_w<SG>(O, w, c.q4, c.k);
    const int64_t up = U + V;
    const int64_t um = U + (int64_t)c.q2 - V;
    U = (up < (int64_t)c.q2) ? up : up - (int64_t)c.q2;
    O = (um < (int64_t)c.q2) ? um : um - (int64_t)c.q2;
}

// Gentleman-Sande butterfly, kern.cu:454-472
template <bool SG>
__device__ __forceinline__ void gs_bfly(int64_t& U, int64_t& V, uint64_t w, const LimbConst& c) {
    const int64_t um = U + (int64_t)c.q2 - V;
    const int64_t O = (um < (int64_t)c.q2) ? um : um - (int64_t)c.q2;
    const int64_t up = U + V;
    V = mont_mul_w<SG>(O, w, c.q4, c.k);
    U = (up < (int64_t)c.q2) ? up : up - (int64_t)c.q2;
}

}  // namespace ckks
