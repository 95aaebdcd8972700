// ntt_fast.cuh -- canonical-output ("fast") batched NTT / iNTT for the fused hot path, sm_100a (B200).
//
// Same two-pass / radix-16-in-registers structure and index algebra as ntt_kernels.cuh, but the butterflies no
// longer reproduce the reference's lazy representatives: outputs are CANONICAL ([0,q), optionally centred),
// which is all the fused mult / rotate path needs (every reference sequence on that path ends in reduce_2q;
// DESIGN.md section 6).  That freedom is spent on the instruction mix the B200 actually has
// (profiles/r01_pipe_microbench.txt: 4 issue slots/clk/SM; IMAD 2/clk, FP64 1.8/clk, IMAD+DFMA co-issue 3.8/clk):
//
//   * scale primes (q < 2^42, 34 of 39 limbs at gold): FP64 ERROR-FREE butterflies.  A coefficient (|v| < 2^51) lives
//     in a double; v*w is formed exactly as p + e = (v*w rounded, fma residual), the quotient by a magic-constant
//     rint(p/q), and r = fma(-c, q, p) + e is the exact integer v*w - c*q with |r| < 0.54 q.  8 FP64-pipe
//     instructions per butterfly, no carries, no conditional corrections, on a pipe the integer kernels leave idle.
//   * 60-bit primes (base + special): Shoup/Harvey lazy butterflies on 64-bit integers, twiddle w with
//     w' = floor(w 2^64 / q): t = umulhi(v, w'), r = v*w - t*q in [0, 2q); one conditional correction per butterfly.
//
// Twiddles are PLAIN (non-Montgomery) powers of psi: the transform is linear, so Montgomery-form data stay in
// Montgomery form.  Tables: shoup[C][N] = {w, w'} (16 B) and dbl[C][N] (8 B, small primes only), plus PACKED copies of
// the last four stages (one thread's 15 twiddles stored lane-interleaved, see TwPacked) so that the block passes read
// them with fully coalesced 128-bit loads.
#pragma once
#include "ntt_kernels.cuh"

namespace ckks {

constexpr double F64_MAGIC = 6755399441055744.0;      // 1.5 * 2^52: x + MAGIC - MAGIC == rint(x) for |x| < 2^51
constexpr uint64_t SMALL_PRIME_LIMIT = 1ull << 42;

// ------------------------------------------------------------------------------------------------------------
// FP64 arithmetic
// ------------------------------------------------------------------------------------------------------------
struct F64C {
    double q, qinv;
};
__device__ __forceinline__ double f64_mulmod(double v, double w, const F64C& c) {
    const double p = __dmul_rn(v, w);
    const double e = __fma_rn(v, w, -p);
    const double m = __dadd_rn(__fma_rn(p, c.qinv, F64_MAGIC), -F64_MAGIC);
    return __dadd_rn(__fma_rn(-m, c.q, p), e);
}
__device__ __forceinline__ double f64_reduce(double v, const F64C& c) {   // -> |r| <= q/2 (+1)
    const double m = __dadd_rn(__fma_rn(v, c.qinv, F64_MAGIC), -F64_MAGIC);
    return __fma_rn(-m, c.q, v);
}
// |x| < 2^51 integer <-> double without the slow conversion pipe
__device__ __forceinline__ double i2d(int64_t x) {
    const uint64_t t = (uint64_t)(x + (1ll << 51)) | 0x4330000000000000ull;
    return __dadd_rn(__longlong_as_double((long long)t), -(4503599627370496.0 + 2251799813685248.0));
}
__device__ __forceinline__ int64_t d2i(double v) {
    const uint64_t t = (uint64_t)__double_as_longlong(__dadd_rn(v, 4503599627370496.0 + 2251799813685248.0));
    return (int64_t)(t & 0x000FFFFFFFFFFFFFull) - (1ll << 51);
}

struct ArithF64 {
    using T = double;
    using TW = double;
    using C = F64C;
    static __device__ __forceinline__ T load(int64_t x) { return i2d(x); }
    static __device__ __forceinline__ int64_t store_lazy(T v, const C& c) { return d2i(f64_reduce(v, c)); }
    // hand-off between the two passes of one transform: the raw double (|v| < 2^51 integer-valued), no conversion
    static __device__ __forceinline__ int64_t store_mid(T v, const C& c) { return (int64_t)__double_as_longlong(v); }
    static __device__ __forceinline__ T load_mid(int64_t x) { return __longlong_as_double((long long)x); }
    static __device__ __forceinline__ int64_t store_canon(T v, const C& c, bool centred) {
        double r = f64_reduce(v, c);
        if (!centred) r = (r < 0.0) ? __dadd_rn(r, c.q) : r;
        return d2i(r);
    }
    // hand-off with the neighbouring fused kernels: raw != 0 means the buffer holds doubles (integer-valued, |v| < 2^51)
    static __device__ __forceinline__ T load_in(int64_t x, int raw) { return raw ? load_mid(x) : load(x); }
    static __device__ __forceinline__ int64_t store_out(T v, const C& c, int raw) { return raw ? store_mid(v, c) : store_canon(v, c, false); }
    static __device__ __forceinline__ T mul(T v, TW w, const C& c) { return f64_mulmod(v, w, c); }
    static __device__ __forceinline__ void ct(T& U, T& V, TW w, const C& c) {
        const T r = f64_mulmod(V, w, c);
        V = __dadd_rn(U, -r);
        U = __dadd_rn(U, r);
    }
    static __device__ __forceinline__ void gs(T& U, T& V, TW w, const C& c) {
        const T t = __dadd_rn(U, -V);
        U = __dadd_rn(U, V);
        V = f64_mulmod(t, w, c);
    }
    // magnitudes double along the sum outputs of GS stages: re-centre once per radix-16 round
    static __device__ __forceinline__ void tame(T& v, const C& c) { v = f64_reduce(v, c); }
    // one whole radix-2 stage (8 butterflies) written "vertically": every step of the 6-deep mulmod chain is issued
    // for all 8 butterflies before the next step, so that 8 independent FP64 chains are in flight per warp
    template <int I>
    static __device__ __forceinline__ void ct_stage(T (&e)[16], const TW (&w)[8], const C& c) {
        constexpr int d = 8 >> I;
        double p[8], r[8], m[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k = ((b / d) * 2 * d) + (b % d);
            p[b] = __dmul_rn(e[k + d], w[k >> (4 - I)]);
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k = ((b / d) * 2 * d) + (b % d);
            r[b] = __fma_rn(e[k + d], w[k >> (4 - I)], -p[b]);
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) m[b] = __fma_rn(p[b], c.qinv, F64_MAGIC);
#pragma unroll
        for (int b = 0; b < 8; ++b) m[b] = __dadd_rn(m[b], -F64_MAGIC);
#pragma unroll
        for (int b = 0; b < 8; ++b) p[b] = __fma_rn(-m[b], c.q, p[b]);
#pragma unroll
        for (int b = 0; b < 8; ++b) p[b] = __dadd_rn(p[b], r[b]);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k = ((b / d) * 2 * d) + (b % d);
            const double u = e[k];
            e[k] = __dadd_rn(u, p[b]);
            e[k + d] = __dadd_rn(u, -p[b]);
        }
    }
    template <int I>   // I = distance exponent: pairs (k, k + 2^I), twiddle index k >> (I+1)
    static __device__ __forceinline__ void gs_stage(T (&e)[16], const TW (&w)[8], const C& c) {
        constexpr int d = 1 << I;
        double t[8], p[8], r[8], m[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k = ((b / d) * 2 * d) + (b % d);
            t[b] = __dadd_rn(e[k], -e[k + d]);
            e[k] = __dadd_rn(e[k], e[k + d]);
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k = ((b / d) * 2 * d) + (b % d);
            p[b] = __dmul_rn(t[b], w[k >> (I + 1)]);
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k = ((b / d) * 2 * d) + (b % d);
            r[b] = __fma_rn(t[b], w[k >> (I + 1)], -p[b]);
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) m[b] = __fma_rn(p[b], c.qinv, F64_MAGIC);
#pragma unroll
        for (int b = 0; b < 8; ++b) m[b] = __dadd_rn(m[b], -F64_MAGIC);
#pragma unroll
        for (int b = 0; b < 8; ++b) p[b] = __fma_rn(-m[b], c.q, p[b]);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k = ((b / d) * 2 * d) + (b % d);
            e[k + d] = __dadd_rn(p[b], r[b]);
        }
    }
    template <int RUN, bool SMEM>
    static __device__ __forceinline__ void load_tw(TW (&w)[8], const TW* __restrict__ p) {
        if (RUN == 1) {
            w[0] = SMEM ? *p : __ldg(p);
        } else {
#pragma unroll
            for (int g = 0; g < RUN; g += 2) {
                const double2 v = SMEM ? *reinterpret_cast<const double2*>(p + g) : __ldg(reinterpret_cast<const double2*>(p + g));
                w[g] = v.x;
                w[g + 1] = v.y;
            }
        }
    }
    template <int I>   // stage I of the packed last group (see TwPacked)
    static __device__ __forceinline__ void load_tw_packed(TW (&w)[8], const TW* __restrict__ wt, int lane) {
        if constexpr (I == 0) {
            w[0] = __ldg(wt + lane);
        } else {
            const double2* __restrict__ gr = reinterpret_cast<const double2*>(wt + 64) + (((1 << (I - 1)) - 1) * 32 + lane);
#pragma unroll
            for (int g = 0; g < (1 << (I - 1)); ++g) {
                const double2 v = __ldg(gr + g * 32);
                w[2 * g] = v.x;
                w[2 * g + 1] = v.y;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------------------
// 64-bit integer Shoup / Harvey arithmetic (any q < 2^62; values kept in [0, 4q) forward, [0, 2q) inverse)
// ------------------------------------------------------------------------------------------------------------
struct U64C {
    uint64_t q, q2;
};
__device__ __forceinline__ uint64_t shoup_mul(uint64_t v, uint64_t w, uint64_t wp, uint64_t q) {
    const uint64_t t = __umul64hi(v, wp);
    return v * w - t * q;   // in [0, 2q) for any v < 2^64
}
struct ArithU64 {
    using T = uint64_t;
    using TW = ulonglong2;   // {w, w'}
    using C = U64C;
    static __device__ __forceinline__ T load(int64_t x) { return (uint64_t)x; }   // expects [0, 4q)
    static __device__ __forceinline__ int64_t store_lazy(T v, const C& c) { return (int64_t)v; }
    static __device__ __forceinline__ int64_t store_mid(T v, const C& c) { return (int64_t)v; }
    static __device__ __forceinline__ T load_mid(int64_t x) { return (uint64_t)x; }
    static __device__ __forceinline__ int64_t store_canon(T v, const C& c, bool centred) {
        v = (v >= c.q2) ? v - c.q2 : v;
        v = (v >= c.q2) ? v - c.q2 : v;           // [0,4q) or even [0,6q) -> [0,2q)
        v = (v >= c.q) ? v - c.q : v;
        int64_t r = (int64_t)v;
        if (centred) r = (r > (int64_t)(c.q >> 1)) ? r - (int64_t)c.q : r;
        return r;
    }
    // integer rows never use the raw-double hand-off: their buffers always hold canonical / lazy integers
    static __device__ __forceinline__ T load_in(int64_t x, int raw) { return load(x); }
    static __device__ __forceinline__ int64_t store_out(T v, const C& c, int raw) { return store_canon(v, c, false); }
    static __device__ __forceinline__ T mul(T v, TW w, const C& c) { return shoup_mul(v, w.x, w.y, c.q); }
    static __device__ __forceinline__ void ct(T& U, T& V, TW w, const C& c) {
        const T u = (U >= c.q2) ? U - c.q2 : U;
        const T t = shoup_mul(V, w.x, w.y, c.q);
        U = u + t;
        V = u - t + c.q2;
    }
    static __device__ __forceinline__ void gs(T& U, T& V, TW w, const C& c) {
        const T s = U + V;
        const T d = U - V + c.q2;
        U = (s >= c.q2) ? s - c.q2 : s;
        V = shoup_mul(d, w.x, w.y, c.q);
    }
    static __device__ __forceinline__ void tame(T& v, const C& c) {}
    template <int I>
    static __device__ __forceinline__ void ct_stage(T (&e)[16], const TW (&w)[8], const C& c) {
        constexpr int d = 8 >> I;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (!(k & d)) ct(e[k], e[k + d], w[k >> (4 - I)], c);
    }
    template <int I>
    static __device__ __forceinline__ void gs_stage(T (&e)[16], const TW (&w)[8], const C& c) {
        constexpr int d = 1 << I;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (!(k & d)) gs(e[k], e[k + d], w[k >> (I + 1)], c);
    }
    template <int RUN, bool SMEM>
    static __device__ __forceinline__ void load_tw(TW (&w)[8], const TW* __restrict__ p) {
#pragma unroll
        for (int g = 0; g < RUN; ++g) w[g] = SMEM ? p[g] : __ldg(p + g);
    }
    template <int I>
    static __device__ __forceinline__ void load_tw_packed(TW (&w)[8], const TW* __restrict__ wt, int lane) {
        const TW* __restrict__ en = wt + ((1 << I) - 1) * 32 + lane;
#pragma unroll
        for (int g = 0; g < (1 << I); ++g) w[g] = __ldg(en + g * 32);
    }
};

// ------------------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk) staging of a CTA's twiddles into shared memory, completion on an mbarrier
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// L2 prefetch (TMA prefetch engine / LSU prefetch): pulls the tile that a CTA about one wave later will load from
// HBM into L2 while this CTA computes, turning that CTA's exposed DRAM latency into an L2 hit.
constexpr int PREFETCH_ROWS_AHEAD = 28;   // x 16 chunks ~ one wave of 3 CTAs x 148 SMs
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
// pulls one line into L1 without keeping a register (the destination is never read)
__device__ __forceinline__ void l1_touch(const void* gptr) {
    unsigned dummy;
    asm volatile("ld.global.nc.L1::evict_last.b32 %0, [%1];" : "=r"(dummy) : "l"(gptr));
}
__device__ __forceinline__ void l2_prefetch_line(const void* gptr) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(gptr));
}
// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): a thread's 16 contiguous coefficients move as 4 full
// 32-byte sectors, so contiguous register tiles go to / come from global memory without a shared-memory detour
__device__ __forceinline__ void ldg256(const int64_t* p, int64_t (&r)[16], int j) {
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r[4 * j]), "=l"(r[4 * j + 1]), "=l"(r[4 * j + 2]), "=l"(r[4 * j + 3])
                 : "l"(p + 4 * j));
}
__device__ __forceinline__ void stg256(int64_t* p, const int64_t (&r)[16], int j) {
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * j), "l"(r[4 * j]), "l"(r[4 * j + 1]),
                 "l"(r[4 * j + 2]), "l"(r[4 * j + 3])
                 : "memory");
}
// where a round's twiddles come from: the global table (index 2^s + ...) or the CTA's staged copy in shared memory
template <class TW>
struct TwGlobal {
    const TW* W;
    int s0;
    unsigned pre;
    static constexpr bool SMEM = false;
    __device__ __forceinline__ const TW* run(int i) const { return W + ((1u << (s0 + i)) + (pre << i)); }
};
// The last four stages of a forward block pass (the first four levels of an inverse one) use 15 twiddles that belong to
// exactly one thread (the thread that holds 16 contiguous coefficients).  In the plain table they are 1 + 2 + 4 + 8
// values at four different places, 8 to 64 bytes per thread: a warp's 128-bit loads touch up to 16 lines each.  The
// PACKED table stores them per warp tile (32 threads = 512 coefficients) lane-interleaved:
//   doubles  [512 per warp tile]: s0[lane] (32) | pad (32) | granule g = 0..6: {w, w}[lane] -- g 0: stage 1, 1-2: stage 2, 3-6: stage 3
//   {w, w'}  [512 per warp tile]: entry e = 0..14 (stage i, index j -> e = 2^i - 1 + j): [e][lane]
// so every load of a warp is one contiguous 256- or 512-byte run.
constexpr int PACK_TILE = 512;
template <class TW>
struct TwPacked {
    const TW* wt;   // this warp tile's packed twiddles
    int lane;
    static constexpr bool SMEM = false;
};
template <class SRC>
struct is_packed { static constexpr bool value = false; };
template <class TW>
struct is_packed<TwPacked<TW>> { static constexpr bool value = true; };

// ------------------------------------------------------------------------------------------------------------
// rounds
// ------------------------------------------------------------------------------------------------------------
template <class A, int I, class SRC>
__device__ __forceinline__ void fast_fwd_stage(typename A::T (&e)[16], const SRC& src, const typename A::C& c) {
    typename A::TW w[8];
    if constexpr (is_packed<SRC>::value)
        A::template load_tw_packed<I>(w, src.wt, src.lane);
    else
        A::template load_tw<(1 << I), SRC::SMEM>(w, src.run(I));
    A::template ct_stage<I>(e, w, c);
}
template <class A, int FIRST, class SRC>
__device__ __forceinline__ void fast_fwd_round(typename A::T (&e)[16], const SRC& src, const typename A::C& c) {
    if constexpr (FIRST <= 0) fast_fwd_stage<A, 0>(e, src, c);
    if constexpr (FIRST <= 1) fast_fwd_stage<A, 1>(e, src, c);
    if constexpr (FIRST <= 2) fast_fwd_stage<A, 2>(e, src, c);
    fast_fwd_stage<A, 3>(e, src, c);
}
template <class A, int I, class SRC>
__device__ __forceinline__ void fast_inv_stage(typename A::T (&e)[16], const SRC& src, const typename A::C& c) {
    typename A::TW w[8];
    if constexpr (is_packed<SRC>::value)
        A::template load_tw_packed<3 - I>(w, src.wt, src.lane);
    else
        A::template load_tw<(1 << (3 - I)), SRC::SMEM>(w, src.run(3 - I));
    A::template gs_stage<I>(e, w, c);
}
template <class A, int NST, class SRC>
__device__ __forceinline__ void fast_inv_round(typename A::T (&e)[16], const SRC& src, const typename A::C& c) {
    fast_inv_stage<A, 0>(e, src, c);
    if constexpr (NST >= 2) fast_inv_stage<A, 1>(e, src, c);
    if constexpr (NST >= 3) fast_inv_stage<A, 2>(e, src, c);
    if constexpr (NST >= 4) fast_inv_stage<A, 3>(e, src, c);
#pragma unroll
    for (int k = 0; k < 16; ++k) A::tame(e[k], c);
}

// shared-memory exchange on the raw 64-bit patterns (double and uint64 share the slot layout)
template <class T>
__device__ __forceinline__ void smx_store(int64_t* sm, const T (&e)[16], int tau, int p) {
    int64_t r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = *reinterpret_cast<const int64_t*>(&e[k]);
    sm_store_field(sm, r, tau, p);
}
template <class T>
__device__ __forceinline__ void smx_load(const int64_t* sm, T (&e)[16], int tau, int p) {
    int64_t r[16];
    sm_load_field(sm, r, tau, p);
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = *reinterpret_cast<const T*>(&r[k]);
}

struct FastArgs {
    int64_t* a;                 // [rows] in place
    long long a_stride;
    const ulonglong2* tw_u64;   // [period][N] {w, w'} plain twiddles (psi for forward, psi^-1 for inverse)
    const double* tw_f64;       // [period][N] the same values as doubles (only read for small primes)
    const int64_t* q;           // [period]
    const int64_t* scal;        // [period] optional per-limb plain multiplier (forward: on load; inverse: at the end)
    const uint64_t* scal_sh;    // [period] its Shoup companion floor(s 2^64 / q)
    int period;                 // constants / twiddles of row r are those of limb r % period
    int logN;
    int centred;                // inverse only: output in (-q/2, q/2] instead of [0, q)
    int force_int;              // 1: integer path for every limb
    // slab view of a [groups][group_rows] block: grid row r -> group r / slab_rows, member r % slab_rows;
    // data row = group * group_rows + slab_t0 + member, limb = slab_t0 + member.  slab_rows == 0: plain rows.
    int slab_rows, group_rows, slab_t0;
    int prefetch;               // rows ahead whose tile is pulled into L2 by every CTA (0 = off)
    int row0;                   // plain rows: limb of grid row r is (row0 + r) % period (row slabs of one batch)
    int in_raw, out_raw;        // FP64 rows: the input / the forward output are raw doubles (fused key switch), not int64
    // PACKED last-group twiddles (TwPacked), or nullptr: the block passes then read the plain tables
    const ulonglong2* twp_u64;  // [period][N]
    const double* twp_f64;      // [period][N]
    const double* qinv;         // [period] 1/q as doubles, or nullptr (computed per CTA: an FP64 division)
    // NTT-domain data in WARP-INTERLEAVED order (fused path only; pointwise consumers do not care, the evaluation key is
    // permuted once): inside every 512-coefficient warp tile, coefficient 16 t + k sits at ((k >> 1) * 32 + t) * 2 + (k & 1),
    // so the block passes move a thread's 16 contiguous coefficients as eight lane-contiguous 128-bit accesses
    // (512 contiguous bytes per warp instruction) instead of four 256-bit accesses 128 bytes apart (32 lines per instruction).
    int perm;
    // lab builds (-DCKKS_LAB) only: bit 0 = skip the butterfly rounds, bit 1 = skip the global loads / stores of the
    // forward passes -- separates "FP64 work" from "memory + overhead" when timing a pass (scripts/ntt_lab.py)
    int lab;
};
#ifdef CKKS_LAB
#define LAB_SKIP_MATH(F) ((F).lab & 1)
#define LAB_SKIP_MEM(F) ((F).lab & 2)
#else
#define LAB_SKIP_MATH(F) false
#define LAB_SKIP_MEM(F) false
#endif

// rows are dispatched LAST ROW FIRST: the 60-bit limbs (base + special primes, the last rows of every period / slab group) take
// ~2.3x the instructions of an FP64 row, so they should not be what the final wave of a launch consists of
__device__ __forceinline__ int grid_row(const FastArgs& F) { return (int)gridDim.y - 1 - (int)blockIdx.y; }
__device__ __forceinline__ unsigned grid_chunk(const FastArgs& F) { return blockIdx.x; }

struct RowId {
    long long data_row;
    int limb;
};
// the tile that a CTA dispatched about `ahead` rows x (chunks per row) tiles later will load: data row (or -1) and chunk
__device__ __forceinline__ long long fast_row_ahead(const FastArgs& F, int ahead, unsigned& chunk) {
    const int r = grid_row(F) - ahead;
    if (r < 0) return -1;
    chunk = blockIdx.x;
    if (F.slab_rows == 0) return r;
    const int g = r / F.slab_rows, m = r - g * F.slab_rows;
    return (long long)g * F.group_rows + F.slab_t0 + m;
}

__device__ __forceinline__ RowId fast_row_of(const FastArgs& F, int r) {
    if (F.slab_rows == 0) return RowId{r, (F.row0 + r) % F.period};
    const int g = r / F.slab_rows, m = r - g * F.slab_rows;
    return RowId{(long long)g * F.group_rows + F.slab_t0 + m, F.slab_t0 + m};
}
__device__ __forceinline__ RowId fast_row(const FastArgs& F) { return fast_row_of(F, grid_row(F)); }
__device__ __forceinline__ bool fast_use_f64(const FastArgs& F, const RowId& rid) {
    return (uint64_t)F.q[rid.limb] < SMALL_PRIME_LIMIT && F.force_int == 0;
}

template <class A>
__device__ __forceinline__ typename A::C make_const(const FastArgs& F, int limb);
template <>
__device__ __forceinline__ F64C make_const<ArithF64>(const FastArgs& F, int limb) {
    const double q = (double)(uint64_t)F.q[limb];
    return F64C{q, F.qinv ? F.qinv[limb] : 1.0 / q};
}
template <>
__device__ __forceinline__ U64C make_const<ArithU64>(const FastArgs& F, int limb) {
    const uint64_t q = (uint64_t)F.q[limb];
    return U64C{q, q << 1};
}
template <class A>
__device__ __forceinline__ const typename A::TW* tw_row(const FastArgs& F, int limb);
template <>
__device__ __forceinline__ const double* tw_row<ArithF64>(const FastArgs& F, int limb) {
    return F.tw_f64 + ((long long)limb << F.logN);
}
template <>
__device__ __forceinline__ const ulonglong2* tw_row<ArithU64>(const FastArgs& F, int limb) {
    return F.tw_u64 + ((long long)limb << F.logN);
}
template <class A>
__device__ __forceinline__ const typename A::TW* twp_row(const FastArgs& F, int limb);
template <>
__device__ __forceinline__ const double* twp_row<ArithF64>(const FastArgs& F, int limb) {
    return F.twp_f64 ? F.twp_f64 + ((long long)limb << F.logN) : nullptr;
}
template <>
__device__ __forceinline__ const ulonglong2* twp_row<ArithU64>(const FastArgs& F, int limb) {
    return F.twp_u64 ? F.twp_u64 + ((long long)limb << F.logN) : nullptr;
}
template <class A>
__device__ __forceinline__ typename A::TW scalar_tw(const FastArgs& F, int limb);
template <>
__device__ __forceinline__ double scalar_tw<ArithF64>(const FastArgs& F, int limb) {
    return (double)F.scal[limb];
}
template <>
__device__ __forceinline__ ulonglong2 scalar_tw<ArithU64>(const FastArgs& F, int limb) {
    return make_ulonglong2((uint64_t)F.scal[limb], F.scal_sh[limb]);
}

// column passes: [data tile | 256 staged twiddles (F64 path) | mbarrier] = 39 KB per CTA; with a 64-register cap 4 CTAs/SM
#ifndef FAST_COL_CTAS
#define FAST_COL_CTAS 4
#endif
constexpr int COL_TW_SLOTS = 256;
constexpr int COL_SMEM_BYTES = SMEM_BYTES + COL_TW_SLOTS * 8 + 16;
// block passes: the (warp-private) exchange buffer only
#ifndef FAST_BLK_CTAS
#define FAST_BLK_CTAS 4
#endif
constexpr int BLK_SMEM_BYTES = SMEM_BYTES;

template <class TW>
struct TwSharedCol {     // column pass: the first 256 table entries, indexed exactly like the global table
    const TW* base;
    int j0;
    unsigned pre;
    static constexpr bool SMEM = true;
    __device__ __forceinline__ const TW* run(int i) const { return base + (1u << (j0 + i)) + (pre << i); }
};

// one thread arms the barrier and issues the bulk copy; everybody waits right before the first use
__device__ __forceinline__ void stage_col_twiddles(const double* __restrict__ W, double* tws, uint64_t* bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 256u * 8u);
        tma_bulk_g2s(tws, W, 256u * 8u, bar);
    }
}

// ---- forward pass A (column pass, stages 0..7) -------------------------------------------------------------
// RESC: the rescale of cc_mult (engine.py:1026-1038) fused into the load.  Row r of the batch is polynomial r / L,
// limb r % L; its coefficients are read from the INPUT ciphertext rows (in[poly], the rows that survive) together with
// the dropped limb r0[poly], and  x = ((in - r0) q0^-1 + [r0 > round_at]) R  goes straight into the first butterflies:
// one exact FP64 product (scale = q0^-1 R mod q is the reference's own table) instead of a separate kernel that
// writes the rescaled polynomial to HBM and reads it back.  60-bit rows evaluate the reference's integer formula.
struct RescaleIn {
    const int64_t* in[4];      // [L][N] rows of the four polynomials, in_stride apart
    long long in_stride;
    const int64_t* r0[4];      // [N] dropped limb of each polynomial
    const int64_t* scale;      // [L] q0^-1 R mod q_t
    long long round_at;
    const int64_t *_2q, *ql, *qh, *kl, *kh;
    int L;
};

// (B = logN - 8 is a template parameter of the column passes: the row stride 2^B becomes an immediate offset of every
//  load / store -- with a run-time shift the address arithmetic was a third of the FP64 path's instructions)
template <class A, int B>
struct RescaleLoad;
template <int B>
struct RescaleLoad<ArithF64, B> {
  static __device__ __forceinline__ void run(const FastArgs& F, const RescaleIn& R, double (&e)[16], int limb,
                                             long long drow, const F64C& c) {
    constexpr int b = B;
    const int tau = threadIdx.x;
    const int g = (int)(drow / R.L);
    const long long base = (long long)grid_chunk(F) * 16 + ((long long)(tau >> 4) << b) + (tau & 15);
    const int64_t* __restrict__ src = R.in[g] + (long long)limb * R.in_stride + base;
    const int64_t* __restrict__ z = R.r0[g] + base;
    const double s1 = (double)R.scale[limb], rm = (double)F.scal[limb];
    int64_t x[16], y[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        x[k] = src[(long long)(16 * k) << b];
        y[k] = z[(long long)(16 * k) << b];
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const double v = f64_mulmod(i2d(x[k] - y[k]), s1, c);
        e[k] = (y[k] > R.round_at) ? __dadd_rn(v, rm) : v;
    }
  }
};
template <int B>
struct RescaleLoad<ArithU64, B> {
  static __device__ __forceinline__ void run(const FastArgs& F, const RescaleIn& R, uint64_t (&e)[16], int limb,
                                             long long drow, const U64C& c) {
    constexpr int b = B;
    const int tau = threadIdx.x;
    const int g = (int)(drow / R.L);
    const long long base = (long long)grid_chunk(F) * 16 + ((long long)(tau >> 4) << b) + (tau & 15);
    const int64_t* __restrict__ src = R.in[g] + (long long)limb * R.in_stride + base;
    const int64_t* __restrict__ z = R.r0[g] + base;
    const LimbConst k = load_limb_const(R._2q, R.ql, R.qh, R.kl, R.kh, limb);
    const int64_t sc = R.scale[limb], q = (int64_t)c.q;
    const ulonglong2 s = scalar_tw<ArithU64>(F, limb);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int64_t x = src[(long long)(16 * i) << b], y = z[(long long)(16 * i) << b];
        int64_t o = reduce_q(mont_mul_ss(x - y, sc, k.q4, k.k) + (y > R.round_at ? 1 : 0), q);
        o += (o < 0) ? q : 0;
        e[i] = ArithU64::mul((uint64_t)o, s, c);
    }
  }
};

template <class A, int B, bool STAGED, bool RESC>
__device__ __forceinline__ void fast_fwd_col_body(const FastArgs& F, const RescaleIn& R, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    constexpr int b = B;
    const typename A::C c = make_const<A>(F, limb);
    int64_t* __restrict__ row0 = F.a + drow * F.a_stride + (long long)grid_chunk(F) * 16;
    const TW* __restrict__ W = tw_row<A>(F, limb);
    TW* tws = reinterpret_cast<TW*>(sm + SMEM_SLOTS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + SMEM_SLOTS + COL_TW_SLOTS);
    if constexpr (STAGED) stage_col_twiddles(reinterpret_cast<const double*>(W), reinterpret_cast<double*>(tws), bar);
    if constexpr (!RESC) {
        unsigned ca = 0;
        const long long ra = F.prefetch ? fast_row_ahead(F, F.prefetch, ca) : -1;
        if (ra >= 0) l2_prefetch_line(F.a + ra * F.a_stride + (long long)ca * 16 + ((long long)tau << b));
    }
    T e[16];
    {
        if constexpr (RESC) {
            RescaleLoad<A, B>::run(F, R, e, limb, drow, c);
        } else {
            const int r0 = tau >> 4, col = tau & 15;
            if (LAB_SKIP_MEM(F)) {
#pragma unroll
                for (int k = 0; k < 16; ++k) e[k] = A::load_in((int64_t)(tau * 16 + k), 0);
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) e[k] = A::load_in(row0[((long long)(r0 + 16 * k) << b) + col], F.in_raw);
            }
            if (F.scal) {
                const TW s = scalar_tw<A>(F, limb);
#pragma unroll
                for (int k = 0; k < 16; ++k) e[k] = A::mul(e[k], s, c);
            }
        }
        if constexpr (STAGED) {
            mbar_wait(bar, 0);
            if (!LAB_SKIP_MATH(F)) fast_fwd_round<A, 0>(e, TwSharedCol<TW>{tws, 0, 0u}, c);
        } else {
            if (!LAB_SKIP_MATH(F)) fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, 0, 0u}, c);
        }
        smx_store(sm, e, tau, 8);
    }
    __syncthreads();
    {
        smx_load(sm, e, tau, 4);
        const int hi = tau >> 4, col = tau & 15;
        if (!LAB_SKIP_MATH(F)) {
            if constexpr (STAGED)
                fast_fwd_round<A, 0>(e, TwSharedCol<TW>{tws, 4, (unsigned)hi}, c);
            else
                fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, 4, (unsigned)hi}, c);
        }
        if (LAB_SKIP_MEM(F)) {      // keep the values alive: one store that never happens
            int64_t acc = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) acc ^= A::store_mid(e[k], c);
            if (acc == 0x7fff123456789abcll) row0[0] = acc;
            return;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) row0[((long long)(hi * 16 + k) << b) + col] = A::store_mid(e[k], c);
    }
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, FAST_COL_CTAS) fast_fwd_colpass(const FastArgs F) {
    pdl_enter();
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    const RescaleIn none{};
    if (fast_use_f64(F, rid))
        fast_fwd_col_body<ArithF64, B, true, false>(F, none, sm, limb, rid.data_row);
    else
        fast_fwd_col_body<ArithU64, B, false, false>(F, none, sm, limb, rid.data_row);
}

// the tensor stage's column pass: rescale fused into the load (rows = 4 polynomials x L limbs, period L, F.scal = R mod q)
template <int B>
__global__ void __launch_bounds__(NTT_THREADS, FAST_COL_CTAS) fast_fwd_colpass_rescale(const FastArgs F, const RescaleIn R) {
    pdl_enter();
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    if (fast_use_f64(F, rid))
        fast_fwd_col_body<ArithF64, B, true, true>(F, R, sm, limb, rid.data_row);
    else
        fast_fwd_col_body<ArithU64, B, false, true>(F, R, sm, limb, rid.data_row);
}

// ---- ModUp basis extension for the fast path ---------------------------------------------------------------------
// One thread owns two coefficients of ONE partition, loads its Garner digits s_0..s_{alpha-1} (exact integers,
// engine.py:654-705) once, and then walks over ALL E target limbs: by Horner's rule in the target field
//      X = s_0 + m_0 (s_1 + m_1 (s_2 + ...)),   ext = X (scale-prime targets) or X * R (60-bit targets)
//      (== extend, engine.py:707-743, up to congruence and, for the scale primes, up to the Montgomery factor that
//      k_ksk_inner_fast no longer divides out)
// in FP64 for scale-prime targets (digits of scale-prime partitions are < 2^44; the single 60-bit digit of the
// base-prime partition is split into 31-bit halves), with the reference's Montgomery chain for 60-bit targets.
// The digits are read once instead of once per target; stores are coalesced 16-byte writes per target row.
struct ExtArgs {
    const int64_t* const* digit_ptrs;   // [P] -> [alpha][N] digit blocks, rows d_stride apart
    long long d_stride;
    const int32_t* alphas;              // [P]
    const int32_t* wide;                // [P] 1 if the partition's digits can exceed 2^51 (base-prime partition)
    const double* const* Hm;            // [P] -> [(alpha-1)][E] doubles: m_i mod q_t
    const double* Rd;                   // [E] R mod q_t
    const double* C31;                  // [E] 2^31 mod q_t
    const int64_t* const* Lenter;       // [P] -> [(alpha-1)][E] (L_i R^2) mod q_t      (60-bit targets)
    const int64_t* Rs;                  // [E] R^2 mod q_t
    const int64_t *q, *_2q, *ql, *qh, *kl, *kh;
    int64_t* out;                       // [P*E][N], row p*E + t
    int E, N;
    int raw;                            // scale-prime targets are written as raw doubles (read back with in_raw)
    const double* qinv;                 // [E] 1/q_t, or nullptr
};

// per-target constants of the extension staged once per CTA (q, 1/q, 2^31 mod q, the Horner multipliers): a thread
// walks over all E targets, and recomputing 1/q (an FP64 division) per target cost as much as the Horner chain itself
constexpr int EXT_MAX_E = 128;
struct ExtShared {
    double q[EXT_MAX_E], qinv[EXT_MAX_E], c31[EXT_MAX_E];
    double hm[7][EXT_MAX_E];
};

template <int AMAX>
__device__ __forceinline__ void ext_target(const ExtArgs& X, const ExtShared& S, int p, int t, int tl, int alpha, bool wide,
                                           const longlong2 (&s)[AMAX], const double (&dx)[AMAX], const double (&dy)[AMAX],
                                           const int64_t* __restrict__ le, int64_t* __restrict__ out) {
    longlong2 r;
    if (S.q[tl] < (double)SMALL_PRIME_LIMIT) {
        const F64C c{S.q[tl], S.qinv[tl]};
        double ax = 0.0, ay = 0.0;
        if (wide) {
            const double c31 = S.c31[tl];
            ax = __dadd_rn(f64_mulmod(dx[0], c31, c), dx[1]);
            ay = __dadd_rn(f64_mulmod(dy[0], c31, c), dy[1]);
        } else {
#pragma unroll
            for (int i = AMAX - 1; i >= 0; --i) {
                if (i == alpha - 1) {
                    ax = dx[i];
                    ay = dy[i];
                } else if (i < alpha - 1) {
                    const double m = S.hm[i][tl];
                    ax = __dadd_rn(f64_mulmod(ax, m, c), dx[i]);
                    ay = __dadd_rn(f64_mulmod(ay, m, c), dy[i]);
                }
            }
        }
        // scale-prime targets stay PLAIN (no Montgomery factor): the transform is linear and the FP64 inner product
        // then needs no R^-1 either -- X * (K R) is already the Montgomery-form product.  |ax| < 2^43: no reduction.
        r.x = X.raw ? (int64_t)__double_as_longlong(ax) : d2i(ax);
        r.y = X.raw ? (int64_t)__double_as_longlong(ay) : d2i(ay);
    } else {
        const LimbConst k = load_limb_const(X._2q, X.ql, X.qh, X.kl, X.kh, t);
        const int64_t q2 = (int64_t)k.q2, rs = X.Rs[t];
        r.x = mont_mul_ss(s[0].x, rs, k.q4, k.k);
        r.y = mont_mul_ss(s[0].y, rs, k.q4, k.k);
#pragma unroll
        for (int i = 0; i < AMAX - 1; ++i) {
            if (i < alpha - 1) {
                const int64_t l = le[(long long)i * X.E + t];
                r.x = lazy_add(r.x, mont_mul_ss(s[i + 1].x, l, k.q4, k.k), q2);
                r.y = lazy_add(r.y, mont_mul_ss(s[i + 1].y, l, k.q4, k.k), q2);
            }
        }
        r.x += (r.x < 0) ? q2 : 0;
        r.y += (r.y < 0) ? q2 : 0;
    }
    *reinterpret_cast<longlong2*>(out + (long long)t * X.N) = r;
}

// AMAX = capacity of the per-thread digit arrays (>= alpha, and >= 2 for the 31-bit split of a wide digit)
template <int AMAX>
__global__ void __launch_bounds__(256) k_extend_fast(const ExtArgs X, int t0, int t1) {
    pdl_enter();
    __shared__ ExtShared S;
    const int p = blockIdx.y;
    const int alpha = X.alphas[p];
    {   // stage the constants of targets [t0, t1) (t1 - t0 <= EXT_MAX_E, checked by the launcher)
        const double* __restrict__ hm = X.Hm[p];
        for (int i = threadIdx.x; i < t1 - t0; i += blockDim.x) {
            const double q = (double)(uint64_t)X.q[t0 + i];
            S.q[i] = q;
            S.qinv[i] = X.qinv ? X.qinv[t0 + i] : 1.0 / q;
            S.c31[i] = X.C31[t0 + i];
#pragma unroll
            for (int k = 0; k < 7; ++k) S.hm[k][i] = (k < alpha - 1 && k < AMAX - 1) ? hm[(long long)k * X.E + t0 + i] : 0.0;
        }
    }
    __syncthreads();
    const long long j = 2ll * (blockIdx.x * 256 + threadIdx.x);
    if (j >= X.N) return;
    const bool wide = X.wide[p] != 0;
    const int64_t* __restrict__ st = X.digit_ptrs[p];
    longlong2 s[AMAX];
    double dx[AMAX], dy[AMAX];
#pragma unroll
    for (int i = 0; i < AMAX; ++i) {
        s[i] = make_longlong2(0, 0);
        if (i < alpha) s[i] = *reinterpret_cast<const longlong2*>(st + (long long)i * X.d_stride + j);
        dx[i] = i2d(s[i].x);
        dy[i] = i2d(s[i].y);
    }
    if (wide) {
        dx[0] = (double)(int)(s[0].x >> 31); dy[0] = (double)(int)(s[0].y >> 31);
        dx[1] = (double)(int)(s[0].x & 0x7FFFFFFF); dy[1] = (double)(int)(s[0].y & 0x7FFFFFFF);
    }
    const int64_t* __restrict__ le = X.Lenter[p];
    int64_t* __restrict__ out = X.out + ((long long)p * X.E) * X.N + j;
    int t = t0;
    for (; t + 3 < t1; t += 4) {   // four independent targets (8 Horner chains) in flight per iteration
        ext_target<AMAX>(X, S, p, t, t - t0, alpha, wide, s, dx, dy, le, out);
        ext_target<AMAX>(X, S, p, t + 1, t + 1 - t0, alpha, wide, s, dx, dy, le, out);
        ext_target<AMAX>(X, S, p, t + 2, t + 2 - t0, alpha, wide, s, dx, dy, le, out);
        ext_target<AMAX>(X, S, p, t + 3, t + 3 - t0, alpha, wide, s, dx, dy, le, out);
    }
    for (; t < t1; ++t) ext_target<AMAX>(X, S, p, t, t - t0, alpha, wide, s, dx, dy, le, out);
}

// ---- evaluation-key inner product for the fast path ------------------------------------------------------------
// acc_h[t] = R^-1 * sum_p ext[p][t] * key_h[p][t]  (== the Montgomery products + running mont_add of
// engine.py:906-937, 832-840 up to congruence).  FP64 for scale-prime rows, Montgomery for the 60-bit rows.
struct InnerArgs {
    const int64_t* ext;                 // [P*E][N] canonical NTT-domain values, row p*E + t
    const int64_t* const* k0;           // [P] row-0 pointers of the key halves (rows k_stride apart)
    const int64_t* const* k1;
    long long k_stride;
    int64_t *acc0, *acc1;               // [E][N]
    const double* Rinv;                 // [E] R^-1 mod q_t
    const int64_t *q, *_2q, *ql, *qh, *kl, *kh;
    int P, E, N, t0;
    int raw;                            // scale-prime rows: ext holds raw doubles and acc is written as raw doubles
    const double* qinv;                 // [E] 1/q_t, or nullptr
};

__global__ void __launch_bounds__(256) k_ksk_inner_fast(const InnerArgs X) {
    pdl_enter();
    const int t = X.t0 + blockIdx.y;
    const long long j = 2ll * (blockIdx.x * 256 + threadIdx.x);
    if (j >= X.N) return;
    const uint64_t q = (uint64_t)X.q[t];
    longlong2 r0, r1;
    if (q < SMALL_PRIME_LIMIT) {
        const F64C c{(double)q, X.qinv ? X.qinv[t] : 1.0 / (double)q};
        double a0x = 0.0, a0y = 0.0, a1x = 0.0, a1y = 0.0;
        for (int p = 0; p < X.P; ++p) {
            const longlong2 e = *reinterpret_cast<const longlong2*>(X.ext + ((long long)p * X.E + t) * X.N + j);
            const longlong2 u = *reinterpret_cast<const longlong2*>(X.k0[p] + (long long)t * X.k_stride + j);
            const longlong2 v = *reinterpret_cast<const longlong2*>(X.k1[p] + (long long)t * X.k_stride + j);
            const double ex = X.raw ? __longlong_as_double(e.x) : i2d(e.x);
            const double ey = X.raw ? __longlong_as_double(e.y) : i2d(e.y);
            a0x = __dadd_rn(a0x, f64_mulmod(ex, i2d(u.x), c));
            a0y = __dadd_rn(a0y, f64_mulmod(ey, i2d(u.y), c));
            a1x = __dadd_rn(a1x, f64_mulmod(ex, i2d(v.x), c));
            a1y = __dadd_rn(a1y, f64_mulmod(ey, i2d(v.y), c));
        }
        // ext is plain for these rows (k_extend_fast), the key is in Montgomery form: the sum already is the
        // Montgomery-form product; at most 13 terms below 0.54 q each, |a| < 2^45, which the inverse transform accepts
        if (X.raw) {
            r0 = make_longlong2(__double_as_longlong(a0x), __double_as_longlong(a0y));
            r1 = make_longlong2(__double_as_longlong(a1x), __double_as_longlong(a1y));
        } else {
            r0 = make_longlong2(d2i(a0x), d2i(a0y));
            r1 = make_longlong2(d2i(a1x), d2i(a1y));
        }
    } else {
        const LimbConst k = load_limb_const(X._2q, X.ql, X.qh, X.kl, X.kh, t);
        const int64_t q2 = (int64_t)k.q2;
        r0 = make_longlong2(0, 0);
        r1 = make_longlong2(0, 0);
        for (int p = 0; p < X.P; ++p) {
            const longlong2 e = *reinterpret_cast<const longlong2*>(X.ext + ((long long)p * X.E + t) * X.N + j);
            const longlong2 u = *reinterpret_cast<const longlong2*>(X.k0[p] + (long long)t * X.k_stride + j);
            const longlong2 v = *reinterpret_cast<const longlong2*>(X.k1[p] + (long long)t * X.k_stride + j);
            r0.x = lazy_add(r0.x, mont_mul_ss(e.x, u.x, k.q4, k.k), q2);
            r0.y = lazy_add(r0.y, mont_mul_ss(e.y, u.y, k.q4, k.k), q2);
            r1.x = lazy_add(r1.x, mont_mul_ss(e.x, v.x, k.q4, k.k), q2);
            r1.y = lazy_add(r1.y, mont_mul_ss(e.y, v.y, k.q4, k.k), q2);
        }
    }
    *reinterpret_cast<longlong2*>(X.acc0 + (long long)t * X.N + j) = r0;
    *reinterpret_cast<longlong2*>(X.acc1 + (long long)t * X.N + j) = r1;
}

// ---- ModDown for the fast path (ordinary rows) -------------------------------------------------------------------
// out_t = [add_t +] (...((D_t - s_0) P_0^-1 - s_1) P_1^-1 ...) mod q_t, canonical, where s_i are the EXACT effective
// special-limb values produced by k_moddown_special (they can exceed 2^51: split into 31-bit halves).
// Same value as engine.py:851-901 (+ :1135-1140 / :947-948), whose ordinary rows end in reduce_2q.
struct ModDownArgs {
    const int64_t* d;                   // [E][N] plain canonical rows (ordinary first)
    const int64_t* eff;                 // [K][N] effective special values (exact integers in [0, 2 q_special))
    const int64_t* add;                 // optional [L][N] (row stride add_stride)
    long long add_stride;
    int64_t* out;                       // [L][N]
    long long out_stride;
    const double* Pinv;                 // [K][E] P_i^-1 mod q_t as doubles
    const double* C31;                  // [E]
    const int64_t* q;
    int L, K, E, N;
    const double* qinv;                 // [E] 1/q_t, or nullptr
    unsigned add_ginv;                  // != 0: the addend is the Galois image of `add` (g^-1 mod 2N), gathered on the fly
};

__global__ void __launch_bounds__(256) k_moddown_fast(const ModDownArgs X) {
    pdl_enter();
    const int t = blockIdx.y;
    const long long j = 2ll * (blockIdx.x * 256 + threadIdx.x);
    if (j >= X.N) return;
    const uint64_t q = (uint64_t)X.q[t];
    const F64C c{(double)q, X.qinv ? X.qinv[t] : 1.0 / (double)q};
    const double c31 = X.C31[t];
    const longlong2 dv = *reinterpret_cast<const longlong2*>(X.d + (long long)t * X.N + j);
    double vx = i2d(dv.x), vy = i2d(dv.y);
    for (int i = 0; i < X.K; ++i) {
        const longlong2 s = *reinterpret_cast<const longlong2*>(X.eff + (long long)i * X.N + j);
        const double sx = __dadd_rn(f64_mulmod((double)(int)(s.x >> 31), c31, c), (double)(int)(s.x & 0x7FFFFFFF));
        const double sy = __dadd_rn(f64_mulmod((double)(int)(s.y >> 31), c31, c), (double)(int)(s.y & 0x7FFFFFFF));
        const double pinv = X.Pinv[(long long)i * X.E + t];
        vx = f64_mulmod(__dadd_rn(vx, -sx), pinv, c);
        vy = f64_mulmod(__dadd_rn(vy, -sy), pinv, c);
    }
    longlong2 r;
    r.x = ArithF64::store_canon(vx, c, false);
    r.y = ArithF64::store_canon(vy, c, false);
    if (X.add) {   // the reference's integer tail mont_add + reduce_2q, exact for ANY addend (conjugate feeds signed values)
        const int64_t qi = (int64_t)q;
        longlong2 a;
        if (X.add_ginv) {   // rotate: the addend is the rotated c0, read from the unrotated ciphertext (canonical rows)
            const int64_t* __restrict__ row = X.add + (long long)t * X.add_stride;
            const unsigned t0 = (X.add_ginv * (unsigned)j) & (2u * X.N - 1), t1 = (X.add_ginv * ((unsigned)j + 1)) & (2u * X.N - 1);
            const int64_t v0 = row[t0 & (X.N - 1)], v1 = row[t1 & (X.N - 1)];
            a.x = (t0 >= (unsigned)X.N && v0 != 0) ? qi - v0 : v0;
            a.y = (t1 >= (unsigned)X.N && v1 != 0) ? qi - v1 : v1;
        } else {
            a = *reinterpret_cast<const longlong2*>(X.add + (long long)t * X.add_stride + j);
        }
        r.x = reduce_q(lazy_add(a.x, r.x, 2 * qi), qi);
        r.y = reduce_q(lazy_add(a.y, r.y, 2 * qi), qi);
    }
    *reinterpret_cast<longlong2*>(X.out + (long long)t * X.out_stride + j) = r;
}

// ---- block passes (forward stages 8..logN-1 / inverse levels 0..B-1), WARP-INDEPENDENT ---------------------------------
// The stages of a block pass only mix coefficients inside 2^B-point sub-blocks (B <= 9), i.e. inside the 512 contiguous
// coefficients a warp owns (32 threads x 16).  Every exchange therefore stays inside the warp's private slice of the
// exchange buffer and needs __syncwarp() only -- no CTA barrier anywhere, so the 8 warps of a CTA drift apart and keep
// the FP64 pipe fed while others wait on loads.  Twiddles go straight from L2 to registers: the stages whose twiddles
// several threads share are broadcast loads from the plain table, the last group (one thread per twiddle) comes from the
// PACKED table as lane-contiguous 128-bit loads.  Nothing is staged through shared memory, no mbarrier, no prologue
// barrier: round 1 profiling showed the L1 data pipe (75-83 % busy, 80 % of it global accesses that touch up to 32 lines
// per warp instruction), not shared memory, to be what the FP64 pipe was waiting for.
//
// The thread that ends (forward) / starts (inverse) with coefficients 16 t .. 16 t + 15 of the tile moves them either as
// four 256-bit accesses (natural order) or, with F.perm, as eight lane-contiguous 128-bit accesses (warp-interleaved order).
template <class T>
__device__ __forceinline__ void tile_store16(int64_t* __restrict__ g, const int64_t (&r)[16], int tau, int perm) {
    if (perm) {
        longlong2* __restrict__ o = reinterpret_cast<longlong2*>(g + (tau >> 5) * PACK_TILE) + (tau & 31);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j * 32] = make_longlong2(r[2 * j], r[2 * j + 1]);
    } else {
        int64_t* o = g + tau * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) stg256(o, r, j);
    }
}
__device__ __forceinline__ void tile_load16(const int64_t* __restrict__ g, int64_t (&r)[16], int tau, int perm) {
    if (perm) {
        const longlong2* __restrict__ in = reinterpret_cast<const longlong2*>(g + (tau >> 5) * PACK_TILE) + (tau & 31);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const longlong2 v = in[j * 32];
            r[2 * j] = v.x;
            r[2 * j + 1] = v.y;
        }
    } else {
        const int64_t* in = g + tau * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) ldg256(in, r, j);
    }
}

// the round on field [3:0] (a thread's 16 contiguous coefficients): packed twiddles when the table is there
template <class A, int FIRST>
__device__ __forceinline__ void fast_fwd_last_round(typename A::T (&e)[16], const FastArgs& F, int limb, unsigned chunk,
                                                    const typename A::C& c) {
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const TW* __restrict__ WP = twp_row<A>(F, limb);
    if (WP)
        fast_fwd_round<A, FIRST>(e, TwPacked<TW>{WP + ((long long)(chunk * 8u + (unsigned)(tau >> 5))) * PACK_TILE, tau & 31}, c);
    else
        fast_fwd_round<A, FIRST>(e, TwGlobal<TW>{tw_row<A>(F, limb), F.logN - 4, (chunk << 8) | (unsigned)tau}, c);
}
template <class A>
__device__ __forceinline__ void fast_inv_first_round(typename A::T (&e)[16], const FastArgs& F, int limb, unsigned chunk,
                                                     const typename A::C& c) {
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const TW* __restrict__ WP = twp_row<A>(F, limb);
    if (WP)
        fast_inv_round<A, 4>(e, TwPacked<TW>{WP + ((long long)(chunk * 8u + (unsigned)(tau >> 5))) * PACK_TILE, tau & 31}, c);
    else
        fast_inv_round<A, 4>(e, TwGlobal<TW>{tw_row<A>(F, limb), F.logN - 4, (chunk << 8) | (unsigned)tau}, c);
}

template <class A, int B>
__device__ __forceinline__ void fast_fwd_block_body(const FastArgs& F, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const unsigned chunk = grid_chunk(F);
    constexpr int logN = B + 8;
    const typename A::C c = make_const<A>(F, limb);
    int64_t* __restrict__ g = F.a + drow * F.a_stride + (long long)chunk * TILE;
    const TW* __restrict__ W = tw_row<A>(F, limb);
    if (tau == 32) {
        unsigned ca = 0;
        const long long ra = F.prefetch ? fast_row_ahead(F, F.prefetch, ca) : -1;
        if (ra >= 0) l2_prefetch_bulk(F.a + ra * F.a_stride + (long long)ca * TILE, TILE * 8u);
    }
    T e[16];
    constexpr int P1 = B - 4;
    if (LAB_SKIP_MEM(F)) {
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = A::load_in((int64_t)(tau * 16 + k), 0);
    } else {
        const int zb = zbase(tau, P1);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = A::load_mid(g[zb | (k << P1)]);
    }
    if constexpr (B > 4) {
        // The last round's twiddles (the warp tile's packed block, 4 or 8 KB = one 128-byte line per lane and half) are
        // needed only after the first rounds, and there is no register to hold them until then: touch the lines now, so
        // that the loads of the last round hit L1 instead of waiting ~600 clk for L2 at each of its four stages
        // (14 % of the FP64 path's stall samples in profiles/r02_ncu_fwd_ntt_v2_steady.txt).
        const TW* __restrict__ WP = twp_row<A>(F, limb);
        if (WP) {
            const char* line = reinterpret_cast<const char*>(WP + ((long long)(chunk * 8u + (unsigned)(tau >> 5))) * PACK_TILE) +
                               (tau & 31) * 128;
            l1_touch(line);
            if constexpr (sizeof(TW) == 16) l1_touch(line + 4096);
        }
    }
    if constexpr (B == 4) {
        fast_fwd_last_round<A, 0>(e, F, limb, chunk, c);
    } else {
        if (!LAB_SKIP_MATH(F))
            fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P1, (chunk << (8 - P1)) | (unsigned)(tau >> P1)}, c);
        smx_store(sm, e, tau, P1);
        __syncwarp();
        if constexpr (B >= 8) {
            constexpr int P2 = B - 8;
            smx_load(sm, e, tau, P2);
            if constexpr (B == 8) {
                if (!LAB_SKIP_MATH(F)) fast_fwd_last_round<A, 0>(e, F, limb, chunk, c);
            } else {
                fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P2, (chunk << (8 - P2)) | (unsigned)(tau >> P2)}, c);
                __syncwarp();
                smx_store(sm, e, tau, P2);
                __syncwarp();
                smx_load(sm, e, tau, 0);
                fast_fwd_last_round<A, 3>(e, F, limb, chunk, c);
            }
        } else {
            smx_load(sm, e, tau, 0);
            fast_fwd_last_round<A, 8 - B>(e, F, limb, chunk, c);
        }
    }
    int64_t r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = A::store_out(e[k], c, F.out_raw);
    if (LAB_SKIP_MEM(F)) {
        int64_t acc = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc ^= r[k];
        if (acc == 0x7fff123456789abcll) g[0] = acc;
        return;
    }
    tile_store16<T>(g, r, tau, F.perm);
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, FAST_BLK_CTAS) fast_fwd_blockpass(const FastArgs F) {
    pdl_enter();
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    if (fast_use_f64(F, rid))
        fast_fwd_block_body<ArithF64, B>(F, sm, rid.limb, rid.data_row);
    else
        fast_fwd_block_body<ArithU64, B>(F, sm, rid.limb, rid.data_row);
}

// inverse: for B == 9 the last level (distance 256) is taken as the top stage of field [8:5], which keeps it warp-private too
template <class A, int B>
__device__ __forceinline__ void fast_inv_block_body(const FastArgs& F, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const unsigned chunk = grid_chunk(F);
    constexpr int logN = B + 8;
    const typename A::C c = make_const<A>(F, limb);
    int64_t* __restrict__ g = F.a + drow * F.a_stride + (long long)chunk * TILE;
    const TW* __restrict__ W = tw_row<A>(F, limb);
    if (tau == 32) {
        unsigned ca = 0;
        const long long ra = F.prefetch ? fast_row_ahead(F, F.prefetch, ca) : -1;
        if (ra >= 0) l2_prefetch_bulk(F.a + ra * F.a_stride + (long long)ca * TILE, TILE * 8u);
    }
    T e[16];
    {
        int64_t r[16];
        tile_load16(g, r, tau, F.perm);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = A::load_in(r[k], F.in_raw);
    }
    fast_inv_first_round<A>(e, F, limb, chunk, c);
    if constexpr (B == 4) {
        int64_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = A::store_mid(e[k], c);
        tile_store16<T>(g, r, tau, 0);
    } else {
        smx_store(sm, e, tau, 0);
        __syncwarp();
        smx_load(sm, e, tau, 4);
        constexpr int NST = (B >= 8) ? 4 : B - 4;
        fast_inv_round<A, NST>(e, TwGlobal<TW>{W, logN - 8, (chunk << 4) | (unsigned)(tau >> 4)}, c);
        if constexpr (B == 9) {
            __syncwarp();
            smx_store(sm, e, tau, 4);
            __syncwarp();
            smx_load(sm, e, tau, 5);
            {   // level with distance 2^8 = top stage of field [8:5]: pairs (k, k+8), one twiddle per 512-point sub-block
                TW w[8];
                w[0] = __ldg(W + (1u << (logN - 9)) + ((chunk << 3) | (unsigned)(tau >> 5)));
                A::template gs_stage<3>(e, w, c);
#pragma unroll
                for (int k = 0; k < 16; ++k) A::tame(e[k], c);
            }
            const int zb = zbase(tau, 5);
#pragma unroll
            for (int k = 0; k < 16; ++k) g[zb | (k << 5)] = A::store_mid(e[k], c);
        } else {
            const int zb = zbase(tau, 4);
#pragma unroll
            for (int k = 0; k < 16; ++k) g[zb | (k << 4)] = A::store_mid(e[k], c);
        }
    }
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, FAST_BLK_CTAS) fast_inv_blockpass(const FastArgs F) {
    pdl_enter();
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    if (fast_use_f64(F, rid))
        fast_inv_block_body<ArithF64, B>(F, sm, rid.limb, rid.data_row);
    else
        fast_inv_block_body<ArithU64, B>(F, sm, rid.limb, rid.data_row);
}


// ---- inverse pass A' (levels b..logN-1), x scalar, canonical out ---------------------------------------------
template <class A, int B, bool STAGED>
__device__ __forceinline__ void fast_inv_col_body(const FastArgs& F, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    constexpr int b = B;
    const typename A::C c = make_const<A>(F, limb);
    int64_t* __restrict__ row0 = F.a + drow * F.a_stride + (long long)grid_chunk(F) * 16;
    const TW* __restrict__ W = tw_row<A>(F, limb);
    TW* tws = reinterpret_cast<TW*>(sm + SMEM_SLOTS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + SMEM_SLOTS + COL_TW_SLOTS);
    if constexpr (STAGED) stage_col_twiddles(reinterpret_cast<const double*>(W), reinterpret_cast<double*>(tws), bar);
    {
        unsigned ca = 0;
        const long long ra = F.prefetch ? fast_row_ahead(F, F.prefetch, ca) : -1;
        if (ra >= 0) l2_prefetch_line(F.a + ra * F.a_stride + (long long)ca * 16 + ((long long)tau << b));
    }
    T e[16];
    {
        const int hi = tau >> 4, col = tau & 15;
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = A::load_mid(row0[((long long)(hi * 16 + k) << b) + col]);
        if constexpr (STAGED) {
            mbar_wait(bar, 0);
            fast_inv_round<A, 4>(e, TwSharedCol<TW>{tws, 4, (unsigned)hi}, c);
        } else {
            fast_inv_round<A, 4>(e, TwGlobal<TW>{W, 4, (unsigned)hi}, c);
        }
        smx_store(sm, e, tau, 4);
    }
    __syncthreads();
    {
        smx_load(sm, e, tau, 8);
        if constexpr (STAGED)
            fast_inv_round<A, 4>(e, TwSharedCol<TW>{tws, 0, 0u}, c);
        else
            fast_inv_round<A, 4>(e, TwGlobal<TW>{W, 0, 0u}, c);
        const TW s = scalar_tw<A>(F, limb);
        const int r0 = tau >> 4, col = tau & 15;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            row0[((long long)(r0 + 16 * k) << b) + col] = A::store_canon(A::mul(e[k], s, c), c, F.centred != 0);
    }
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, FAST_COL_CTAS) fast_inv_colpass(const FastArgs F) {
    pdl_enter();
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    if (fast_use_f64(F, rid))
        fast_inv_col_body<ArithF64, B, true>(F, sm, limb, rid.data_row);
    else
        fast_inv_col_body<ArithU64, B, false>(F, sm, limb, rid.data_row);
}

// ---- table construction ----------------------------------------------------------------------------------------
// plain canonical twiddles [C][N] -> {w, floor(w 2^64 / q)} and double(w)
__global__ void fast_tables_kernel(const int64_t* __restrict__ plain, const int64_t* __restrict__ q,
                                   ulonglong2* __restrict__ sh, double* __restrict__ dbl, int N) {
    const int i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const uint64_t w = (uint64_t)plain[(long long)i * N + j];
    const unsigned __int128 num = ((unsigned __int128)w) << 64;
    const uint64_t wp = (uint64_t)(num / (unsigned __int128)(uint64_t)q[i]);
    sh[(long long)i * N + j] = make_ulonglong2(w, wp);
    if (dbl) dbl[(long long)i * N + j] = (double)w;
}

// packed last-group tables (TwPacked) from the plain fast tables: thread T = coefficient index / 16
__global__ void fast_pack_kernel(const ulonglong2* __restrict__ sh, const double* __restrict__ dbl, ulonglong2* __restrict__ psh,
                                 double* __restrict__ pdbl, int logN) {
    const int limb = blockIdx.y;
    const int T = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = 1 << logN;
    if (T >= (N >> 4)) return;
    const long long row = (long long)limb << logN;
    const long long wt = row + (long long)(T >> 5) * PACK_TILE;
    const int lane = T & 31;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < (1 << i); ++j) {
            const long long src = row + (1ll << (logN - 4 + i)) + ((long long)T << i) + j;
            if (psh) psh[wt + ((1 << i) - 1 + j) * 32 + lane] = sh[src];
            if (pdbl) {
                if (i == 0) pdbl[wt + lane] = dbl[src];
                else pdbl[wt + 64 + ((((1 << (i - 1)) - 1) + (j >> 1)) * 32 + lane) * 2 + (j & 1)] = dbl[src];
            }
        }
    }
    if (psh) psh[wt + 15 * 32 + lane] = make_ulonglong2(0, 0);
    if (pdbl) pdbl[wt + 32 + lane] = 0.0;
}

// natural <-> warp-interleaved order of NTT-domain rows (FastArgs::perm); used once per evaluation key
__global__ void fast_perm_kernel(const int64_t* __restrict__ in, long long in_stride, int64_t* __restrict__ out, long long out_stride,
                                 int N, int inverse) {
    const long long row = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // natural index
    if (i >= N) return;
    const int tile = i >> 9, t = (i >> 4) & 31, k = i & 15;
    const int p = (tile << 9) + (((k >> 1) * 32 + t) << 1) + (k & 1);
    if (inverse) out[row * out_stride + i] = in[row * in_stride + p];
    else out[row * out_stride + p] = in[row * in_stride + i];
}

}  // namespace ckks
