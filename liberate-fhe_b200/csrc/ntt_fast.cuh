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
// Montgomery form.  Tables: shoup[C][N] = {w, w'} (16 B) and dbl[C][N] (8 B, small primes only).
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
template <class TW>
struct TwShared {
    const TW* const* stage;   // stage[j] = shared-memory array of pass-local stage j
    int j0;                   // pass-local stage of i == 0
    unsigned pre;             // CTA-local index bits above the field
    static constexpr bool SMEM = true;
    __device__ __forceinline__ const TW* run(int i) const { return stage[j0 + i] + (pre << i); }
};

// ------------------------------------------------------------------------------------------------------------
// rounds
// ------------------------------------------------------------------------------------------------------------
template <class A, int I, class SRC>
__device__ __forceinline__ void fast_fwd_stage(typename A::T (&e)[16], const SRC& src, const typename A::C& c) {
    typename A::TW w[8];
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
    int force_int;              // 1: integer path for every limb; m > 1: for every m-th row (INT + FP64 pipes side by side)
    // slab view of a [groups][group_rows] block: grid row r -> group r / slab_rows, member r % slab_rows;
    // data row = group * group_rows + slab_t0 + member, limb = slab_t0 + member.  slab_rows == 0: plain rows.
    int slab_rows, group_rows, slab_t0;
    int prefetch;               // rows ahead whose tile is pulled into L2 by every CTA (0 = off)
    int swap_grid;              // one-tile-per-CTA kernels: blockIdx.x = row, blockIdx.y = chunk
    int row0;                   // plain rows: limb of grid row r is (row0 + r) % period (row slabs of one batch)
    int in_raw, out_raw;        // FP64 rows: the input / the forward output are raw doubles (fused key switch), not int64
    // slab views of a key switch: partition g owns target limbs [own_row0[g], own_row0[g] + own_alpha[g]) (own_row0 < 0:
    // none on this device).  Those rows are NOT transformed: their NTT is the tensor product's d2 itself.
    const int32_t* own_row0;
    const int32_t* own_alpha;
};

// grid mapping of the one-tile-per-CTA kernels: (chunk, row) by default; swapped (row, chunk) when F.swap_grid is
// set, so that consecutive CTAs belong to consecutive rows and the integer-path rows (60-bit limbs) are spread
// finely among the FP64 rows -- the two kinds of tile load different pipes and overlap when they share an SM.
__device__ __forceinline__ int grid_row(const FastArgs& F) { return F.swap_grid ? blockIdx.x : blockIdx.y; }
__device__ __forceinline__ int grid_rows(const FastArgs& F) { return F.swap_grid ? gridDim.x : gridDim.y; }
__device__ __forceinline__ unsigned grid_chunk(const FastArgs& F) { return F.swap_grid ? blockIdx.y : blockIdx.x; }

struct RowId {
    long long data_row;
    int limb;
};
__device__ __forceinline__ bool own_target(const int32_t* row0, const int32_t* alpha, int part, int t) {
    if (!row0) return false;
    const int r0 = row0[part];
    return r0 >= 0 && t >= r0 && t < r0 + alpha[part];
}
__device__ __forceinline__ bool fast_skip_own(const FastArgs& F) {   // CTA-uniform
    if (!F.own_row0 || F.slab_rows == 0) return false;
    const int r = grid_row(F), g = r / F.slab_rows;
    return own_target(F.own_row0, F.own_alpha, g, F.slab_t0 + (r - g * F.slab_rows));
}
// the tile that a CTA dispatched about `ahead` rows x (chunks per row) tiles later will load: data row (or -1) and chunk
__device__ __forceinline__ long long fast_row_ahead(const FastArgs& F, int ahead, unsigned& chunk) {
    int r;
    if (F.swap_grid) {   // dispatch order: row fastest, then chunk
        const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + (long long)ahead * gridDim.y;
        const int c = (int)(lin / gridDim.x);
        if (c >= (int)gridDim.y) return -1;
        r = (int)(lin - (long long)c * gridDim.x);
        chunk = (unsigned)c;
    } else {
        r = blockIdx.y + ahead;
        if (r >= (int)gridDim.y) return -1;
        chunk = blockIdx.x;
    }
    if (F.slab_rows == 0) return r;
    const int g = r / F.slab_rows, m = r - g * F.slab_rows;
    return (long long)g * F.group_rows + F.slab_t0 + m;
}

__device__ __forceinline__ bool fast_use_f64(const FastArgs& F, const RowId& rid);
__device__ __forceinline__ RowId fast_row(const FastArgs& F) {
    const int r = grid_row(F);
    if (F.slab_rows == 0) return RowId{r, (F.row0 + r) % F.period};
    const int g = r / F.slab_rows, m = r - g * F.slab_rows;
    return RowId{(long long)g * F.group_rows + F.slab_t0 + m, F.slab_t0 + m};
}
__device__ __forceinline__ bool fast_use_f64(const FastArgs& F, const RowId& rid) {
    if ((uint64_t)F.q[rid.limb] >= SMALL_PRIME_LIMIT) return false;
    if (F.force_int == 0) return true;
    if (F.force_int == 1) return false;
    return (rid.data_row % F.force_int) != (F.force_int - 1);
}

template <class A>
__device__ __forceinline__ typename A::C make_const(uint64_t q);
template <>
__device__ __forceinline__ F64C make_const<ArithF64>(uint64_t q) {
    return F64C{(double)q, 1.0 / (double)q};
}
template <>
__device__ __forceinline__ U64C make_const<ArithU64>(uint64_t q) {
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
__device__ __forceinline__ typename A::TW scalar_tw(const FastArgs& F, int limb);
template <>
__device__ __forceinline__ double scalar_tw<ArithF64>(const FastArgs& F, int limb) {
    return (double)F.scal[limb];
}
template <>
__device__ __forceinline__ ulonglong2 scalar_tw<ArithU64>(const FastArgs& F, int limb) {
    return make_ulonglong2((uint64_t)F.scal[limb], F.scal_sh[limb]);
}

// shared memory of the fast kernels: [data tile | staged twiddles (F64 path) | mbarrier]
#ifndef FAST_CTAS_PER_SM
#define FAST_CTAS_PER_SM 3
#endif
constexpr int FAST_TW_SLOTS = 4096;
constexpr int FAST_SMEM_BYTES = SMEM_BYTES + FAST_TW_SLOTS * 8 + 16;
// column passes stage only 256 twiddles: 39 KB per CTA and a 64-register cap give 4 CTAs/SM (col pass 94 -> 89 us per 380 limbs)
#ifndef FAST_COL_CTAS
#define FAST_COL_CTAS 4
#endif
constexpr int COL_TW_SLOTS = 256;
// "hybrid" block passes (B >= 7): only the twiddles that threads share (the first four stages of the pass) are staged;
// the later stages' twiddles are used by exactly one thread each and are read from L2 directly -- 41 KB per CTA, 4 CTAs/SM
constexpr int HYB_TW_SLOTS = 512;
constexpr int HYB_SMEM_BYTES = SMEM_BYTES + HYB_TW_SLOTS * 8 + 16;
constexpr int COL_SMEM_BYTES = SMEM_BYTES + COL_TW_SLOTS * 8 + 16;

template <class TW>
struct TwSharedBlock {   // block pass: stage j (global stage 8+j) holds 2^(j+unit_log) twiddles, stages back to back
    const TW* base;
    int unit_log, j0;
    unsigned pre;
    static constexpr bool SMEM = true;
    __device__ __forceinline__ const TW* run(int i) const {
        return base + (((1u << (j0 + i)) - 1u) << unit_log) + (pre << i);
    }
};
template <class TW>
struct TwSharedCol {     // column pass: the first 256 table entries, indexed exactly like the global table
    const TW* base;
    int j0;
    unsigned pre;
    static constexpr bool SMEM = true;
    __device__ __forceinline__ const TW* run(int i) const { return base + (1u << (j0 + i)) + (pre << i); }
};

// one thread arms the barrier and issues the bulk copies; everybody waits right before the first use
__device__ __forceinline__ void stage_block_twiddles(const double* __restrict__ W, double* tws, uint64_t* bar, int B,
                                                     unsigned chunk) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int unit_log = 12 - B;
        mbar_expect_tx(bar, (unsigned)(((1u << 12) - (1u << unit_log)) * 8u));
        for (int j = 0; j < B; ++j) {
            const unsigned cnt = 1u << (j + unit_log);
            tma_bulk_g2s(tws + (((1u << j) - 1u) << unit_log), W + (1u << (8 + j)) + (size_t)chunk * cnt, cnt * 8u, bar);
        }
    }
}
// the first `nst` stages only (the shared ones; the unshared twiddles of the later stages then come straight from L2)
__device__ __forceinline__ void stage_block_twiddles_first(const double* __restrict__ W, double* tws, uint64_t* bar, int B,
                                                           unsigned chunk, int nst) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int unit_log = 12 - B;
        mbar_expect_tx(bar, (unsigned)((((1u << nst) - 1u) << unit_log) * 8u));
        for (int j = 0; j < nst; ++j) {
            const unsigned cnt = 1u << (j + unit_log);
            tma_bulk_g2s(tws + (((1u << j) - 1u) << unit_log), W + (1u << (8 + j)) + (size_t)chunk * cnt, cnt * 8u, bar);
        }
    }
}
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

template <class A>
__device__ __forceinline__ void rescale_load(const FastArgs& F, const RescaleIn& R, typename A::T (&e)[16], int limb,
                                             long long drow, const typename A::C& c);
template <>
__device__ __forceinline__ void rescale_load<ArithF64>(const FastArgs& F, const RescaleIn& R, double (&e)[16], int limb,
                                                       long long drow, const F64C& c) {
    const int tau = threadIdx.x, b = F.logN - 8;
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
template <>
__device__ __forceinline__ void rescale_load<ArithU64>(const FastArgs& F, const RescaleIn& R, uint64_t (&e)[16], int limb,
                                                       long long drow, const U64C& c) {
    const int tau = threadIdx.x, b = F.logN - 8;
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

template <class A, bool STAGED, bool RESC>
__device__ __forceinline__ void fast_fwd_col_body(const FastArgs& F, const RescaleIn& R, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const int b = F.logN - 8;
    const typename A::C c = make_const<A>((uint64_t)F.q[limb]);
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
            rescale_load<A>(F, R, e, limb, drow, c);
        } else {
            const int r0 = tau >> 4, col = tau & 15;
#pragma unroll
            for (int k = 0; k < 16; ++k) e[k] = A::load_in(row0[((long long)(r0 + 16 * k) << b) + col], F.in_raw);
            if (F.scal) {
                const TW s = scalar_tw<A>(F, limb);
#pragma unroll
                for (int k = 0; k < 16; ++k) e[k] = A::mul(e[k], s, c);
            }
        }
        if constexpr (STAGED) {
            mbar_wait(bar, 0);
            fast_fwd_round<A, 0>(e, TwSharedCol<TW>{tws, 0, 0u}, c);
        } else {
            fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, 0, 0u}, c);
        }
        smx_store(sm, e, tau, 8);
    }
    __syncthreads();
    {
        smx_load(sm, e, tau, 4);
        const int hi = tau >> 4, col = tau & 15;
        if constexpr (STAGED)
            fast_fwd_round<A, 0>(e, TwSharedCol<TW>{tws, 4, (unsigned)hi}, c);
        else
            fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, 4, (unsigned)hi}, c);
#pragma unroll
        for (int k = 0; k < 16; ++k) row0[((long long)(hi * 16 + k) << b) + col] = A::store_mid(e[k], c);
    }
}

template <int DUMMY>
__global__ void __launch_bounds__(NTT_THREADS, FAST_COL_CTAS) fast_fwd_colpass(const FastArgs F) {
    extern __shared__ __align__(16) int64_t sm[];
    if (fast_skip_own(F)) return;
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    const RescaleIn none{};
    if (fast_use_f64(F, rid))
        fast_fwd_col_body<ArithF64, true, false>(F, none, sm, limb, rid.data_row);
    else
        fast_fwd_col_body<ArithU64, false, false>(F, none, sm, limb, rid.data_row);
}

// the tensor stage's column pass: rescale fused into the load (rows = 4 polynomials x L limbs, period L, F.scal = R mod q)
template <int DUMMY>
__global__ void __launch_bounds__(NTT_THREADS, FAST_COL_CTAS) fast_fwd_colpass_rescale(const FastArgs F, const RescaleIn R) {
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    if (fast_use_f64(F, rid))
        fast_fwd_col_body<ArithF64, true, true>(F, R, sm, limb, rid.data_row);
    else
        fast_fwd_col_body<ArithU64, false, true>(F, R, sm, limb, rid.data_row);
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
    const int32_t* own_row0;            // [P] or nullptr: first own target limb of each partition (skipped), < 0 = none
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
    if (own_target(X.own_row0, X.alphas, p, t)) return;   // the key switch takes this row from the tensor product
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
    __shared__ ExtShared S;
    const int p = blockIdx.y;
    const int alpha = X.alphas[p];
    {   // stage the constants of targets [t0, t1) (t1 - t0 <= EXT_MAX_E, checked by the launcher)
        const double* __restrict__ hm = X.Hm[p];
        for (int i = threadIdx.x; i < t1 - t0; i += blockDim.x) {
            const double q = (double)(uint64_t)X.q[t0 + i];
            S.q[i] = q;
            S.qinv[i] = 1.0 / q;
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
    const int32_t *own_row0, *alphas;   // [P] own target rows of each partition (see FastArgs), or nullptr
    const int64_t* d2hat;               // [L][N] NTT-domain d2 = NTT(X) R (lazy integers) used for the own rows
};

__global__ void __launch_bounds__(256) k_ksk_inner_fast(const InnerArgs X) {
    const int t = X.t0 + blockIdx.y;
    const long long j = 2ll * (blockIdx.x * 256 + threadIdx.x);
    if (j >= X.N) return;
    const uint64_t q = (uint64_t)X.q[t];
    longlong2 r0, r1;
    if (q < SMALL_PRIME_LIMIT) {
        const F64C c{(double)q, 1.0 / (double)q};
        double a0x = 0.0, a0y = 0.0, a1x = 0.0, a1y = 0.0;
        for (int p = 0; p < X.P; ++p) {
            const bool own = own_target(X.own_row0, X.alphas, p, t);
            const longlong2 e = own ? *reinterpret_cast<const longlong2*>(X.d2hat + (long long)t * X.N + j)
                                    : *reinterpret_cast<const longlong2*>(X.ext + ((long long)p * X.E + t) * X.N + j);
            const longlong2 u = *reinterpret_cast<const longlong2*>(X.k0[p] + (long long)t * X.k_stride + j);
            const longlong2 v = *reinterpret_cast<const longlong2*>(X.k1[p] + (long long)t * X.k_stride + j);
            double ex, ey;
            if (own) {   // d2hat carries the Montgomery factor; the extended rows of this kernel are plain
                ex = f64_mulmod(i2d(e.x), X.Rinv[t], c);
                ey = f64_mulmod(i2d(e.y), X.Rinv[t], c);
            } else {
                ex = X.raw ? __longlong_as_double(e.x) : i2d(e.x);
                ey = X.raw ? __longlong_as_double(e.y) : i2d(e.y);
            }
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
            const bool own = own_target(X.own_row0, X.alphas, p, t);
            const longlong2 e = own ? *reinterpret_cast<const longlong2*>(X.d2hat + (long long)t * X.N + j)
                                    : *reinterpret_cast<const longlong2*>(X.ext + ((long long)p * X.E + t) * X.N + j);
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
};

__global__ void __launch_bounds__(256) k_moddown_fast(const ModDownArgs X) {
    const int t = blockIdx.y;
    const long long j = 2ll * (blockIdx.x * 256 + threadIdx.x);
    if (j >= X.N) return;
    const uint64_t q = (uint64_t)X.q[t];
    const F64C c{(double)q, 1.0 / (double)q};
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
        const longlong2 a = *reinterpret_cast<const longlong2*>(X.add + (long long)t * X.add_stride + j);
        const int64_t qi = (int64_t)q;
        r.x = reduce_q(lazy_add(a.x, r.x, 2 * qi), qi);
        r.y = reduce_q(lazy_add(a.y, r.y, 2 * qi), qi);
    }
    *reinterpret_cast<longlong2*>(X.out + (long long)t * X.out_stride + j) = r;
}

// ---- forward pass B (block pass, stages 8..logN-1), canonical [0,q) out --------------------------------------
template <class A, int B, bool STAGED>
__device__ __forceinline__ void fast_fwd_block_body(const FastArgs& F, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const unsigned chunk = grid_chunk(F);
    constexpr int logN = B + 8;
    const typename A::C c = make_const<A>((uint64_t)F.q[limb]);
    int64_t* __restrict__ g = F.a + drow * F.a_stride + (long long)chunk * TILE;
    const TW* __restrict__ W = tw_row<A>(F, limb);
    TW* tws = reinterpret_cast<TW*>(sm + SMEM_SLOTS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + SMEM_SLOTS + FAST_TW_SLOTS);
    if constexpr (STAGED) stage_block_twiddles(reinterpret_cast<const double*>(W), reinterpret_cast<double*>(tws), bar, B, chunk);
    if (tau == 32) {
        unsigned ca = 0;
        const long long ra = F.prefetch ? fast_row_ahead(F, F.prefetch, ca) : -1;
        if (ra >= 0) l2_prefetch_bulk(F.a + ra * F.a_stride + (long long)ca * TILE, TILE * 8u);
    }
    T e[16];
    constexpr int P1 = B - 4;
    {
        const int zb = zbase(tau, P1);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = A::load_mid(g[zb | (k << P1)]);
        if constexpr (STAGED) {
            mbar_wait(bar, 0);
            fast_fwd_round<A, 0>(e, TwSharedBlock<TW>{tws, 12 - B, 0, (unsigned)(tau >> P1)}, c);
        } else {
            fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P1, (chunk << (8 - P1)) | (unsigned)(tau >> P1)}, c);
        }
    }
    if constexpr (B >= 8) {
        constexpr int P2 = B - 8;
        smx_store(sm, e, tau, P1);
        __syncthreads();
        smx_load(sm, e, tau, P2);
        if constexpr (STAGED)
            fast_fwd_round<A, 0>(e, TwSharedBlock<TW>{tws, 12 - B, 4, (unsigned)(tau >> P2)}, c);
        else
            fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P2, (chunk << (8 - P2)) | (unsigned)(tau >> P2)}, c);
        if constexpr (B == 9) {
            __syncthreads();
            smx_store(sm, e, tau, P2);
            __syncthreads();
            smx_load(sm, e, tau, 0);
            if constexpr (STAGED)
                fast_fwd_round<A, 3>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
            else
                fast_fwd_round<A, 3>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
        }
        __syncthreads();
    } else if constexpr (B > 4) {
        smx_store(sm, e, tau, P1);
        __syncthreads();
        smx_load(sm, e, tau, 0);
        if constexpr (STAGED)
            fast_fwd_round<A, 8 - B>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
        else
            fast_fwd_round<A, 8 - B>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
        __syncthreads();
    }
    {   // canonical values back through shared memory for coalesced 128-bit stores
        int64_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = A::store_out(e[k], c, F.out_raw);
        sm_store_field(sm, r, tau, 0);
    }
    __syncthreads();
    sm_to_global(sm, g, tau);
}

// ---- forward pass B, WARP-INDEPENDENT form -------------------------------------------------------------------------
// Stages 8..logN-1 only mix coefficients inside 2^B-point sub-blocks (B <= 9), i.e. inside the 512 contiguous
// coefficients a warp owns (32 threads x 16).  Every exchange therefore stays inside the warp's private slice of the
// exchange buffer and needs __syncwarp() only -- no CTA barrier anywhere, so the 8 warps of a CTA drift apart and keep
// the FP64 pipe fed while others wait on loads.  After the last round a thread holds 16 contiguous coefficients,
// which leave as four 256-bit stores (no staging pass through shared memory).
template <class A, int B, bool STAGED, bool HYB = false>
__device__ __forceinline__ void fast_fwd_block_body_w(const FastArgs& F, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const unsigned chunk = grid_chunk(F);
    constexpr int logN = B + 8;
    constexpr bool S2 = STAGED && !HYB;   // rounds after the first take their twiddles from shared memory
    const typename A::C c = make_const<A>((uint64_t)F.q[limb]);
    int64_t* __restrict__ g = F.a + drow * F.a_stride + (long long)chunk * TILE;
    const TW* __restrict__ W = tw_row<A>(F, limb);
    TW* tws = reinterpret_cast<TW*>(sm + SMEM_SLOTS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + SMEM_SLOTS + (HYB ? HYB_TW_SLOTS : FAST_TW_SLOTS));
    if constexpr (STAGED && HYB)
        stage_block_twiddles_first(reinterpret_cast<const double*>(W), reinterpret_cast<double*>(tws), bar, B, chunk, 4);
    else if constexpr (STAGED)
        stage_block_twiddles(reinterpret_cast<const double*>(W), reinterpret_cast<double*>(tws), bar, B, chunk);
    if (tau == 32) {
        unsigned ca = 0;
        const long long ra = F.prefetch ? fast_row_ahead(F, F.prefetch, ca) : -1;
        if (ra >= 0) l2_prefetch_bulk(F.a + ra * F.a_stride + (long long)ca * TILE, TILE * 8u);
    }
    T e[16];
    constexpr int P1 = B - 4;
    {
        const int zb = zbase(tau, P1);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = A::load_mid(g[zb | (k << P1)]);
        if constexpr (STAGED) {
            mbar_wait(bar, 0);
            fast_fwd_round<A, 0>(e, TwSharedBlock<TW>{tws, 12 - B, 0, (unsigned)(tau >> P1)}, c);
        } else {
            fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P1, (chunk << (8 - P1)) | (unsigned)(tau >> P1)}, c);
        }
    }
    if constexpr (B >= 8) {
        constexpr int P2 = B - 8;
        smx_store(sm, e, tau, P1);
        __syncwarp();
        smx_load(sm, e, tau, P2);
        if constexpr (S2)
            fast_fwd_round<A, 0>(e, TwSharedBlock<TW>{tws, 12 - B, 4, (unsigned)(tau >> P2)}, c);
        else
            fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P2, (chunk << (8 - P2)) | (unsigned)(tau >> P2)}, c);
        if constexpr (B == 9) {
            __syncwarp();
            smx_store(sm, e, tau, P2);
            __syncwarp();
            smx_load(sm, e, tau, 0);
            if constexpr (S2)
                fast_fwd_round<A, 3>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
            else
                fast_fwd_round<A, 3>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
        }
    } else if constexpr (B > 4) {
        smx_store(sm, e, tau, P1);
        __syncwarp();
        smx_load(sm, e, tau, 0);
        if constexpr (S2)
            fast_fwd_round<A, 8 - B>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
        else
            fast_fwd_round<A, 8 - B>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
    }
    {   // the thread's 16 contiguous canonical coefficients: four full-sector 256-bit stores
        int64_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = A::store_out(e[k], c, F.out_raw);
        int64_t* o = g + tau * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) stg256(o, r, j);
    }
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, FAST_CTAS_PER_SM) fast_fwd_blockpass_w(const FastArgs F) {
    extern __shared__ __align__(16) int64_t sm[];
    if (fast_skip_own(F)) return;
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    if (fast_use_f64(F, rid))
        fast_fwd_block_body_w<ArithF64, B, true>(F, sm, limb, rid.data_row);
    else
        fast_fwd_block_body_w<ArithU64, B, false>(F, sm, limb, rid.data_row);
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, 4) fast_fwd_blockpass_h(const FastArgs F) {
    extern __shared__ __align__(16) int64_t sm[];
    if (fast_skip_own(F)) return;
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    if (fast_use_f64(F, rid))
        fast_fwd_block_body_w<ArithF64, B, true, true>(F, sm, limb, rid.data_row);
    else
        fast_fwd_block_body_w<ArithU64, B, false, true>(F, sm, limb, rid.data_row);
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, FAST_CTAS_PER_SM) fast_fwd_blockpass(const FastArgs F) {
    extern __shared__ __align__(16) int64_t sm[];
    if (fast_skip_own(F)) return;
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    if (fast_use_f64(F, rid))
        fast_fwd_block_body<ArithF64, B, true>(F, sm, limb, rid.data_row);
    else
        fast_fwd_block_body<ArithU64, B, false>(F, sm, limb, rid.data_row);
}

// ---- PERSISTENT forward block pass: TMA double-buffered tiles ----------------------------------------------------
// One CTA walks a contiguous range of tiles.  While tile i is being transformed out of registers, the 32 KB of
// tile i+1 are already in flight (one cp.async.bulk into the other buffer, completion on an mbarrier), so the
// DRAM/L2 latency that the one-tile-per-CTA kernels expose at every CTA start is hidden behind compute.  Tiles are
// ordered so that consecutive tiles share (limb, chunk): the staged twiddles are reused by all G rows of a batch
// that belong to the same limb (the P partitions of a key switch, the 4 polynomials of a tensor stage, ...).
struct TileId {
    long long drow;
    int limb, chunk, group;
};
__device__ __forceinline__ TileId decode_tile(const FastArgs& F, long long id, int G, int chunks) {
    const int group = (int)(id / G), g = (int)(id - (long long)group * G);
    const int m = group / chunks, chunk = group - m * chunks;
    TileId t;
    t.group = group;
    t.chunk = chunk;
    if (F.slab_rows == 0) {
        t.limb = m;
        t.drow = (long long)g * F.period + m;
    } else {
        t.limb = F.slab_t0 + m;
        t.drow = (long long)g * F.group_rows + F.slab_t0 + m;
    }
    return t;
}
__device__ __forceinline__ TileId decode_tile32(const FastArgs& F, int id, int G, int chunks) {   // 32-bit index maths
    const int group = id / G, g = id - group * G;
    const int m = group / chunks, chunk = group - m * chunks;
    TileId t;
    t.group = group;
    t.chunk = chunk;
    if (F.slab_rows == 0) {
        t.limb = m;
        t.drow = (long long)g * F.period + m;
    } else {
        t.limb = F.slab_t0 + m;
        t.drow = (long long)g * F.group_rows + F.slab_t0 + m;
    }
    return t;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// persistent kernel shared memory:
// [exchange buffer (padded tile) | PERSIST_SLOTS raw tile slots | 2 twiddle stages | mbarriers]
constexpr int PERSIST_SLOTS = 3;
constexpr int PERSIST_SMEM_BYTES = SMEM_BYTES + PERSIST_SLOTS * TILE * 8 + 2 * FAST_TW_SLOTS * 8 + 64;

template <class A, int B, bool STAGED>
__device__ __forceinline__ void persist_fwd_block_compute(const FastArgs& F, typename A::T (&e)[16], int64_t* xb,
                                                          const typename A::TW* tws, const TileId& tl,
                                                          int64_t* __restrict__ gout) {
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    constexpr int logN = B + 8;
    constexpr int P1 = B - 4;
    const unsigned chunk = (unsigned)tl.chunk;
    const typename A::C c = make_const<A>((uint64_t)F.q[tl.limb]);
    const TW* __restrict__ W = tw_row<A>(F, tl.limb);
    if constexpr (STAGED)
        fast_fwd_round<A, 0>(e, TwSharedBlock<TW>{tws, 12 - B, 0, (unsigned)(tau >> P1)}, c);
    else
        fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P1, (chunk << (8 - P1)) | (unsigned)(tau >> P1)}, c);
    if constexpr (B >= 8) {
        constexpr int P2 = B - 8;
        smx_store(xb, e, tau, P1);
        __syncthreads();
        smx_load(xb, e, tau, P2);
        if constexpr (STAGED)
            fast_fwd_round<A, 0>(e, TwSharedBlock<TW>{tws, 12 - B, 4, (unsigned)(tau >> P2)}, c);
        else
            fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P2, (chunk << (8 - P2)) | (unsigned)(tau >> P2)}, c);
        if constexpr (B == 9) {
            __syncthreads();
            smx_store(xb, e, tau, P2);
            __syncthreads();
            smx_load(xb, e, tau, 0);
            if constexpr (STAGED)
                fast_fwd_round<A, 3>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
            else
                fast_fwd_round<A, 3>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
        }
        __syncthreads();
    } else if constexpr (B > 4) {
        smx_store(xb, e, tau, P1);
        __syncthreads();
        smx_load(xb, e, tau, 0);
        if constexpr (STAGED)
            fast_fwd_round<A, 8 - B>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
        else
            fast_fwd_round<A, 8 - B>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
        __syncthreads();
    }
    {
        int64_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = A::store_out(e[k], c, F.out_raw);
        sm_store_field(xb, r, tau, 0);
    }
    __syncthreads();
    sm_to_global(xb, gout, tau);
}

// One CTA per SM.  A ring of PERSIST_SLOTS raw tiles keeps the next tiles' loads (TMA bulk copies) in flight while
// the current tile is transformed, and the twiddles of the NEXT (limb, chunk) group are staged into the second
// twiddle buffer while the current group is processed: after the prologue no wait on L2/DRAM is exposed.
template <int B>
__global__ void __launch_bounds__(NTT_THREADS, 1) fast_fwd_blockpass_persist(const FastArgs F, long long total_tiles, int G) {
    extern __shared__ __align__(16) int64_t sm[];
    int64_t* xb = sm;
    int64_t* ring = sm + SMEM_SLOTS;
    double* tws = reinterpret_cast<double*>(ring + PERSIST_SLOTS * TILE);      // [2][FAST_TW_SLOTS]
    uint64_t* bar_data = reinterpret_cast<uint64_t*>(tws + 2 * FAST_TW_SLOTS);  // [PERSIST_SLOTS]
    uint64_t* bar_tw = bar_data + PERSIST_SLOTS;                               // [2]
    const int tau = threadIdx.x;
    constexpr int P1 = B - 4;
    const int chunks = (1 << (B + 8)) / TILE;
    const long long t_begin = total_tiles * blockIdx.x / gridDim.x;
    const long long t_end = total_tiles * (blockIdx.x + 1) / gridDim.x;
    if (t_begin >= t_end) return;
    if (tau == 0) {
        for (int i = 0; i < PERSIST_SLOTS; ++i) mbar_init(&bar_data[i], 1);
        mbar_init(&bar_tw[0], 1);
        mbar_init(&bar_tw[1], 1);
    }
    __syncthreads();
    auto issue_data = [&](long long id, int slot) {   // thread 0 only
        const TileId t = decode_tile(F, id, G, chunks);
        const int64_t* src = F.a + t.drow * F.a_stride + (long long)t.chunk * TILE;
        fence_async_smem();
        mbar_expect_tx(&bar_data[slot], TILE * 8u);
        tma_bulk_g2s(ring + slot * TILE, src, TILE * 8u, &bar_data[slot]);
    };
    auto group_staged = [&](int limb) { return (uint64_t)F.q[limb] < SMALL_PRIME_LIMIT && F.force_int != 1; };
    auto issue_tw = [&](long long id, int tslot) {    // thread 0 only; id = first tile of the group
        const TileId t = decode_tile(F, id, G, chunks);
        if (!group_staged(t.limb)) return;
        const double* W = F.tw_f64 + ((long long)t.limb << (B + 8));
        const int unit_log = 12 - B;
        double* dst = tws + tslot * FAST_TW_SLOTS;
        fence_async_smem();
        mbar_expect_tx(&bar_tw[tslot], (unsigned)(((1u << 12) - (1u << unit_log)) * 8u));
        for (int j = 0; j < B; ++j) {
            const unsigned cnt = 1u << (j + unit_log);
            tma_bulk_g2s(dst + (((1u << j) - 1u) << unit_log), W + (1u << (8 + j)) + (size_t)t.chunk * cnt, cnt * 8u, &bar_tw[tslot]);
        }
    };
    if (tau == 0) {
        issue_tw(t_begin, 0);
        for (int i = 0; i < PERSIST_SLOTS && t_begin + i < t_end; ++i) issue_data(t_begin + i, i);
    }
    long long cur_group = -1;
    unsigned gcount = 0;            // groups seen by this CTA; group k stages into buffer k & 1
    unsigned tw_w0 = 0, tw_w1 = 0;  // completed phases per twiddle buffer
    unsigned it = 0;
    for (int id = t_begin; id < t_end; ++id, ++it) {
        const int slot = it % PERSIST_SLOTS;
        const TileId tl = decode_tile(F, id, G, chunks);
        const RowId rid{tl.drow, tl.limb};
        const bool f64 = fast_use_f64(F, rid);
        bool new_group = false;
        if (tl.group != cur_group) {   // readers of buffer (gcount+1)&1 (two groups back) left at the loop-end barrier
            cur_group = tl.group;
            new_group = true;
            const int next_first = (tl.group + 1) * G;
            if (tau == 0 && next_first < t_end) issue_tw(next_first, (gcount + 1) & 1);
            ++gcount;
        }
        const int tslot = (gcount - 1) & 1;
        mbar_wait(&bar_data[slot], (it / PERSIST_SLOTS) & 1);
        int64_t* gout = F.a + tl.drow * F.a_stride + (long long)tl.chunk * TILE;
        const int64_t* raw = ring + slot * TILE;
        const int zb = zbase(tau, P1);
        if (new_group && group_staged(tl.limb)) {
            mbar_wait(&bar_tw[tslot], (tslot ? tw_w1 : tw_w0) & 1);
            if (tslot) ++tw_w1; else ++tw_w0;
        }
        if (f64) {
            double e[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) e[k] = ArithF64::load_mid(raw[zb | (k << P1)]);
            __syncthreads();   // slot consumed
            if (tau == 0 && id + PERSIST_SLOTS < t_end) issue_data(id + PERSIST_SLOTS, slot);
            persist_fwd_block_compute<ArithF64, B, true>(F, e, xb, tws + tslot * FAST_TW_SLOTS, tl, gout);
        } else {
            uint64_t e[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) e[k] = ArithU64::load_mid(raw[zb | (k << P1)]);
            __syncthreads();
            if (tau == 0 && id + PERSIST_SLOTS < t_end) issue_data(id + PERSIST_SLOTS, slot);
            persist_fwd_block_compute<ArithU64, B, false>(F, e, xb, nullptr, tl, gout);
        }
        __syncthreads();   // exchange buffer free; on a group change the older twiddle buffer is free
    }
}

// ---- PERSISTENT, SOFTWARE-PIPELINED forward block pass ("pp") ---------------------------------------------------------
// The one-tile-per-CTA kernels run load -> compute -> store back to back in every CTA, and the CTAs of an SM fall
// into step: the memory system is idle while they compute and the FP64 pipe is idle while they load (ncu: FP64 pipe
// 42 % busy, no dominant stall).  Here a CTA walks a contiguous range of tiles and every thread PREFETCHES THE NEXT
// TILE'S 16 COEFFICIENTS INTO REGISTERS before it transforms the current one, so the load latency of tile i+1 hides
// behind the butterflies of tile i.  Warps are independent inside a tile (see fast_fwd_block_body_w): the only CTA
// barrier is at a change of (limb, chunk) group, where the twiddle buffer of the group before last is handed back
// to the TMA engine (twiddles are double-buffered and staged one group ahead).  Tiles are ordered so that the G
// rows sharing a limb (the partitions of a key switch) are consecutive and reuse the staged twiddles.
// Static split of the (limb-major) tile list over the persistent CTAs, by COST rather than by count: a tile of a
// 60-bit limb (integer Shoup path) takes about PP_INT_WEIGHT times as long as an FP64 tile, and those limbs sit at
// the end of the list -- an even split by count leaves the last CTAs with integer tiles only (measured: 35 % idle).
constexpr int PP_INT_WEIGHT = 3;
struct TileRange {
    int begin, end;
};
__device__ __forceinline__ TileRange pp_tile_range(const FastArgs& F, int nlimbs, int tiles_per_limb) {
    const int l0 = F.slab_rows ? F.slab_t0 : 0;
    auto weight = [&](int l) { return ((uint64_t)F.q[l0 + l] < SMALL_PRIME_LIMIT && F.force_int != 1) ? 1 : PP_INT_WEIGHT; };
    long long W = 0;
    for (int l = 0; l < nlimbs; ++l) W += (long long)weight(l) * tiles_per_limb;
    const long long lo = W * blockIdx.x / gridDim.x, hi = W * (blockIdx.x + 1) / gridDim.x;
    auto to_tile = [&](long long x) {
        long long acc = 0;
        int tiles = 0;
        for (int l = 0; l < nlimbs; ++l) {
            const int w = weight(l);
            const long long seg = (long long)w * tiles_per_limb;
            if (x < acc + seg) return tiles + (int)((x - acc) / w);
            acc += seg;
            tiles += tiles_per_limb;
        }
        return tiles;
    };
    return TileRange{to_tile(lo), to_tile(hi)};
}
constexpr int PP_SMEM_BYTES = SMEM_BYTES + 2 * FAST_TW_SLOTS * 8 + 32;

template <class A, int B, bool STAGED>
__device__ __forceinline__ void pp_fwd_block_tile(const FastArgs& F, const int64_t (&raw)[16], int64_t* xb,
                                                  const typename A::TW* tws, const TileId& tl, int64_t* __restrict__ g) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    constexpr int logN = B + 8;
    constexpr int P1 = B - 4;
    const unsigned chunk = (unsigned)tl.chunk;
    const typename A::C c = make_const<A>((uint64_t)F.q[tl.limb]);
    const TW* __restrict__ W = tw_row<A>(F, tl.limb);
    T e[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = A::load_mid(raw[k]);
    if constexpr (STAGED)
        fast_fwd_round<A, 0>(e, TwSharedBlock<TW>{tws, 12 - B, 0, (unsigned)(tau >> P1)}, c);
    else
        fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P1, (chunk << (8 - P1)) | (unsigned)(tau >> P1)}, c);
    if constexpr (B >= 8) {
        constexpr int P2 = B - 8;
        __syncwarp();   // the warp's slice of the exchange buffer is free (previous tile fully read)
        smx_store(xb, e, tau, P1);
        __syncwarp();
        smx_load(xb, e, tau, P2);
        if constexpr (STAGED)
            fast_fwd_round<A, 0>(e, TwSharedBlock<TW>{tws, 12 - B, 4, (unsigned)(tau >> P2)}, c);
        else
            fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, logN - 4 - P2, (chunk << (8 - P2)) | (unsigned)(tau >> P2)}, c);
        if constexpr (B == 9) {
            __syncwarp();
            smx_store(xb, e, tau, P2);
            __syncwarp();
            smx_load(xb, e, tau, 0);
            if constexpr (STAGED)
                fast_fwd_round<A, 3>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
            else
                fast_fwd_round<A, 3>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
        }
    } else if constexpr (B > 4) {
        __syncwarp();
        smx_store(xb, e, tau, P1);
        __syncwarp();
        smx_load(xb, e, tau, 0);
        if constexpr (STAGED)
            fast_fwd_round<A, 8 - B>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
        else
            fast_fwd_round<A, 8 - B>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
    }
    int64_t r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = A::store_out(e[k], c, F.out_raw);
    int64_t* o = g + tau * 16;
#pragma unroll
    for (int j = 0; j < 4; ++j) stg256(o, r, j);
}

// thread 0: stage the block-pass twiddles of (limb, chunk) into `dst` (FP64 rows only), completion on `bar`
template <int B>
__device__ __forceinline__ void pp_issue_block_twiddles(const FastArgs& F, int limb, int chunk, double* dst, uint64_t* bar) {
    const double* W = F.tw_f64 + ((long long)limb << (B + 8));
    constexpr int unit_log = 12 - B;
    fence_async_smem();
    mbar_expect_tx(bar, (unsigned)(((1u << 12) - (1u << unit_log)) * 8u));
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const unsigned cnt = 1u << (j + unit_log);
        tma_bulk_g2s(dst + (((1u << j) - 1u) << unit_log), W + (1u << (8 + j)) + (size_t)chunk * cnt, cnt * 8u, bar);
    }
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, 2) fast_fwd_blockpass_pp(const FastArgs F, long long total_tiles, int G) {
    extern __shared__ __align__(16) int64_t sm[];
    int64_t* xb = sm;
    double* tws = reinterpret_cast<double*>(sm + SMEM_SLOTS);                   // [2][FAST_TW_SLOTS]
    uint64_t* bar_tw = reinterpret_cast<uint64_t*>(tws + 2 * FAST_TW_SLOTS);    // [2]
    const int tau = threadIdx.x;
    constexpr int P1 = B - 4;
    const int chunks = (1 << (B + 8)) / TILE;
    const TileRange range = pp_tile_range(F, (int)(total_tiles / ((long long)chunks * G)), chunks * G);
    const int t_begin = range.begin, t_end = range.end;
    if (t_begin >= t_end) return;
    auto staged = [&](int limb) { return (uint64_t)F.q[limb] < SMALL_PRIME_LIMIT && F.force_int != 1; };
    if (tau == 0) {
        mbar_init(&bar_tw[0], 1);
        mbar_init(&bar_tw[1], 1);
    }
    __syncthreads();
    const int zb = zbase(tau, P1);
    int64_t nxt[16];
    {
        const TileId t0 = decode_tile32(F, t_begin, G, chunks);
        if (tau == 0 && staged(t0.limb)) pp_issue_block_twiddles<B>(F, t0.limb, t0.chunk, tws, &bar_tw[0]);
        const int64_t* __restrict__ src = F.a + t0.drow * F.a_stride + (long long)t0.chunk * TILE;
#pragma unroll
        for (int k = 0; k < 16; ++k) nxt[k] = src[zb | (k << P1)];
    }
    int cur_group = -1;
    unsigned gcount = 0;                 // groups seen by this CTA; group k uses twiddle buffer k & 1
    unsigned ph0 = 0, ph1 = 0;           // completed phases of the two twiddle barriers
#pragma unroll 1
    for (int id = t_begin; id < t_end; ++id) {
        const TileId tl = decode_tile32(F, id, G, chunks);
        int64_t cur[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) cur[k] = nxt[k];
        if (id + 1 < t_end) {            // prefetch the next tile while this one is transformed
            const TileId tn = decode_tile32(F, id + 1, G, chunks);
            const int64_t* __restrict__ src = F.a + tn.drow * F.a_stride + (long long)tn.chunk * TILE;
#pragma unroll
            for (int k = 0; k < 16; ++k) nxt[k] = src[zb | (k << P1)];
        }
        if (tl.group != cur_group) {
            cur_group = tl.group;
            __syncthreads();             // every warp is done with the group before: its twiddle buffer can be refilled
            const int next_first = (tl.group + 1) * G;
            if (tau == 0 && next_first < t_end) {
                const TileId tg = decode_tile32(F, next_first, G, chunks);
                if (staged(tg.limb)) pp_issue_block_twiddles<B>(F, tg.limb, tg.chunk, tws + ((gcount + 1) & 1) * FAST_TW_SLOTS, &bar_tw[(gcount + 1) & 1]);
            }
            if (staged(tl.limb)) {
                const unsigned b = gcount & 1;
                mbar_wait(&bar_tw[b], (b ? ph1 : ph0) & 1);
                if (b) ++ph1; else ++ph0;
            }
            ++gcount;
        }
        const unsigned tb = (gcount - 1) & 1;
        int64_t* gout = F.a + tl.drow * F.a_stride + (long long)tl.chunk * TILE;
        const RowId rid{tl.drow, tl.limb};
        if (fast_use_f64(F, rid))
            pp_fwd_block_tile<ArithF64, B, true>(F, cur, xb, tws + tb * FAST_TW_SLOTS, tl, gout);
        else
            pp_fwd_block_tile<ArithU64, B, false>(F, cur, xb, nullptr, tl, gout);
    }
}

// ---- inverse pass B' (levels 0..B-1) ---------------------------------------------------------------------------
template <class A, int B, bool STAGED>
__device__ __forceinline__ void fast_inv_block_body(const FastArgs& F, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const unsigned chunk = grid_chunk(F);
    constexpr int logN = B + 8;
    const typename A::C c = make_const<A>((uint64_t)F.q[limb]);
    int64_t* __restrict__ g = F.a + drow * F.a_stride + (long long)chunk * TILE;
    const TW* __restrict__ W = tw_row<A>(F, limb);
    TW* tws = reinterpret_cast<TW*>(sm + SMEM_SLOTS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + SMEM_SLOTS + FAST_TW_SLOTS);
    if constexpr (STAGED) stage_block_twiddles(reinterpret_cast<const double*>(W), reinterpret_cast<double*>(tws), bar, B, chunk);
    if (tau == 32) {
        unsigned ca = 0;
        const long long ra = F.prefetch ? fast_row_ahead(F, F.prefetch, ca) : -1;
        if (ra >= 0) l2_prefetch_bulk(F.a + ra * F.a_stride + (long long)ca * TILE, TILE * 8u);
    }
    T e[16];
    global_to_sm(sm, g, tau);
    __syncthreads();
    {
        int64_t r[16];
        sm_load_field(sm, r, tau, 0);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = A::load_in(r[k], F.in_raw);
    }
    if constexpr (STAGED) {
        mbar_wait(bar, 0);
        fast_inv_round<A, 4>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
    } else {
        fast_inv_round<A, 4>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
    }
    if constexpr (B == 4) {
        int64_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = A::store_mid(e[k], c);
        __syncthreads();
        sm_store_field(sm, r, tau, 0);
        __syncthreads();
        sm_to_global(sm, g, tau);
    } else {
        __syncthreads();
        smx_store(sm, e, tau, 0);
        __syncthreads();
        smx_load(sm, e, tau, 4);
        constexpr int NST = (B >= 8) ? 4 : B - 4;
        if constexpr (STAGED)
            fast_inv_round<A, NST>(e, TwSharedBlock<TW>{tws, 12 - B, B - 8, (unsigned)(tau >> 4)}, c);
        else
            fast_inv_round<A, NST>(e, TwGlobal<TW>{W, logN - 8, (chunk << 4) | (unsigned)(tau >> 4)}, c);
        if constexpr (B == 9) {
            __syncthreads();
            smx_store(sm, e, tau, 4);
            __syncthreads();
            smx_load(sm, e, tau, 8);
            if constexpr (STAGED)
                fast_inv_round<A, 1>(e, TwSharedBlock<TW>{tws, 12 - B, B - 12, 0u}, c);
            else
                fast_inv_round<A, 1>(e, TwGlobal<TW>{W, logN - 12, chunk}, c);
            const int zb = zbase(tau, 8);
#pragma unroll
            for (int k = 0; k < 16; ++k) g[zb | (k << 8)] = A::store_mid(e[k], c);
        } else {
            const int zb = zbase(tau, 4);
#pragma unroll
            for (int k = 0; k < 16; ++k) g[zb | (k << 4)] = A::store_mid(e[k], c);
        }
    }
}

// ---- tensor product of cc_mult (engine.py:1095-1101) fused into the load of the inverse block pass ---------------------
// Row r of the [3L] batch is d_(r / L), limb r % L:  d0 = x0 y0,  d1 = x0 y1 + x1 y0,  d2 = x1 y1, where x[4][L][N] holds
// the NTT-domain polynomials x0, x1, y0, y1 (Montgomery form: transforms of x R).  Scale-prime rows multiply in FP64
// (x holds raw doubles there; the product carries R^2, removed by the exit scalar N^-1 R^-2), 60-bit rows use the
// reference's Montgomery product (exit scalar N^-1 R^-1).  If d2hat != nullptr the NTT-domain d2 is also kept
// (row-major [L][N]) for the key switch, which then skips the transform of every partition's own limbs.
struct TensorIn {
    const int64_t* x;          // [4][L][N]
    long long poly_stride;     // L * N
    const int64_t *_2q, *ql, *qh, *kl, *kh;
    int64_t* d2hat;            // optional [L][N]
    int L;
};
__device__ __forceinline__ void ldg256v(const int64_t* p, int64_t (&r)[4]) {
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3]) : "l"(p));
}
__device__ __forceinline__ void stg256v(int64_t* p, const int64_t (&r)[4]) {
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(r[0]), "l"(r[1]), "l"(r[2]), "l"(r[3]) : "memory");
}
template <class A>
__device__ __forceinline__ void tensor_load(const FastArgs& F, const TensorIn& Tn, typename A::T (&e)[16], int limb,
                                            long long drow, long long off, const typename A::C& c);
template <>
__device__ __forceinline__ void tensor_load<ArithF64>(const FastArgs& F, const TensorIn& Tn, double (&e)[16], int limb,
                                                      long long drow, long long off, const F64C& c) {
    const int g = (int)(drow / Tn.L);
    const int64_t* base = Tn.x + (long long)limb * F.a_stride + off;
    const int64_t* pa = base + ((g == 2) ? 1 : 0) * Tn.poly_stride;
    const int64_t* pb = base + ((g == 0) ? 2 : 3) * Tn.poly_stride;
#pragma unroll
    for (int j = 0; j < 4; ++j) {      // four coefficients at a time keeps the live operand registers low
        int64_t a[4], b[4];
        ldg256v(pa + 4 * j, a);
        ldg256v(pb + 4 * j, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) e[4 * j + i] = f64_mulmod(__longlong_as_double(a[i]), __longlong_as_double(b[i]), c);
    }
    if (g == 1) {
        const int64_t* pc = base + 1 * Tn.poly_stride;
        const int64_t* pd = base + 2 * Tn.poly_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t a[4], b[4];
            ldg256v(pc + 4 * j, a);
            ldg256v(pd + 4 * j, b);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                e[4 * j + i] = __dadd_rn(e[4 * j + i], f64_mulmod(__longlong_as_double(a[i]), __longlong_as_double(b[i]), c));
        }
    }
    if (g == 2 && Tn.d2hat) {
        int64_t* o = Tn.d2hat + (long long)limb * F.a_stride + off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = (int64_t)__double_as_longlong(e[4 * j + i]);
            stg256v(o + 4 * j, r);
        }
    }
}
template <>
__device__ __forceinline__ void tensor_load<ArithU64>(const FastArgs& F, const TensorIn& Tn, uint64_t (&e)[16], int limb,
                                                      long long drow, long long off, const U64C& c) {
    const int g = (int)(drow / Tn.L);
    const int64_t* base = Tn.x + (long long)limb * F.a_stride + off;
    const int64_t* pa = base + ((g == 2) ? 1 : 0) * Tn.poly_stride;
    const int64_t* pb = base + ((g == 0) ? 2 : 3) * Tn.poly_stride;
    const LimbConst k = load_limb_const(Tn._2q, Tn.ql, Tn.qh, Tn.kl, Tn.kh, limb);
    const int64_t q2 = (int64_t)k.q2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int64_t a[4], b[4];
        ldg256v(pa + 4 * j, a);
        ldg256v(pb + 4 * j, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) e[4 * j + i] = (uint64_t)mont_mul_ss(a[i], b[i], k.q4, k.k);   // lazy, in [0, 2q)
    }
    if (g == 1) {
        const int64_t* pc = base + 1 * Tn.poly_stride;
        const int64_t* pd = base + 2 * Tn.poly_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t a[4], b[4];
            ldg256v(pc + 4 * j, a);
            ldg256v(pd + 4 * j, b);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                e[4 * j + i] = (uint64_t)lazy_add((int64_t)e[4 * j + i], mont_mul_ss(a[i], b[i], k.q4, k.k), q2);
        }
    }
    if (g == 2 && Tn.d2hat) {
        int64_t* o = Tn.d2hat + (long long)limb * F.a_stride + off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = (int64_t)e[4 * j + i];
            stg256v(o + 4 * j, r);
        }
    }
}

// ---- inverse pass B', WARP-INDEPENDENT form (see fast_fwd_block_body_w) ------------------------------------------------
// A thread starts from its 16 contiguous coefficients (four 256-bit loads), exchanges stay inside the warp; for
// B == 9 the last level (distance 256) is taken as the top stage of field [8:5], which keeps it warp-private too.
template <class A, int B, bool STAGED, bool TENS, bool HYB = false>
__device__ __forceinline__ void fast_inv_block_body_w(const FastArgs& F, const TensorIn& Tn, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const unsigned chunk = grid_chunk(F);
    constexpr int logN = B + 8;
    const typename A::C c = make_const<A>((uint64_t)F.q[limb]);
    int64_t* __restrict__ g = F.a + drow * F.a_stride + (long long)chunk * TILE;
    const TW* __restrict__ W = tw_row<A>(F, limb);
    TW* tws = reinterpret_cast<TW*>(sm + SMEM_SLOTS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + SMEM_SLOTS + (HYB ? HYB_TW_SLOTS : FAST_TW_SLOTS));
    constexpr bool S1 = STAGED && !HYB;   // the first round's twiddles (one thread each) come from shared memory
    if constexpr (STAGED && HYB)
        stage_block_twiddles_first(reinterpret_cast<const double*>(W), reinterpret_cast<double*>(tws), bar, B, chunk, B - 4);
    else if constexpr (STAGED)
        stage_block_twiddles(reinterpret_cast<const double*>(W), reinterpret_cast<double*>(tws), bar, B, chunk);
    if (tau == 32) {
        unsigned ca = 0;
        const long long ra = F.prefetch ? fast_row_ahead(F, F.prefetch, ca) : -1;
        if (ra >= 0) l2_prefetch_bulk(F.a + ra * F.a_stride + (long long)ca * TILE, TILE * 8u);
    }
    T e[16];
    if constexpr (TENS) {
        tensor_load<A>(F, Tn, e, limb, drow, (long long)chunk * TILE + tau * 16, c);
    } else {
        int64_t r[16];
        const int64_t* in = g + tau * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) ldg256(in, r, j);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = A::load_in(r[k], F.in_raw);
    }
    if constexpr (S1) {
        mbar_wait(bar, 0);
        fast_inv_round<A, 4>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
    } else {
        fast_inv_round<A, 4>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
        if constexpr (STAGED) mbar_wait(bar, 0);   // hybrid: the staged (shared) stages are needed from the next round on
    }
    if constexpr (B == 4) {
        int64_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = A::store_mid(e[k], c);
        int64_t* o = g + tau * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) stg256(o, r, j);
    } else {
        smx_store(sm, e, tau, 0);
        __syncwarp();
        smx_load(sm, e, tau, 4);
        constexpr int NST = (B >= 8) ? 4 : B - 4;
        if constexpr (STAGED)
            fast_inv_round<A, NST>(e, TwSharedBlock<TW>{tws, 12 - B, B - 8, (unsigned)(tau >> 4)}, c);
        else
            fast_inv_round<A, NST>(e, TwGlobal<TW>{W, logN - 8, (chunk << 4) | (unsigned)(tau >> 4)}, c);
        if constexpr (B == 9) {
            __syncwarp();
            smx_store(sm, e, tau, 4);
            __syncwarp();
            smx_load(sm, e, tau, 5);
            {   // level with distance 2^8 = top stage of field [8:5]: pairs (k, k+8), one twiddle per 512-point sub-block
                TW w[8];
                const unsigned sub = (unsigned)(tau >> 5);
                if constexpr (STAGED)
                    w[0] = tws[sub];
                else
                    w[0] = __ldg(W + (1u << (logN - 9)) + ((chunk << 3) | sub));
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
__global__ void __launch_bounds__(NTT_THREADS, FAST_CTAS_PER_SM) fast_inv_blockpass_w(const FastArgs F) {
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    const TensorIn none{};
    if (fast_use_f64(F, rid))
        fast_inv_block_body_w<ArithF64, B, true, false>(F, none, sm, limb, rid.data_row);
    else
        fast_inv_block_body_w<ArithU64, B, false, false>(F, none, sm, limb, rid.data_row);
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, 4) fast_inv_blockpass_h(const FastArgs F) {
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    const TensorIn none{};
    if (fast_use_f64(F, rid))
        fast_inv_block_body_w<ArithF64, B, true, false, true>(F, none, sm, limb, rid.data_row);
    else
        fast_inv_block_body_w<ArithU64, B, false, false, true>(F, none, sm, limb, rid.data_row);
}

// the tensor stage's inverse block pass: tensor product fused into the load (rows = 3 x L, period L)
template <int B>
__global__ void __launch_bounds__(NTT_THREADS, FAST_CTAS_PER_SM) fast_inv_blockpass_tensor(const FastArgs F, const TensorIn Tn) {
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    if (fast_use_f64(F, rid))
        fast_inv_block_body_w<ArithF64, B, true, true>(F, Tn, sm, limb, rid.data_row);
    else
        fast_inv_block_body_w<ArithU64, B, false, true>(F, Tn, sm, limb, rid.data_row);
}

template <int B>
__global__ void __launch_bounds__(NTT_THREADS, FAST_CTAS_PER_SM) fast_inv_blockpass(const FastArgs F) {
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    if (fast_use_f64(F, rid))
        fast_inv_block_body<ArithF64, B, true>(F, sm, limb, rid.data_row);
    else
        fast_inv_block_body<ArithU64, B, false>(F, sm, limb, rid.data_row);
}

// ---- PERSISTENT, SOFTWARE-PIPELINED inverse block pass (see fast_fwd_blockpass_pp) --------------------------------------
template <class A, int B, bool STAGED>
__device__ __forceinline__ void pp_inv_block_tile(const FastArgs& F, const int64_t (&raw)[16], int64_t* xb,
                                                  const typename A::TW* tws, const TileId& tl, int64_t* __restrict__ g) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    constexpr int logN = B + 8;
    const unsigned chunk = (unsigned)tl.chunk;
    const typename A::C c = make_const<A>((uint64_t)F.q[tl.limb]);
    const TW* __restrict__ W = tw_row<A>(F, tl.limb);
    T e[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = A::load_in(raw[k], F.in_raw);
    if constexpr (STAGED)
        fast_inv_round<A, 4>(e, TwSharedBlock<TW>{tws, 12 - B, B - 4, (unsigned)tau}, c);
    else
        fast_inv_round<A, 4>(e, TwGlobal<TW>{W, logN - 4, (chunk << 8) | (unsigned)tau}, c);
    if constexpr (B == 4) {
        int64_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = A::store_mid(e[k], c);
        int64_t* o = g + tau * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) stg256(o, r, j);
    } else {
        __syncwarp();   // the warp's slice of the exchange buffer is free (previous tile fully read)
        smx_store(xb, e, tau, 0);
        __syncwarp();
        smx_load(xb, e, tau, 4);
        constexpr int NST = (B >= 8) ? 4 : B - 4;
        if constexpr (STAGED)
            fast_inv_round<A, NST>(e, TwSharedBlock<TW>{tws, 12 - B, B - 8, (unsigned)(tau >> 4)}, c);
        else
            fast_inv_round<A, NST>(e, TwGlobal<TW>{W, logN - 8, (chunk << 4) | (unsigned)(tau >> 4)}, c);
        if constexpr (B == 9) {
            __syncwarp();
            smx_store(xb, e, tau, 4);
            __syncwarp();
            smx_load(xb, e, tau, 5);
            {
                TW w[8];
                const unsigned sub = (unsigned)(tau >> 5);
                if constexpr (STAGED)
                    w[0] = tws[sub];
                else
                    w[0] = __ldg(W + (1u << (logN - 9)) + ((chunk << 3) | sub));
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
__global__ void __launch_bounds__(NTT_THREADS, 2) fast_inv_blockpass_pp(const FastArgs F, long long total_tiles, int G) {
    extern __shared__ __align__(16) int64_t sm[];
    int64_t* xb = sm;
    double* tws = reinterpret_cast<double*>(sm + SMEM_SLOTS);                   // [2][FAST_TW_SLOTS]
    uint64_t* bar_tw = reinterpret_cast<uint64_t*>(tws + 2 * FAST_TW_SLOTS);    // [2]
    const int tau = threadIdx.x;
    const int chunks = (1 << (B + 8)) / TILE;
    const TileRange range = pp_tile_range(F, (int)(total_tiles / ((long long)chunks * G)), chunks * G);
    const int t_begin = range.begin, t_end = range.end;
    if (t_begin >= t_end) return;
    auto staged = [&](int limb) { return (uint64_t)F.q[limb] < SMALL_PRIME_LIMIT && F.force_int != 1; };
    if (tau == 0) {
        mbar_init(&bar_tw[0], 1);
        mbar_init(&bar_tw[1], 1);
    }
    __syncthreads();
    int64_t nxt[16];
    {
        const TileId t0 = decode_tile32(F, t_begin, G, chunks);
        if (tau == 0 && staged(t0.limb)) pp_issue_block_twiddles<B>(F, t0.limb, t0.chunk, tws, &bar_tw[0]);
        const int64_t* src = F.a + t0.drow * F.a_stride + (long long)t0.chunk * TILE + tau * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) ldg256(src, nxt, j);
    }
    long long cur_group = -1;
    unsigned gcount = 0;
    unsigned ph0 = 0, ph1 = 0;
#pragma unroll 1
    for (int id = t_begin; id < t_end; ++id) {
        const TileId tl = decode_tile32(F, id, G, chunks);
        int64_t cur[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) cur[k] = nxt[k];
        if (id + 1 < t_end) {
            const TileId tn = decode_tile32(F, id + 1, G, chunks);
            const int64_t* src = F.a + tn.drow * F.a_stride + (long long)tn.chunk * TILE + tau * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j) ldg256(src, nxt, j);
        }
        if (tl.group != cur_group) {
            cur_group = tl.group;
            __syncthreads();
            const int next_first = (tl.group + 1) * G;
            if (tau == 0 && next_first < t_end) {
                const TileId tg = decode_tile32(F, next_first, G, chunks);
                if (staged(tg.limb)) pp_issue_block_twiddles<B>(F, tg.limb, tg.chunk, tws + ((gcount + 1) & 1) * FAST_TW_SLOTS, &bar_tw[(gcount + 1) & 1]);
            }
            if (staged(tl.limb)) {
                const unsigned b = gcount & 1;
                mbar_wait(&bar_tw[b], (b ? ph1 : ph0) & 1);
                if (b) ++ph1; else ++ph0;
            }
            ++gcount;
        }
        const unsigned tb = (gcount - 1) & 1;
        int64_t* gout = F.a + tl.drow * F.a_stride + (long long)tl.chunk * TILE;
        const RowId rid{tl.drow, tl.limb};
        if (fast_use_f64(F, rid))
            pp_inv_block_tile<ArithF64, B, true>(F, cur, xb, tws + tb * FAST_TW_SLOTS, tl, gout);
        else
            pp_inv_block_tile<ArithU64, B, false>(F, cur, xb, nullptr, tl, gout);
    }
}

// ---- inverse pass A' (levels b..logN-1), x scalar, canonical out ---------------------------------------------
template <class A, bool STAGED>
__device__ __forceinline__ void fast_inv_col_body(const FastArgs& F, int64_t* sm, int limb, long long drow) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const int b = F.logN - 8;
    const typename A::C c = make_const<A>((uint64_t)F.q[limb]);
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

template <int DUMMY>
__global__ void __launch_bounds__(NTT_THREADS, FAST_COL_CTAS) fast_inv_colpass(const FastArgs F) {
    extern __shared__ __align__(16) int64_t sm[];
    const RowId rid = fast_row(F);
    const int limb = rid.limb;
    if (fast_use_f64(F, rid))
        fast_inv_col_body<ArithF64, true>(F, sm, limb, rid.data_row);
    else
        fast_inv_col_body<ArithU64, false>(F, sm, limb, rid.data_row);
}

// ---- PERSISTENT, SOFTWARE-PIPELINED column passes -------------------------------------------------------------------------
// Same idea as fast_fwd_blockpass_pp for the strided passes: a CTA walks tiles (256 rows x 16 columns of one limb row),
// prefetching the next tile's 16 coefficients per thread into registers while the current tile is transformed.  The
// exchange between the two radix-16 rounds crosses warps here, so there is one CTA barrier per tile; the exchange
// buffer is double-buffered so that no second barrier is needed.  Tile order: all column chunks of a row, then the
// next of the G rows that share the limb, then the next limb: the 2 KB of column-pass twiddles are staged once per
// limb (double-buffered, one limb ahead).
constexpr int PPC_SMEM_BYTES = 2 * SMEM_BYTES + 2 * 256 * 8 + 32;

struct ColTile {
    long long drow;
    int limb, ch, group;
};
__device__ __forceinline__ ColTile decode_col_tile(const FastArgs& F, int id, int G, int chunks) {
    const int per_group = G * chunks;
    const int group = id / per_group, rem = id - group * per_group;
    const int g = rem / chunks;
    ColTile t;
    t.group = group;
    t.ch = rem - g * chunks;
    if (F.slab_rows == 0) {
        t.limb = group;
        t.drow = (long long)g * F.period + group;
    } else {
        t.limb = F.slab_t0 + group;
        t.drow = (long long)g * F.group_rows + F.slab_t0 + group;
    }
    return t;
}
__device__ __forceinline__ void pp_issue_col_twiddles(const FastArgs& F, int limb, double* dst, uint64_t* bar) {
    fence_async_smem();
    mbar_expect_tx(bar, 256u * 8u);
    tma_bulk_g2s(dst, F.tw_f64 + ((long long)limb << F.logN), 256u * 8u, bar);
}

template <class A, bool STAGED>
__device__ __forceinline__ void pp_fwd_col_tile(const FastArgs& F, const int64_t (&raw)[16], int64_t* xb,
                                                const typename A::TW* tws, const ColTile& tl) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const int b = F.logN - 8;
    const typename A::C c = make_const<A>((uint64_t)F.q[tl.limb]);
    int64_t* __restrict__ row0 = F.a + tl.drow * F.a_stride + (long long)tl.ch * 16;
    const TW* __restrict__ W = tw_row<A>(F, tl.limb);
    T e[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = A::load_in(raw[k], F.in_raw);
    if (F.scal) {
        const TW s = scalar_tw<A>(F, tl.limb);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = A::mul(e[k], s, c);
    }
    if constexpr (STAGED)
        fast_fwd_round<A, 0>(e, TwSharedCol<TW>{tws, 0, 0u}, c);
    else
        fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, 0, 0u}, c);
    smx_store(xb, e, tau, 8);
    __syncthreads();
    smx_load(xb, e, tau, 4);
    const int hi = tau >> 4, col = tau & 15;
    if constexpr (STAGED)
        fast_fwd_round<A, 0>(e, TwSharedCol<TW>{tws, 4, (unsigned)hi}, c);
    else
        fast_fwd_round<A, 0>(e, TwGlobal<TW>{W, 4, (unsigned)hi}, c);
#pragma unroll
    for (int k = 0; k < 16; ++k) row0[((long long)(hi * 16 + k) << b) + col] = A::store_mid(e[k], c);
}

template <class A, bool STAGED>
__device__ __forceinline__ void pp_inv_col_tile(const FastArgs& F, const int64_t (&raw)[16], int64_t* xb,
                                                const typename A::TW* tws, const ColTile& tl) {
    using T = typename A::T;
    using TW = typename A::TW;
    const int tau = threadIdx.x;
    const int b = F.logN - 8;
    const typename A::C c = make_const<A>((uint64_t)F.q[tl.limb]);
    int64_t* __restrict__ row0 = F.a + tl.drow * F.a_stride + (long long)tl.ch * 16;
    const TW* __restrict__ W = tw_row<A>(F, tl.limb);
    T e[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = A::load_mid(raw[k]);
    const int hi = tau >> 4;
    if constexpr (STAGED)
        fast_inv_round<A, 4>(e, TwSharedCol<TW>{tws, 4, (unsigned)hi}, c);
    else
        fast_inv_round<A, 4>(e, TwGlobal<TW>{W, 4, (unsigned)hi}, c);
    smx_store(xb, e, tau, 4);
    __syncthreads();
    smx_load(xb, e, tau, 8);
    if constexpr (STAGED)
        fast_inv_round<A, 4>(e, TwSharedCol<TW>{tws, 0, 0u}, c);
    else
        fast_inv_round<A, 4>(e, TwGlobal<TW>{W, 0, 0u}, c);
    const TW s = scalar_tw<A>(F, tl.limb);
    const int r0 = tau >> 4, col = tau & 15;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        row0[((long long)(r0 + 16 * k) << b) + col] = A::store_canon(A::mul(e[k], s, c), c, F.centred != 0);
}

template <bool FWD>
__global__ void __launch_bounds__(NTT_THREADS, 2) fast_colpass_pp(const FastArgs F, long long total_tiles, int G) {
    extern __shared__ __align__(16) int64_t sm[];
    double* tws = reinterpret_cast<double*>(sm + 2 * SMEM_SLOTS);               // [2][256]
    uint64_t* bar_tw = reinterpret_cast<uint64_t*>(tws + 2 * 256);               // [2]
    const int tau = threadIdx.x;
    const int b = F.logN - 8;
    const int chunks = (1 << F.logN) / TILE;
    const TileRange range = pp_tile_range(F, (int)(total_tiles / ((long long)chunks * G)), chunks * G);
    const int t_begin = range.begin, t_end = range.end;
    if (t_begin >= t_end) return;
    auto staged = [&](int limb) { return (uint64_t)F.q[limb] < SMALL_PRIME_LIMIT && F.force_int != 1; };
    if (tau == 0) {
        mbar_init(&bar_tw[0], 1);
        mbar_init(&bar_tw[1], 1);
    }
    __syncthreads();
    // element k of a thread: forward reads rows (tau>>4) + 16k, inverse rows 16 (tau>>4) + k, column tau & 15
    const long long off0 = FWD ? ((long long)(tau >> 4) << b) + (tau & 15) : ((long long)((tau >> 4) * 16) << b) + (tau & 15);
    const long long kstep = FWD ? (16ll << b) : (1ll << b);
    int64_t nxt[16];
    {
        const ColTile t0 = decode_col_tile(F, t_begin, G, chunks);
        if (tau == 0 && staged(t0.limb)) pp_issue_col_twiddles(F, t0.limb, tws, &bar_tw[0]);
        const int64_t* __restrict__ src = F.a + t0.drow * F.a_stride + (long long)t0.ch * 16 + off0;
#pragma unroll
        for (int k = 0; k < 16; ++k) nxt[k] = src[k * kstep];
    }
    long long cur_group = -1;
    unsigned gcount = 0, ph0 = 0, ph1 = 0, it = 0;
#pragma unroll 1
    for (int id = t_begin; id < t_end; ++id, ++it) {
        const ColTile tl = decode_col_tile(F, id, G, chunks);
        int64_t cur[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) cur[k] = nxt[k];
        if (id + 1 < t_end) {
            const ColTile tn = decode_col_tile(F, id + 1, G, chunks);
            const int64_t* __restrict__ src = F.a + tn.drow * F.a_stride + (long long)tn.ch * 16 + off0;
#pragma unroll
            for (int k = 0; k < 16; ++k) nxt[k] = src[k * kstep];
        }
        if (tl.group != cur_group) {
            cur_group = tl.group;
            __syncthreads();             // every warp is done with the limb before: its twiddle buffer can be refilled
            const int next_first = (tl.group + 1) * G * chunks;
            if (tau == 0 && next_first < t_end) {
                const ColTile tg = decode_col_tile(F, next_first, G, chunks);
                if (staged(tg.limb)) pp_issue_col_twiddles(F, tg.limb, tws + ((gcount + 1) & 1) * 256, &bar_tw[(gcount + 1) & 1]);
            }
            if (staged(tl.limb)) {
                const unsigned bb = gcount & 1;
                mbar_wait(&bar_tw[bb], (bb ? ph1 : ph0) & 1);
                if (bb) ++ph1; else ++ph0;
            }
            ++gcount;
        }
        const double* tw = tws + ((gcount - 1) & 1) * 256;
        int64_t* xb = sm + (it & 1) * SMEM_SLOTS;   // alternate exchange buffers: one barrier per tile is enough
        const RowId rid{tl.drow, tl.limb};
        if (fast_use_f64(F, rid)) {
            if constexpr (FWD) pp_fwd_col_tile<ArithF64, true>(F, cur, xb, tw, tl);
            else pp_inv_col_tile<ArithF64, true>(F, cur, xb, tw, tl);
        } else {
            if constexpr (FWD) pp_fwd_col_tile<ArithU64, false>(F, cur, xb, nullptr, tl);
            else pp_inv_col_tile<ArithU64, false>(F, cur, xb, nullptr, tl);
        }
    }
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

}  // namespace ckks
