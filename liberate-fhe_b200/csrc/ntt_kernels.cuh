// ntt_kernels.cuh -- two-pass batched negacyclic NTT / iNTT for sm_100a.
//
// Replaces the reference's one-launch-per-stage kernels (ntt_cuda_kernel / intt_cuda_kernel,
// src/liberate/ntt/ntt_cuda_kernel.cu:236-275, 433-473, launched logN times at :318-322, :521-525)
// and the host-side "fused" launch sequences enter_ntt / intt / intt_exit / intt_exit_reduce /
// intt_exit_reduce_signed (:349-423, :476-548, :709-973) by TWO kernels per transform:
//
//   forward : pass A (column kernel, stages 0..7)    -> pass B (block kernel, stages 8..logN-1)
//   inverse : pass B' (block kernel, levels 0..b-1)  -> pass A' (column kernel, levels b..logN-1,
//                                                                  then xN^-1, redc, reduce, signed)
//
// Every CTA owns 4096 coefficients of one limb (256 threads x 16 coefficients in registers) and runs
// radix-16 rounds (4 butterfly stages, 32 butterflies per thread) entirely in registers; rounds
// exchange data through (padded) shared memory.  The limb row crosses HBM/L2 twice per transform
// instead of 2*logN times, and there are no index tables at all (the reference streams 12 bytes of
// even/odd/psi tables per butterfly).  Twiddles are read from a COMPACT per-limb table
// W[C][N] (psi^bitrev(i), Montgomery form) -- 1/logN of the reference's painted psi[C][logN][N/2].
//
// Bit-exactness: every butterfly evaluates exactly the integer expression of the reference
// (mont.cuh), so even the lazy [0,2q) representatives agree.  SG=true keeps that true for inputs
// that are negative (the reference feeds signed values through its butterflies in a few places,
// e.g. create_rotation_key engine.py:1163-1165 and the Garner-extended digits engine.py:733-740).
//
// Local index algebra (shared by all four kernels).  A CTA works on a 12-bit local index z:
//   block kernels : z = position inside the 4096-coefficient chunk;   span = logN, jz = chunk*4096 + z
//   column kernels: z = row*16 + column, 256 rows x 16 columns tile;   span = 12,   jz = z
// A radix-16 round on the 4-bit field [p+3..p] of z gives thread tau the 16 elements
//   z = ((tau >> p) << (p+4)) | (k << p) | (tau & (2^p - 1)),  k = 0..15
// and, with pre = jz_base >> (p+4) (the index bits above the field) and s0 = span - 4 - p,
//   forward stage s0+i pairs (k, k + (8>>i)) with twiddle  W[2^(s0+i) + (pre << i) + (k >> (4-i))]
//   inverse level  with distance 2^i uses the same table index with i' = 3-i.
// (cctx.py:89-142 paint_butterfly_forward/backward written in closed form.)
#pragma once
#include "mont.cuh"

namespace ckks {

constexpr int NTT_THREADS = 256;
constexpr int TILE = 4096;  // coefficients per CTA
// shared-memory padding: 2 slots per 16 coefficients keeps 16-byte alignment for 128-bit accesses and
// makes the 16-consecutive-per-thread pattern conflict-free (stride 18 slots = 36 words).
__device__ __forceinline__ int pad_idx(int z) { return z + ((z >> 4) << 1); }
constexpr int SMEM_SLOTS = TILE + (TILE >> 4) * 2;
constexpr int SMEM_BYTES = SMEM_SLOTS * 8;

__device__ __forceinline__ int zbase(int tau, int p) { return ((tau >> p) << (p + 4)) | (tau & ((1 << p) - 1)); }

// -------------------------------------------------------------------------------------------------
// register rounds
// -------------------------------------------------------------------------------------------------
template <int RUN>
__device__ __forceinline__ void load_tw(uint64_t (&w)[8], const int64_t* __restrict__ wp) {
    if (RUN == 1) {
        w[0] = (uint64_t)__ldg(wp);
    } else {
#pragma unroll
        for (int g = 0; g < RUN; g += 2) {
            const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(wp + g));
            w[g] = (uint64_t)v.x;
            w[g + 1] = (uint64_t)v.y;
        }
    }
}

// forward stages i = FIRST..3 of the notional radix-16 round
template <int FIRST, bool SG>
__device__ __forceinline__ void fwd_round(int64_t (&e)[16], const int64_t* __restrict__ W, int s0, unsigned pre,
                                          const LimbConst& c) {
#pragma unroll
    for (int i = FIRST; i < 4; ++i) {
        const int d = 8 >> i;
        const int64_t* wp = W + ((1u << (s0 + i)) + (pre << i));
        uint64_t w[8];
        if (i == 0) load_tw<1>(w, wp);
        if (i == 1) load_tw<2>(w, wp);
        if (i == 2) load_tw<4>(w, wp);
        if (i == 3) load_tw<8>(w, wp);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (!(k & d)) ct_bfly<SG>(e[k], e[k + d], w[k >> (4 - i)], c);
    }
}

// inverse levels with distance 2^i, i = 0..NST-1 of the notional radix-16 round
template <int NST, bool SG>
__device__ __forceinline__ void inv_round(int64_t (&e)[16], const int64_t* __restrict__ W, int s0, unsigned pre,
                                          const LimbConst& c) {
#pragma unroll
    for (int i = 0; i < NST; ++i) {
        const int d = 1 << i;
        const int ip = 3 - i;
        const int64_t* wp = W + ((1u << (s0 + ip)) + (pre << ip));
        uint64_t w[8];
        if (ip == 0) load_tw<1>(w, wp);
        if (ip == 1) load_tw<2>(w, wp);
        if (ip == 2) load_tw<4>(w, wp);
        if (ip == 3) load_tw<8>(w, wp);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (!(k & d)) gs_bfly<SG>(e[k], e[k + d], w[k >> (i + 1)], c);
    }
}

// -------------------------------------------------------------------------------------------------
// shared-memory exchange helpers
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sm_store_field(int64_t* sm, const int64_t (&e)[16], int tau, int p) {
    const int zb = zbase(tau, p);
    if (p == 0) {
#pragma unroll
        for (int k = 0; k < 16; k += 2)
            *reinterpret_cast<longlong2*>(sm + pad_idx(zb + k)) = make_longlong2(e[k], e[k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) sm[pad_idx(zb | (k << p))] = e[k];
    }
}
__device__ __forceinline__ void sm_load_field(const int64_t* sm, int64_t (&e)[16], int tau, int p) {
    const int zb = zbase(tau, p);
    if (p == 0) {
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const longlong2 v = *reinterpret_cast<const longlong2*>(sm + pad_idx(zb + k));
            e[k] = v.x;
            e[k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = sm[pad_idx(zb | (k << p))];
    }
}
// coalesced 128-bit copy between the (padded) tile in shared memory and a contiguous global chunk
__device__ __forceinline__ void sm_to_global(const int64_t* sm, int64_t* __restrict__ g, int tau) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int z = 2 * tau + it * 512;
        *reinterpret_cast<longlong2*>(g + z) = *reinterpret_cast<const longlong2*>(sm + pad_idx(z));
    }
}
__device__ __forceinline__ void global_to_sm(int64_t* sm, const int64_t* __restrict__ g, int tau) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int z = 2 * tau + it * 512;
        *reinterpret_cast<longlong2*>(sm + pad_idx(z)) = *reinterpret_cast<const longlong2*>(g + z);
    }
}

struct NttArgs {
    int64_t* a;            // [C rows] data, in place
    long long a_stride;    // elements between rows
    const int64_t* tw;     // compact twiddles [C][N] (psi for forward, psi^-1 for inverse), Montgomery form
    long long tw_stride;
    const int64_t* _2q;    // per-limb constants (device arrays of length C), the reference's parameter pack
    const int64_t* ql;
    const int64_t* qh;
    const int64_t* kl;
    const int64_t* kh;
    const int64_t* scal;   // forward: Rs (R^2 mod q) when `enter` is fused, else nullptr; inverse: N^-1 * R mod q
    int logN;
    int exit_mode;         // inverse only: 0 intt, 1 +redc, 2 +reduce, 3 +make_signed
};

// -------------------------------------------------------------------------------------------------
// forward pass A: stages 0..7 on a 256-row x 16-column tile (row stride 2^(logN-8) coefficients)
// grid (2^(logN-8)/16, C)
// -------------------------------------------------------------------------------------------------
template <bool SG, bool ENTER>
__global__ void __launch_bounds__(NTT_THREADS) ntt_fwd_colpass(const NttArgs A) {
    extern __shared__ __align__(16) int64_t sm[];
    const int tau = threadIdx.x;
    const int limb = blockIdx.y;
    const int b = A.logN - 8;
    const LimbConst c = load_limb_const(A._2q, A.ql, A.qh, A.kl, A.kh, limb);
    int64_t* __restrict__ row0 = A.a + (long long)limb * A.a_stride + (long long)blockIdx.x * 16;
    const int64_t* __restrict__ W = A.tw + (long long)limb * A.tw_stride;

    int64_t e[16];
    {   // round 1: field bits 11..8 (rows r = (tau>>4) + 16k), stages 0..3, no index bits above the field
        const int r0 = tau >> 4, col = tau & 15;
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = row0[((long long)(r0 + 16 * k) << b) + col];
        if (ENTER) {
            const uint64_t rs = (uint64_t)A.scal[limb];
#pragma unroll
            for (int k = 0; k < 16; ++k) e[k] = mont_mul_w<true>(e[k], rs, c.q4, c.k);
        }
        fwd_round<0, SG>(e, W, 0, 0u, c);
        sm_store_field(sm, e, tau, 8);
    }
    __syncthreads();
    {   // round 2: field bits 7..4 (rows r = 16*(tau>>4) + k), stages 4..7, pre = tau>>4
        sm_load_field(sm, e, tau, 4);
        const int hi = tau >> 4, col = tau & 15;
        fwd_round<0, SG>(e, W, 4, (unsigned)hi, c);
#pragma unroll
        for (int k = 0; k < 16; ++k) row0[((long long)(hi * 16 + k) << b) + col] = e[k];
    }
}

// -------------------------------------------------------------------------------------------------
// forward pass B: stages 8..logN-1 on 4096 contiguous coefficients (2^(12-B) blocks of 2^B)
// grid (N/4096, C).  B = logN - 8 in [4, 9].
// rounds (field low bit p, first stage): B=4:[0] 5:[1,0*] 6:[2,0*] 7:[3,0*] 8:[4,0] 9:[5,1,0*]  (* partial)
// -------------------------------------------------------------------------------------------------
template <int B, bool SG>
__global__ void __launch_bounds__(NTT_THREADS) ntt_fwd_blockpass(const NttArgs A) {
    extern __shared__ __align__(16) int64_t sm[];
    const int tau = threadIdx.x;
    const int limb = blockIdx.y;
    const unsigned chunk = blockIdx.x;
    const int logN = B + 8;
    const LimbConst c = load_limb_const(A._2q, A.ql, A.qh, A.kl, A.kh, limb);
    int64_t* __restrict__ g = A.a + (long long)limb * A.a_stride + (long long)chunk * TILE;
    const int64_t* __restrict__ W = A.tw + (long long)limb * A.tw_stride;
    int64_t e[16];

    constexpr int P1 = (B >= 4) ? B - 4 : 0;  // first (full) round
    {
        const int zb = zbase(tau, P1);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = g[zb | (k << P1)];
        const unsigned pre = (chunk << (8 - P1)) | (unsigned)(tau >> P1);
        fwd_round<0, SG>(e, W, logN - 4 - P1, pre, c);
    }
    if constexpr (B >= 8) {  // second full round
        constexpr int P2 = (B >= 8) ? B - 8 : 0;
        sm_store_field(sm, e, tau, P1);
        __syncthreads();
        sm_load_field(sm, e, tau, P2);
        const unsigned pre = (chunk << (8 - P2)) | (unsigned)(tau >> P2);
        fwd_round<0, SG>(e, W, logN - 4 - P2, pre, c);
        if constexpr (B == 9) {  // partial round: last stage (distance 1) on field [3..0]
            __syncthreads();
            sm_store_field(sm, e, tau, P2);
            __syncthreads();
            sm_load_field(sm, e, tau, 0);
            fwd_round<3, SG>(e, W, logN - 4, (chunk << 8) | (unsigned)tau, c);
        }
        __syncthreads();
        sm_store_field(sm, e, tau, 0);
    } else if constexpr (B > 4) {  // partial round with rem = B-4 stages on field [3..0]
        sm_store_field(sm, e, tau, P1);
        __syncthreads();
        sm_load_field(sm, e, tau, 0);
        fwd_round<8 - B, SG>(e, W, logN - 4, (chunk << 8) | (unsigned)tau, c);
        __syncthreads();
        sm_store_field(sm, e, tau, 0);
    } else {
        sm_store_field(sm, e, tau, 0);
    }
    __syncthreads();
    sm_to_global(sm, g, tau);
}

// -------------------------------------------------------------------------------------------------
// inverse pass B': levels 0..B-1 on 4096 contiguous coefficients.  grid (N/4096, C)
// rounds: B=4:[0] 5:[0,4*] 6:[0,4*] 7:[0,4*] 8:[0,4] 9:[0,4,8*]   (* partial: first B-4 / 1 levels)
// -------------------------------------------------------------------------------------------------
template <int B, bool SG>
__global__ void __launch_bounds__(NTT_THREADS) ntt_inv_blockpass(const NttArgs A) {
    extern __shared__ __align__(16) int64_t sm[];
    const int tau = threadIdx.x;
    const int limb = blockIdx.y;
    const unsigned chunk = blockIdx.x;
    const int logN = B + 8;
    const LimbConst c = load_limb_const(A._2q, A.ql, A.qh, A.kl, A.kh, limb);
    int64_t* __restrict__ g = A.a + (long long)limb * A.a_stride + (long long)chunk * TILE;
    const int64_t* __restrict__ W = A.tw + (long long)limb * A.tw_stride;
    int64_t e[16];

    global_to_sm(sm, g, tau);
    __syncthreads();
    sm_load_field(sm, e, tau, 0);
    inv_round<4, SG>(e, W, logN - 4, (chunk << 8) | (unsigned)tau, c);
    if constexpr (B == 4) {
        __syncthreads();
        sm_store_field(sm, e, tau, 0);
        __syncthreads();
        sm_to_global(sm, g, tau);
        return;
    } else {
    __syncthreads();
    sm_store_field(sm, e, tau, 0);
    __syncthreads();
    sm_load_field(sm, e, tau, 4);
    {
        const unsigned pre = (chunk << 4) | (unsigned)(tau >> 4);
        constexpr int NST = (B >= 8) ? 4 : B - 4;
        inv_round<NST, SG>(e, W, logN - 8, pre, c);
    }
    if constexpr (B == 9) {
        __syncthreads();
        sm_store_field(sm, e, tau, 4);
        __syncthreads();
        sm_load_field(sm, e, tau, 8);
        inv_round<1, SG>(e, W, logN - 12, chunk, c);
        const int zb = zbase(tau, 8);
#pragma unroll
        for (int k = 0; k < 16; ++k) g[zb | (k << 8)] = e[k];
    } else {
        const int zb = zbase(tau, 4);
#pragma unroll
        for (int k = 0; k < 16; ++k) g[zb | (k << 4)] = e[k];
    }
    }
}

// -------------------------------------------------------------------------------------------------
// inverse pass A': levels b..logN-1 on a 256-row x 16-column tile, then the exit chain
//   x N^-1 (mont_enter with Ninv, kern.cu:527-529) [, mont_redc :754-766] [, reduce :817-832] [, make_signed :884-902]
// grid (2^(logN-8)/16, C)
// -------------------------------------------------------------------------------------------------
template <bool SG>
__global__ void __launch_bounds__(NTT_THREADS) ntt_inv_colpass(const NttArgs A) {
    extern __shared__ __align__(16) int64_t sm[];
    const int tau = threadIdx.x;
    const int limb = blockIdx.y;
    const int b = A.logN - 8;
    const LimbConst c = load_limb_const(A._2q, A.ql, A.qh, A.kl, A.kh, limb);
    int64_t* __restrict__ row0 = A.a + (long long)limb * A.a_stride + (long long)blockIdx.x * 16;
    const int64_t* __restrict__ W = A.tw + (long long)limb * A.tw_stride;
    int64_t e[16];
    {   // field bits 7..4: rows 16*(tau>>4) + k
        const int hi = tau >> 4, col = tau & 15;
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = row0[((long long)(hi * 16 + k) << b) + col];
        inv_round<4, SG>(e, W, 4, (unsigned)hi, c);
        sm_store_field(sm, e, tau, 4);
    }
    __syncthreads();
    {   // field bits 11..8: rows (tau>>4) + 16k
        sm_load_field(sm, e, tau, 8);
        inv_round<4, SG>(e, W, 0, 0u, c);
        const uint64_t ninv = (uint64_t)A.scal[limb];
        const int mode = A.exit_mode;
        const int64_t q = (int64_t)(c.q2 >> 1);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            int64_t v = mont_mul_w<SG>(e[k], ninv, c.q4, c.k);
            if (mode >= 1) v = mont_redc(v, c.q4, c.k);
            if (mode >= 2) v = reduce_q(v, q);
            if (mode >= 3) v = make_signed(v, q);
            e[k] = v;
        }
        const int r0 = tau >> 4, col = tau & 15;
#pragma unroll
        for (int k = 0; k < 16; ++k) row0[((long long)(r0 + 16 * k) << b) + col] = e[k];
    }
}

}  // namespace ckks
