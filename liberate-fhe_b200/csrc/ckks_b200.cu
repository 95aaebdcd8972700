// ckks_b200.cu -- kernels + C ABI of libckks_b200.so (see include/ckks_b200.h for the contract and the
// reference file:line each entry point replaces).  sm_100a only; no torch types anywhere.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ckks_b200.h"
#include "mont.cuh"
#include "ntt_kernels.cuh"
#include "ntt_fast.cuh"
#include "csprng.cuh"

using namespace ckks;

#define CKKS_ABI_VERSION 4

namespace {

constexpr int EW_THREADS = 256;

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
long long g_launches = 0;   // kernels launched by this library since load (ckks_launch_count)
inline int launch_status() { ++g_launches; return (int)cudaGetLastError(); }
int g_pdl = 1;   // ckks_set_option(19, v): hot-path kernels are launched with programmatic stream serialization (pdl_enter())
template <class... KArgs, class... Args>
inline int launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    ++g_launches;
    return (int)(e != cudaSuccess ? e : cudaGetLastError());
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool row_ok(const void* p, long long stride) { return aligned16(p) && (stride % 2 == 0); }

// two coefficients per thread, 128-bit accesses; grid (N/2/EW_THREADS, C)
__device__ __forceinline__ longlong2 ld2(const int64_t* p) { return *reinterpret_cast<const longlong2*>(p); }
__device__ __forceinline__ void st2(int64_t* p, longlong2 v) { *reinterpret_cast<longlong2*>(p) = v; }

struct MontPack {
    const int64_t* _2q;
    const int64_t* ql;
    const int64_t* qh;
    const int64_t* kl;
    const int64_t* kh;
};
__device__ __forceinline__ LimbConst lc(const MontPack& m, int i) { return load_limb_const(m._2q, m.ql, m.qh, m.kl, m.kh, i); }

// ---------------------------------------------------------------------------------------------
// level-1 elementwise kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_mont_mult(const int64_t* __restrict__ a, long long as, const int64_t* __restrict__ b, long long bs,
                            int64_t* __restrict__ c, long long cs, int N, MontPack m) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, i);
    const longlong2 x = ld2(a + i * as + j), y = ld2(b + i * bs + j);
    st2(c + i * cs + j, make_longlong2(mont_mul_ss(x.x, y.x, k.q4, k.k), mont_mul_ss(x.y, y.y, k.q4, k.k)));
}

__global__ void k_mont_enter(int64_t* __restrict__ a, long long as, const int64_t* __restrict__ Rs, int N, MontPack m) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, i);
    const int64_t r = Rs[i];
    longlong2 x = ld2(a + i * as + j);
    x.x = mont_mul_ss(x.x, r, k.q4, k.k);
    x.y = mont_mul_ss(x.y, r, k.q4, k.k);
    st2(a + i * as + j, x);
}

__global__ void k_mont_redc(int64_t* __restrict__ a, long long as, int N, MontPack m) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, i);
    longlong2 x = ld2(a + i * as + j);
    x.x = mont_redc(x.x, k.q4, k.k);
    x.y = mont_redc(x.y, k.q4, k.k);
    st2(a + i * as + j, x);
}

// mode 0 reduce_2q, 1 make_signed, 2 make_unsigned
template <int MODE>
__global__ void k_unary(int64_t* __restrict__ a, long long as, int N, const int64_t* __restrict__ _2q) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const int64_t q = _2q[i] >> 1;
    longlong2 x = ld2(a + i * as + j);
    if (MODE == 0) { x.x = reduce_q(x.x, q); x.y = reduce_q(x.y, q); }
    if (MODE == 1) { x.x = make_signed(x.x, q); x.y = make_signed(x.y, q); }
    if (MODE == 2) { x.x += q; x.y += q; }
    st2(a + i * as + j, x);
}

template <bool SUB, bool RED = false>
__global__ void k_addsub(const int64_t* __restrict__ a, long long as, const int64_t* __restrict__ b, long long bs,
                         int64_t* __restrict__ c, long long cs, int N, const int64_t* __restrict__ _2q) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const int64_t q2 = _2q[i];
    const longlong2 x = ld2(a + i * as + j), y = ld2(b + i * bs + j);
    longlong2 r;
    r.x = SUB ? lazy_sub(x.x, y.x, q2) : lazy_add(x.x, y.x, q2);
    r.y = SUB ? lazy_sub(x.y, y.y, q2) : lazy_add(x.y, y.y, q2);
    if (RED) {   // + reduce_2q (kern.cu:664-680) in the same pass
        r.x = reduce_q(r.x, q2 >> 1);
        r.y = reduce_q(r.y, q2 >> 1);
    }
    st2(c + i * cs + j, r);
}

__global__ void k_tile_unsigned(const int64_t* __restrict__ a, int64_t* __restrict__ dst, long long ds, int N,
                                const int64_t* __restrict__ _2q) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const int64_t q = _2q[i] >> 1;
    longlong2 x = ld2(a + j);
    x.x += q;
    x.y += q;
    st2(dst + i * ds + j, x);
}

// painted psi[C][logN][N/2] -> compact [C][N]  (cctx.py:89-142: block i of stage/level lvl uses one twiddle)
__global__ void k_compact(const int64_t* __restrict__ painted, int64_t* __restrict__ compact, int logN, int forward) {
    const int i = blockIdx.y;
    const int idx = blockIdx.x * EW_THREADS + threadIdx.x;  // compact index in [0, N)
    const int N = 1 << logN;
    if (idx >= N) return;
    int64_t v = 0;
    if (idx > 0) {
        const int lg = 31 - __clz(idx);           // idx = 2^lg + blk
        const int blk = idx - (1 << lg);
        // forward: stage lg, m = 2^lg blocks of t = N/(2m) butterflies; inverse: level with h = 2^lg, t = N/(2h)
        const int t = N >> (lg + 1);
        const int lvl = forward ? lg : (logN - 1 - lg);
        v = painted[((long long)i * logN + lvl) * (N / 2) + (long long)blk * t];
    }
    compact[(long long)i * N + idx] = v;
}

// ---------------------------------------------------------------------------------------------
// level-2 fused kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_rescale(const int64_t* __restrict__ in, long long is, const int64_t* __restrict__ r0,
                          int64_t* __restrict__ out, long long os, int N, const int64_t* __restrict__ scale,
                          int64_t round_at, int canon, MontPack m) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, i);
    const int64_t sc = scale[i], q = (int64_t)(k.q2 >> 1);
    const longlong2 x = ld2(in + i * is + j), r = ld2(r0 + j);
    longlong2 o;
    o.x = reduce_q(mont_mul_ss(x.x - r.x, sc, k.q4, k.k) + (r.x > round_at ? 1 : 0), q);
    o.y = reduce_q(mont_mul_ss(x.y - r.y, sc, k.q4, k.k) + (r.y > round_at ? 1 : 0), q);
    if (canon) {  // the reference leaves slightly negative representatives here; the fused path wants [0,q)
        o.x += (o.x < 0) ? q : 0;
        o.y += (o.y < 0) ? q : 0;
    }
    st2(out + i * os + j, o);
}

// rescale with a per-limb Montgomery scalar folded in before (mult_scalar, engine.py:2052-2098: mont_enter_scalar + reduce_2q,
// then rescale) and / or after it (level_up, engine.py:1410-1467: rescale, then mont_enter_scalar + reduce_2q): the same
// integers as the three-kernel sequences, one pass.  r0 is the dropped limb AFTER the pre-scaling (the caller scales that row).
template <bool PRE, bool POST>
__global__ void k_rescale_scaled(const int64_t* __restrict__ in, long long is, const int64_t* __restrict__ r0,
                                 int64_t* __restrict__ out, long long os, int N, const int64_t* __restrict__ pre,
                                 const int64_t* __restrict__ scale, int64_t round_at, const int64_t* __restrict__ post,
                                 MontPack m) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, i);
    const int64_t sc = scale[i], q = (int64_t)(k.q2 >> 1);
    longlong2 x = ld2(in + i * is + j);
    const longlong2 r = ld2(r0 + j);
    if (PRE) {
        const int64_t s = pre[i];
        x.x = reduce_q(mont_mul_ss(x.x, s, k.q4, k.k), q);
        x.y = reduce_q(mont_mul_ss(x.y, s, k.q4, k.k), q);
    }
    longlong2 o;
    o.x = reduce_q(mont_mul_ss(x.x - r.x, sc, k.q4, k.k) + (r.x > round_at ? 1 : 0), q);
    o.y = reduce_q(mont_mul_ss(x.y - r.y, sc, k.q4, k.k) + (r.y > round_at ? 1 : 0), q);
    if (POST) {
        const int64_t s = post[i];
        o.x = reduce_q(mont_mul_ss(o.x, s, k.q4, k.k), q);
        o.y = reduce_q(mont_mul_ss(o.y, s, k.q4, k.k), q);
    }
    st2(out + i * os + j, o);
}

// plaintext x ciphertext in the NTT domain (mc_mult, engine.py:2100-2140): d0 = mont(p, c0), d1 = mont(p, c1), one pass
__global__ void k_pc_product(const int64_t* __restrict__ p, const int64_t* __restrict__ c0, const int64_t* __restrict__ c1,
                             long long is, int64_t* __restrict__ d0, int64_t* __restrict__ d1, long long os, int N, MontPack m) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, i);
    const long long o = i * is + j;
    const longlong2 a = ld2(p + o), b0 = ld2(c0 + o), b1 = ld2(c1 + o);
    longlong2 r0, r1;
    r0.x = mont_mul_ss(a.x, b0.x, k.q4, k.k);
    r0.y = mont_mul_ss(a.y, b0.y, k.q4, k.k);
    r1.x = mont_mul_ss(a.x, b1.x, k.q4, k.k);
    r1.y = mont_mul_ss(a.y, b1.y, k.q4, k.k);
    st2(d0 + i * os + j, r0);
    st2(d1 + i * os + j, r1);
}

// plaintext + ciphertext (mc_add, engine.py:2142-2175: mont_enter_scale(pt), mont_enter(c0), mont_add, mont_redc, reduce_2q)
__global__ void k_pc_add(const int64_t* __restrict__ p, long long ps, const int64_t* __restrict__ c0, long long cs,
                         int64_t* __restrict__ out, long long os, int N, const int64_t* __restrict__ Rs_scale,
                         const int64_t* __restrict__ Rs, MontPack m) {
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, i);
    const int64_t rp = Rs_scale[i], rc = Rs[i], q = (int64_t)(k.q2 >> 1);
    const longlong2 a = ld2(p + i * ps + j), b = ld2(c0 + i * cs + j);
    longlong2 o;
    o.x = reduce_q(mont_redc(lazy_add(mont_mul_ss(a.x, rp, k.q4, k.k), mont_mul_ss(b.x, rc, k.q4, k.k), (int64_t)k.q2), k.q4, k.k), q);
    o.y = reduce_q(mont_redc(lazy_add(mont_mul_ss(a.y, rp, k.q4, k.k), mont_mul_ss(b.y, rc, k.q4, k.k), (int64_t)k.q2), k.q4, k.k), q);
    st2(out + i * os + j, o);
}

__global__ void k_tensor(const int64_t* __restrict__ x0, const int64_t* __restrict__ x1, const int64_t* __restrict__ y0,
                         const int64_t* __restrict__ y1, long long is, int64_t* __restrict__ d0,
                         int64_t* __restrict__ d1, int64_t* __restrict__ d2, long long os, int N, MontPack m) {
    pdl_enter();
    const int i = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, i);
    const long long o = i * is + j;
    const longlong2 a0 = ld2(x0 + o), a1 = ld2(x1 + o), b0 = ld2(y0 + o), b1 = ld2(y1 + o);
    longlong2 r0, r1, r2;
    r0.x = mont_mul_ss(a0.x, b0.x, k.q4, k.k);
    r0.y = mont_mul_ss(a0.y, b0.y, k.q4, k.k);
    r1.x = lazy_add(mont_mul_ss(a0.x, b1.x, k.q4, k.k), mont_mul_ss(a1.x, b0.x, k.q4, k.k), (int64_t)k.q2);
    r1.y = lazy_add(mont_mul_ss(a0.y, b1.y, k.q4, k.k), mont_mul_ss(a1.y, b0.y, k.q4, k.k), (int64_t)k.q2);
    r2.x = mont_mul_ss(a1.x, b1.x, k.q4, k.k);
    r2.y = mont_mul_ss(a1.y, b1.y, k.q4, k.k);
    const long long oo = i * os + j;
    st2(d0 + oo, r0);
    st2(d1 + oo, r1);
    st2(d2 + oo, r2);
}

constexpr int MAX_ALPHA = 8;

// one thread per coefficient; the alpha rows of the partition are walked sequentially (engine.py:672-702)
__global__ void k_garner(const int64_t* __restrict__ a, long long as, int64_t* __restrict__ st, long long ss, int alpha,
                         int N, const int64_t* __restrict__ Ysc, const int64_t* __restrict__ Ltri, MontPack m) {
    const int j = blockIdx.x * EW_THREADS + threadIdx.x;
    if (j >= N) return;
    int64_t s[MAX_ALPHA], av[MAX_ALPHA];
#pragma unroll
    for (int r = 0; r < MAX_ALPHA; ++r)
        if (r < alpha) av[r] = a[r * as + j];
#pragma unroll
    for (int r = 0; r < MAX_ALPHA; ++r) s[r] = av[0];
#pragma unroll
    for (int i = 0; i < MAX_ALPHA - 1; ++i) {
        if (i < alpha - 1) {
            const LimbConst k = lc(m, i + 1);
            const int64_t Y = mont_mul_ss(av[i + 1] - s[i + 1], Ysc[i], k.q4, k.k);
            s[i + 1] = Y;
#pragma unroll
            for (int r = i + 2; r < MAX_ALPHA; ++r) {
                if (r < alpha) {
                    const LimbConst kr = lc(m, r);
                    s[r] += mont_mul_ss(Y, Ltri[i * alpha + r], kr.q4, kr.k);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < MAX_ALPHA; ++r)
        if (r < alpha) st[r * ss + j] = s[r];
}

// grid (N/2/EW_THREADS, E): target limb t
__global__ void k_extend(const int64_t* __restrict__ st, long long ss, int alpha, int64_t* __restrict__ out,
                         long long os, int E, int N, const int64_t* __restrict__ Rs, const int64_t* __restrict__ Lenter,
                         int canon, MontPack m) {
    const int t = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, t);
    const int64_t q2 = (int64_t)k.q2;
    longlong2 v = ld2(st + j);
    const int64_t rs = Rs[t];
    longlong2 acc;
    acc.x = mont_mul_ss(v.x, rs, k.q4, k.k);
    acc.y = mont_mul_ss(v.y, rs, k.q4, k.k);
    for (int i = 0; i < alpha - 1; ++i) {
        v = ld2(st + (i + 1) * ss + j);
        const int64_t le = Lenter[(long long)i * E + t];
        acc.x = lazy_add(acc.x, mont_mul_ss(v.x, le, k.q4, k.k), q2);
        acc.y = lazy_add(acc.y, mont_mul_ss(v.y, le, k.q4, k.k), q2);
    }
    if (canon) {  // lazy chain can end slightly below zero (signed digits); bring into [0, 2q)
        acc.x += (acc.x < 0) ? q2 : 0;
        acc.y += (acc.y < 0) ? q2 : 0;
    }
    st2(out + t * os + j, acc);
}

// evaluation-key inner product over ALL parts in one pass: acc_i[t] = (+)_p mont(ext[p][t], key_i[p][t])
// ext: [parts*E rows]; key pointers: device arrays of `parts` row-0 pointers (row stride ks).
__global__ void k_ksk_inner(const int64_t* __restrict__ ext, long long es, int parts, const int64_t* const* __restrict__ k0p,
                            const int64_t* const* __restrict__ k1p, long long ks, int64_t* __restrict__ a0,
                            int64_t* __restrict__ a1, long long as, int E, int N, MontPack m) {
    const int t = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, t);
    const int64_t q2 = (int64_t)k.q2;
    longlong2 s0 = make_longlong2(0, 0), s1 = make_longlong2(0, 0);
    for (int p = 0; p < parts; ++p) {
        const longlong2 e = ld2(ext + ((long long)p * E + t) * es + j);
        const longlong2 u = ld2(k0p[p] + t * ks + j), v = ld2(k1p[p] + t * ks + j);
        s0.x = lazy_add(s0.x, mont_mul_ss(e.x, u.x, k.q4, k.k), q2);
        s0.y = lazy_add(s0.y, mont_mul_ss(e.y, u.y, k.q4, k.k), q2);
        s1.x = lazy_add(s1.x, mont_mul_ss(e.x, v.x, k.q4, k.k), q2);
        s1.y = lazy_add(s1.y, mont_mul_ss(e.y, v.y, k.q4, k.k), q2);
    }
    st2(a0 + t * as + j, s0);
    st2(a1 + t * as + j, s1);
}

__global__ void k_ksk_acc(const int64_t* __restrict__ ext, long long es, const int64_t* __restrict__ k0,
                          const int64_t* __restrict__ k1, long long ks, int64_t* __restrict__ a0,
                          int64_t* __restrict__ a1, long long as, int N, int first, MontPack m) {
    const int t = blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, t);
    const int64_t q2 = (int64_t)k.q2;
    const longlong2 e = ld2(ext + t * es + j), u = ld2(k0 + t * ks + j), v = ld2(k1 + t * ks + j);
    longlong2 p0, p1;
    p0.x = mont_mul_ss(e.x, u.x, k.q4, k.k);
    p0.y = mont_mul_ss(e.y, u.y, k.q4, k.k);
    p1.x = mont_mul_ss(e.x, v.x, k.q4, k.k);
    p1.y = mont_mul_ss(e.y, v.y, k.q4, k.k);
    if (!first) {
        const longlong2 c0 = ld2(a0 + t * as + j), c1 = ld2(a1 + t * as + j);
        p0.x = lazy_add(c0.x, p0.x, q2);
        p0.y = lazy_add(c0.y, p0.y, q2);
        p1.x = lazy_add(c1.x, p1.x, q2);
        p1.y = lazy_add(c1.y, p1.y, q2);
    }
    st2(a0 + t * as + j, p0);
    st2(a1 + t * as + j, p1);
}

// ModDown, part 1: the chain on the K special rows (engine.py:863-890 restricted to rows >= L).
// eff[i][j] = value of special row E-1-i at the moment step i reads it.  One thread per coefficient.
__global__ void k_moddown_special(const int64_t* __restrict__ d, long long ds, int L, int K, int N,
                                  const int64_t* __restrict__ PiR, int64_t* __restrict__ eff, MontPack m) {
    pdl_enter();
    const int j = blockIdx.x * EW_THREADS + threadIdx.x;
    if (j >= N) return;
    const int E = L + K;
    int64_t s[MAX_ALPHA];
#pragma unroll
    for (int p = 0; p < MAX_ALPHA; ++p) {
        if (p < K) {
            const int u = E - 1 - p;  // row index
            const LimbConst k = lc(m, u);
            const int64_t q2 = (int64_t)k.q2, q = (int64_t)(k.q2 >> 1);
            int64_t v = d[u * ds + j];
#pragma unroll
            for (int i = 0; i < MAX_ALPHA; ++i) {
                if (i < p) {
                    v = lazy_sub(v, s[i], q2);
                    v = mont_mul_ss(v, PiR[(long long)i * E + u], k.q4, k.k);
                    v = reduce_q(v, q);
                }
            }
            s[p] = v;
            eff[(long long)p * N + j] = v;
        }
    }
}

// coefficient `o` of the Galois image of a canonical row: out[(g j) mod N] = +-in[j]  <=>  out[o] = +-in[(ginv o) mod N], with the
// minus sign when (ginv o) mod 2N >= N; then make_unsigned + reduce_2q (engine.py:1196-1200), i.e. q - v for a flipped v != 0.
// Lets the kernels that READ a rotated polynomial (Garner digits, the ModDown tail's addend) take it straight from the
// unrotated ciphertext: rotate never writes the rotated ciphertext to HBM.
__device__ __forceinline__ int64_t galois_gather(const int64_t* __restrict__ row, unsigned o, unsigned ginv, unsigned N, int64_t q) {
    const unsigned t = (ginv * o) & (2u * N - 1);
    const int64_t v = row[t & (N - 1)];
    return (t >= N && v != 0) ? q - v : v;
}

// ModDown, part 2: ordinary rows.  grid (N/2/EW_THREADS, L)
__global__ void k_moddown_ordinary(const int64_t* __restrict__ d, long long ds, int L, int K, int N,
                                   const int64_t* __restrict__ Rs, const int64_t* __restrict__ PiR,
                                   const int64_t* __restrict__ eff, const int64_t* __restrict__ add, long long adds,
                                   int64_t* __restrict__ out, long long os, int row0, MontPack m, unsigned add_ginv) {
    pdl_enter();
    const int t = row0 + blockIdx.y;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const int E = L + K;
    const LimbConst k = lc(m, t);
    const int64_t q2 = (int64_t)k.q2, q = (int64_t)(k.q2 >> 1);
    const int64_t rs = Rs[t];
    longlong2 v = ld2(d + t * ds + j);
    v.x = mont_mul_ss(v.x, rs, k.q4, k.k);
    v.y = mont_mul_ss(v.y, rs, k.q4, k.k);
    for (int i = 0; i < K; ++i) {
        const longlong2 p = ld2(eff + (long long)i * N + j);
        const int64_t pir = PiR[(long long)i * E + t];
        v.x = reduce_q(mont_mul_ss(lazy_sub(v.x, mont_mul_ss(p.x, rs, k.q4, k.k), q2), pir, k.q4, k.k), q);
        v.y = reduce_q(mont_mul_ss(lazy_sub(v.y, mont_mul_ss(p.y, rs, k.q4, k.k), q2), pir, k.q4, k.k), q);
    }
    v.x = reduce_q(mont_redc(v.x, k.q4, k.k), q);
    v.y = reduce_q(mont_redc(v.y, k.q4, k.k), q);
    if (add) {
        longlong2 a;
        if (add_ginv) {
            a.x = galois_gather(add + t * adds, (unsigned)j, add_ginv, (unsigned)N, q);
            a.y = galois_gather(add + t * adds, (unsigned)j + 1, add_ginv, (unsigned)N, q);
        } else {
            a = ld2(add + t * adds + j);
        }
        v.x = reduce_q(lazy_add(a.x, v.x, q2), q);
        v.y = reduce_q(lazy_add(a.y, v.y, q2), q);
    }
    st2(out + t * os + j, v);
}

__global__ void k_automorphism(const int64_t* __restrict__ in, long long is, int64_t* __restrict__ out, long long os,
                               int N, unsigned g, int canon, const int64_t* __restrict__ _2q) {
    const int i = blockIdx.y;
    const unsigned j = blockIdx.x * EW_THREADS + threadIdx.x;
    if (j >= (unsigned)N) return;
    const unsigned pj = (g * j) & (2u * N - 1);  // g*j mod 2N (N power of two; 32-bit wrap is harmless)
    const unsigned dst = pj & (N - 1);
    int64_t v = in[i * is + j];
    if (pj >= (unsigned)N) v = -v;
    if (canon) {
        const int64_t q = _2q[i] >> 1;
        v = reduce_q(v + q, q);
    }
    out[i * os + dst] = v;
}

// all local partitions in one launch: grid (N/EW_THREADS, nlocal); partition p covers rows [row0[p], row0[p]+alpha[p])
__global__ void k_garner_batched(const int64_t* __restrict__ a, long long as, int64_t* __restrict__ st, long long ss, int N,
                                 const int32_t* __restrict__ row0, const int32_t* __restrict__ alphas,
                                 const int64_t* const* __restrict__ Yp, const int64_t* const* __restrict__ Lp, MontPack m,
                                 unsigned ginv, const int64_t* __restrict__ q_rows) {
    pdl_enter();
    const int p = blockIdx.y;
    const int j = blockIdx.x * EW_THREADS + threadIdx.x;
    if (j >= N) return;
    const int r0 = row0[p], alpha = alphas[p];
    const int64_t* __restrict__ Ysc = Yp[p];
    const int64_t* __restrict__ Ltri = Lp[p];
    int64_t s[MAX_ALPHA], av[MAX_ALPHA];
#pragma unroll
    for (int r = 0; r < MAX_ALPHA; ++r)
        if (r < alpha)
            av[r] = ginv ? galois_gather(a + (long long)(r0 + r) * as, (unsigned)j, ginv, (unsigned)N, q_rows[r0 + r])
                         : a[(long long)(r0 + r) * as + j];
#pragma unroll
    for (int r = 0; r < MAX_ALPHA; ++r) s[r] = av[0];
#pragma unroll
    for (int i = 0; i < MAX_ALPHA - 1; ++i) {
        if (i < alpha - 1) {
            const LimbConst k = lc(m, r0 + i + 1);
            const int64_t Y = mont_mul_ss(av[i + 1] - s[i + 1], Ysc[i], k.q4, k.k);
            s[i + 1] = Y;
#pragma unroll
            for (int r = i + 2; r < MAX_ALPHA; ++r) {
                if (r < alpha) {
                    const LimbConst kr = lc(m, r0 + r);
                    s[r] += mont_mul_ss(Y, Ltri[i * alpha + r], kr.q4, kr.k);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < MAX_ALPHA; ++r)
        if (r < alpha) st[(long long)(r0 + r) * ss + j] = s[r];
}

// every partition extended to the E target rows in one launch: grid (N/2/EW_THREADS, nparts*E), canonicalised to [0,2q)
__global__ void k_extend_batched(const int64_t* const* __restrict__ states, long long ss, const int32_t* __restrict__ alphas,
                                 int64_t* __restrict__ out, long long os, int E, int N, const int64_t* __restrict__ Rs,
                                 const int64_t* const* __restrict__ Lenters, MontPack m) {
    const int row = blockIdx.y;
    const int p = row / E, t = row - p * E;
    const int j = 2 * (blockIdx.x * EW_THREADS + threadIdx.x);
    if (j >= N) return;
    const LimbConst k = lc(m, t);
    const int64_t q2 = (int64_t)k.q2;
    const int64_t* __restrict__ st = states[p];
    const int64_t* __restrict__ Lenter = Lenters[p];
    const int alpha = alphas[p];
    longlong2 v = ld2(st + j);
    const int64_t rs = Rs[t];
    longlong2 acc;
    acc.x = mont_mul_ss(v.x, rs, k.q4, k.k);
    acc.y = mont_mul_ss(v.y, rs, k.q4, k.k);
    for (int i = 0; i < alpha - 1; ++i) {
        v = ld2(st + (long long)(i + 1) * ss + j);
        const int64_t le = Lenter[(long long)i * E + t];
        acc.x = lazy_add(acc.x, mont_mul_ss(v.x, le, k.q4, k.k), q2);
        acc.y = lazy_add(acc.y, mont_mul_ss(v.y, le, k.q4, k.k), q2);
    }
    acc.x += (acc.x < 0) ? q2 : 0;
    acc.y += (acc.y < 0) ? q2 : 0;
    st2(out + (long long)row * os + j, acc);
}

inline dim3 ew_grid(int N, int C) { return dim3((N / 2 + EW_THREADS - 1) / EW_THREADS, C); }
inline dim3 col_grid(int N) { return dim3((N + EW_THREADS - 1) / EW_THREADS); }

template <int B>
static int launch_fwd_block(const NttArgs& A, dim3 grid, cudaStream_t st) {
    ntt_fwd_blockpass<B, true><<<grid, NTT_THREADS, SMEM_BYTES, st>>>(A);
    return launch_status();
}
template <int B>
static int launch_inv_block(const NttArgs& A, dim3 grid, cudaStream_t st) {
    ntt_inv_blockpass<B, true><<<grid, NTT_THREADS, SMEM_BYTES, st>>>(A);
    return launch_status();
}

// ---- tuning knobs (ckks_set_option) -------------------------------------------------------------------------------
int g_prefetch = 0;       // 2: L2 prefetch distance in rows (0 = off: with the row-slab pipelines the next tile is in L2 anyway)
int g_slab_mb = 100;      // 9: MB of extended rows per key-switch slab (L2 residency vs grid size)
int g_pipes = 3;          // 10: internal side streams used by the slab pipelines (1 = everything on the caller's stream)
int g_ntt_slab_mb = 24;   // 11: MB of rows per slab of a big batched transform
int g_fuse_rescale = 1;   // 12: rescale fused into the tensor stage's column pass
int g_split_tail = 1;     // 16: inverse transform + ModDown of the two output polynomials on two streams
int g_packed = 1;         // 17: block passes read the last-group twiddles from the packed tables (when the caller passes them)
int g_min_slabs = 2;      // 22: the key switch's forward chain runs in at least this many slabs (so that small shards also pipeline)
int g_perm = 1;           // 18: the executor keeps NTT-domain data in warp-interleaved order (needs permuted key copies)
#ifdef CKKS_LAB
int g_skip = 0;           // 5 (lab builds only): measurement -- bit 0 skips the column pass, bit 1 the block pass
int g_lab = 0;            // 21 (lab builds only): FastArgs::lab -- bit 0 no butterflies, bit 1 no global loads / stores
#else
constexpr int g_skip = 0;
#endif
inline bool aligned32(const void* p, long long stride) { return (((uintptr_t)p) & 31) == 0 && (stride & 3) == 0; }

// g^-1 mod 2N for odd g (Newton iteration on 32-bit words; 2N <= 2^18)
unsigned galois_inverse(unsigned g, unsigned N) {
    unsigned x = g;                          // correct to 3 bits for odd g
    for (int i = 0; i < 5; ++i) x *= 2u - g * x;
    return x & (2u * N - 1);
}
int sm_count() {
    static int n[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    return n[dev] > 0 ? n[dev] : 148;
}
template <int B>
static int launch_fast_col_b(bool fwd, const FastArgs& F, dim3 grid, cudaStream_t st) {
    if (fwd) {
        cudaFuncSetAttribute(fast_fwd_colpass<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, COL_SMEM_BYTES);
        return launch_k(fast_fwd_colpass<B>, grid, dim3(NTT_THREADS), COL_SMEM_BYTES, st, F);
    } else {
        cudaFuncSetAttribute(fast_inv_colpass<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, COL_SMEM_BYTES);
        return launch_k(fast_inv_colpass<B>, grid, dim3(NTT_THREADS), COL_SMEM_BYTES, st, F);
    }
}
static int launch_fast_col(bool fwd, const FastArgs& F, dim3 grid, cudaStream_t st) {
    switch (F.logN - 8) {
        case 4: return launch_fast_col_b<4>(fwd, F, grid, st);
        case 5: return launch_fast_col_b<5>(fwd, F, grid, st);
        case 6: return launch_fast_col_b<6>(fwd, F, grid, st);
        case 7: return launch_fast_col_b<7>(fwd, F, grid, st);
        case 8: return launch_fast_col_b<8>(fwd, F, grid, st);
        case 9: return launch_fast_col_b<9>(fwd, F, grid, st);
    }
    return CKKS_E_LOGN;
}
template <int B>
static int launch_col_rescale_b(const FastArgs& F, const RescaleIn& R, dim3 grid, cudaStream_t st) {
    cudaFuncSetAttribute(fast_fwd_colpass_rescale<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, COL_SMEM_BYTES);
    return launch_k(fast_fwd_colpass_rescale<B>, grid, dim3(NTT_THREADS), COL_SMEM_BYTES, st, F, R);
}
static int launch_col_rescale(const FastArgs& F, const RescaleIn& R, dim3 grid, cudaStream_t st) {
    switch (F.logN - 8) {
        case 4: return launch_col_rescale_b<4>(F, R, grid, st);
        case 5: return launch_col_rescale_b<5>(F, R, grid, st);
        case 6: return launch_col_rescale_b<6>(F, R, grid, st);
        case 7: return launch_col_rescale_b<7>(F, R, grid, st);
        case 8: return launch_col_rescale_b<8>(F, R, grid, st);
        case 9: return launch_col_rescale_b<9>(F, R, grid, st);
    }
    return CKKS_E_LOGN;
}
template <int B>
static int launch_fast_block(bool fwd, const FastArgs& F, dim3 grid, cudaStream_t st) {
    if (!aligned32(F.a, F.a_stride)) return CKKS_E_ALIGN;   // 256-bit accesses: rows must be 32-byte aligned
    if (fwd) {
        cudaFuncSetAttribute(fast_fwd_blockpass<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLK_SMEM_BYTES);
        return launch_k(fast_fwd_blockpass<B>, grid, dim3(NTT_THREADS), BLK_SMEM_BYTES, st, F);
    } else {
        cudaFuncSetAttribute(fast_inv_blockpass<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, BLK_SMEM_BYTES);
        return launch_k(fast_inv_blockpass<B>, grid, dim3(NTT_THREADS), BLK_SMEM_BYTES, st, F);
    }
}
static int launch_fast_block_any(bool fwd, FastArgs F, dim3 grid, cudaStream_t st) {
    if (!g_packed) { F.twp_u64 = nullptr; F.twp_f64 = nullptr; }
    switch (F.logN - 8) {
        case 4: return launch_fast_block<4>(fwd, F, grid, st);
        case 5: return launch_fast_block<5>(fwd, F, grid, st);
        case 6: return launch_fast_block<6>(fwd, F, grid, st);
        case 7: return launch_fast_block<7>(fwd, F, grid, st);
        case 8: return launch_fast_block<8>(fwd, F, grid, st);
        case 9: return launch_fast_block<9>(fwd, F, grid, st);
    }
    return CKKS_E_LOGN;
}

// ---- side streams: independent slabs of one call run on up to MAX_PIPES internal streams so that the tail of one
// kernel overlaps the head of the next slab's kernels (small grids leave SMs idle at wave boundaries otherwise).
// fork(): the side streams wait for everything already queued on the caller's stream; join(): the caller's stream
// waits for the side streams.  Streams and events are created once per device and reused: the library is
// SINGLE-THREADED PER DEVICE (one host thread issues the calls for a device; torch's model and the reference's).
// PipeScope joins on every exit path, so an error return between fork and join cannot leave a capture open.
constexpr int MAX_PIPES = 4;
struct SidePipes {
    bool ready = false;
    cudaStream_t s[MAX_PIPES];
    cudaEvent_t fork_ev, join_ev[MAX_PIPES];
};
SidePipes g_side[64];
SidePipes* side_pipes() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SidePipes& p = g_side[dev];
    if (!p.ready) {
        for (int i = 0; i < MAX_PIPES; ++i) {
            if (cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&p.join_ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&p.fork_ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        p.ready = true;
    }
    return &p;
}
struct PipeScope {
    SidePipes* p = nullptr;
    cudaStream_t main_st = nullptr;
    int n = 0;
    bool open = false;
    int fork(SidePipes* pipes, cudaStream_t st, int count) {
        p = pipes; main_st = st; n = count;
        if (cudaEventRecord(p->fork_ev, main_st) != cudaSuccess) return (int)cudaGetLastError();
        open = true;
        for (int i = 0; i < n; ++i)
            if (cudaStreamWaitEvent(p->s[i], p->fork_ev, 0) != cudaSuccess) return (int)cudaGetLastError();
        return 0;
    }
    int join() {
        if (!open) return 0;
        open = false;
        int rc = 0;
        for (int i = 0; i < n; ++i) {
            if (cudaEventRecord(p->join_ev[i], p->s[i]) != cudaSuccess) rc = (int)cudaGetLastError();
            if (cudaStreamWaitEvent(main_st, p->join_ev[i], 0) != cudaSuccess) rc = (int)cudaGetLastError();
        }
        return rc;
    }
    ~PipeScope() { join(); }
};

// one batched fast transform; big batches run as row slabs on the internal streams: the first pass of slab s+1 overlaps
// the second pass of slab s and, more importantly, the hand-off between the two passes of a slab stays in L2 instead of
// going through HBM twice (DRAM traffic = the algorithmic 16 B per coefficient).
static int fast_transform(bool fwd, const FastArgs& F0, int rows, cudaStream_t st) {
    const int logN = F0.logN;
    const long long as = F0.a_stride;
    const long long bytes = (long long)rows * 8 << logN;
    const int nslabs = (int)((bytes + ((long long)g_ntt_slab_mb << 20) - 1) / ((long long)g_ntt_slab_mb << 20));
    SidePipes* pipes = (bytes > (96ll << 20) && g_pipes > 1 && !g_skip) ? side_pipes() : nullptr;
    auto passes = [&](const FastArgs& F, dim3 grid, cudaStream_t s) -> int {
        FastArgs Fb = F;
        if (fwd) {
            if (!(g_skip & 1)) { int rc = launch_fast_col(true, F, grid, s); if (rc) return rc; }
            if (g_skip & 2) return 0;
            Fb.scal = nullptr;
            return launch_fast_block_any(true, Fb, grid, s);
        }
        if (!(g_skip & 2)) { int rc = launch_fast_block_any(false, Fb, grid, s); if (rc) return rc; }
        if (g_skip & 1) return 0;
        return launch_fast_col(false, F, grid, s);
    };
    if (pipes) {
        const int npipes = g_pipes < nslabs ? g_pipes : nslabs;
        const int per = (rows + nslabs - 1) / nslabs;
        PipeScope scope;
        int rc = scope.fork(pipes, st, npipes);
        if (rc) return rc;
        int k = 0;
        for (int r0 = 0; r0 < rows; r0 += per, ++k) {
            const int r1 = r0 + per < rows ? r0 + per : rows;
            FastArgs Fs = F0;
            Fs.a = F0.a + (long long)r0 * as;
            Fs.row0 = (F0.row0 + r0) % F0.period;
            rc = passes(Fs, dim3((1 << logN) / TILE, r1 - r0), pipes->s[k % npipes]);
            if (rc) return rc;
        }
        return scope.join();
    }
    return passes(F0, dim3((1 << logN) / TILE, rows), st);
}

}  // namespace

#define RC(x)              \
    do {                   \
        int _rc = (x);     \
        if (_rc) return _rc; \
    } while (0)

#define CHECK_PTRS(...)                                  \
    do {                                                 \
        const void* _p[] = {__VA_ARGS__};                \
        for (const void* x : _p)                         \
            if (!x) return CKKS_E_BADARG;                \
    } while (0)

static FastArgs fast_args(int64_t* a, long long as, const void* tw_u64, const double* tw_f64, const void* twp_u64,
                          const double* twp_f64, const int64_t* q, const double* qinv, const int64_t* scal,
                          const uint64_t* scal_sh, int period, int logN) {
    FastArgs F{};
    F.a = a; F.a_stride = as;
    F.tw_u64 = reinterpret_cast<const ulonglong2*>(tw_u64); F.tw_f64 = tw_f64;
    F.twp_u64 = reinterpret_cast<const ulonglong2*>(twp_u64); F.twp_f64 = twp_f64;
    F.q = q; F.qinv = qinv; F.scal = scal; F.scal_sh = scal_sh;
    F.period = period; F.logN = logN;
    F.prefetch = g_prefetch;
#ifdef CKKS_LAB
    F.lab = g_lab;
#endif
    return F;
}
static int fast_check(const void* a, long long as, int rows, int period, int logN, const void* tw_u64, const double* tw_f64,
                      const void* twp_u64, const double* twp_f64, int force_int) {
    if (rows <= 0 || period <= 0) return CKKS_E_BADARG;
    if (!force_int && !tw_f64) return CKKS_E_BADARG;
    if (logN < 12 || logN > 17) return CKKS_E_LOGN;
    if (!row_ok(a, as) || !aligned16(tw_u64) || (tw_f64 && !aligned16(tw_f64)) || (twp_u64 && !aligned16(twp_u64)) ||
        (twp_f64 && !aligned16(twp_f64)))
        return CKKS_E_ALIGN;
    return 0;
}

extern "C" {

int ckks_abi_version(void) { return CKKS_ABI_VERSION; }

int64_t ckks_launch_count(void) { return (int64_t)g_launches; }

static int* option_slot(int key) {
    switch (key) {
        case 2: return &g_prefetch;
#ifdef CKKS_LAB
        case 5: return &g_skip;
        case 21: return &g_lab;
#endif
        case 9: return &g_slab_mb;
        case 10: return &g_pipes;
        case 11: return &g_ntt_slab_mb;
        case 12: return &g_fuse_rescale;
        case 16: return &g_split_tail;
        case 17: return &g_packed;
        case 18: return &g_perm;
        case 19: return &g_pdl;
        case 22: return &g_min_slabs;
    }
    return nullptr;
}
int ckks_get_option(int key) {
    const int* p = option_slot(key);
    return p ? *p : CKKS_E_BADARG;
}
int ckks_set_option(int key, int value) {
    int* p = option_slot(key);
    if (!p) return CKKS_E_BADARG;
    if ((key == 9 || key == 11) && value < 1) return CKKS_E_BADARG;
    if (key == 10 && (value < 1 || value > MAX_PIPES)) return CKKS_E_BADARG;
    *p = value;
    return 0;
}

int ckks_mont_mult(const int64_t* a, int64_t as, const int64_t* b, int64_t bs, int64_t* c, int64_t cs, int C, int N,
                   const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(a, b, c, ql, qh, kl, kh);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(a, as) || !row_ok(b, bs) || !row_ok(c, cs)) return CKKS_E_ALIGN;
    k_mont_mult<<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(a, as, b, bs, c, cs, N, MontPack{nullptr, ql, qh, kl, kh});
    return launch_status();
}

int ckks_mont_enter(int64_t* a, int64_t as, const int64_t* Rs, int C, int N, const int64_t* ql, const int64_t* qh,
                    const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(a, Rs, ql, qh, kl, kh);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(a, as)) return CKKS_E_ALIGN;
    k_mont_enter<<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(a, as, Rs, N, MontPack{nullptr, ql, qh, kl, kh});
    return launch_status();
}

int ckks_mont_redc(int64_t* a, int64_t as, int C, int N, const int64_t* ql, const int64_t* qh, const int64_t* kl,
                   const int64_t* kh, void* stream) {
    CHECK_PTRS(a, ql, qh, kl, kh);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(a, as)) return CKKS_E_ALIGN;
    k_mont_redc<<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(a, as, N, MontPack{nullptr, ql, qh, kl, kh});
    return launch_status();
}

#define UNARY_IMPL(NAME, MODE)                                                                          \
    int NAME(int64_t* a, int64_t as, int C, int N, const int64_t* _2q, void* stream) {                  \
        CHECK_PTRS(a, _2q);                                                                             \
        if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;                                          \
        if (!row_ok(a, as)) return CKKS_E_ALIGN;                                                        \
        k_unary<MODE><<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(a, as, N, _2q);                      \
        return launch_status();                                                                         \
    }
UNARY_IMPL(ckks_reduce_2q, 0)
UNARY_IMPL(ckks_make_signed, 1)
UNARY_IMPL(ckks_make_unsigned, 2)

int ckks_mont_add(const int64_t* a, int64_t as, const int64_t* b, int64_t bs, int64_t* c, int64_t cs, int C, int N,
                  const int64_t* _2q, void* stream) {
    CHECK_PTRS(a, b, c, _2q);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(a, as) || !row_ok(b, bs) || !row_ok(c, cs)) return CKKS_E_ALIGN;
    k_addsub<false><<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(a, as, b, bs, c, cs, N, _2q);
    return launch_status();
}
int ckks_addsub_reduce(const int64_t* a, int64_t as, const int64_t* b, int64_t bs, int64_t* c, int64_t cs, int C, int N,
                       const int64_t* _2q, int sub, void* stream) {
    CHECK_PTRS(a, b, c, _2q);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(a, as) || !row_ok(b, bs) || !row_ok(c, cs)) return CKKS_E_ALIGN;
    if (sub) k_addsub<true, true><<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(a, as, b, bs, c, cs, N, _2q);
    else k_addsub<false, true><<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(a, as, b, bs, c, cs, N, _2q);
    return launch_status();
}

int ckks_mont_sub(const int64_t* a, int64_t as, const int64_t* b, int64_t bs, int64_t* c, int64_t cs, int C, int N,
                  const int64_t* _2q, void* stream) {
    CHECK_PTRS(a, b, c, _2q);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(a, as) || !row_ok(b, bs) || !row_ok(c, cs)) return CKKS_E_ALIGN;
    k_addsub<true><<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(a, as, b, bs, c, cs, N, _2q);
    return launch_status();
}

int ckks_tile_unsigned(const int64_t* a, int64_t* dst, int64_t ds, int C, int N, const int64_t* _2q, void* stream) {
    CHECK_PTRS(a, dst, _2q);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!aligned16(a) || !row_ok(dst, ds)) return CKKS_E_ALIGN;
    k_tile_unsigned<<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(a, dst, ds, N, _2q);
    return launch_status();
}

int ckks_compact_twiddles(const int64_t* painted, int64_t* compact, int C, int logN, int forward, void* stream) {
    CHECK_PTRS(painted, compact);
    if (C <= 0 || logN < 1 || logN > 20) return CKKS_E_BADARG;
    const int N = 1 << logN;
    k_compact<<<dim3((N + EW_THREADS - 1) / EW_THREADS, C), EW_THREADS, 0, S(stream)>>>(painted, compact, logN, forward);
    return launch_status();
}

// ---- NTT -------------------------------------------------------------------------------------------
int ckks_ntt(int64_t* a, int64_t as, int C, int logN, const int64_t* tw, int64_t tws, const int64_t* Rs,
             const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh,
             void* stream) {
    CHECK_PTRS(a, tw, _2q, ql, qh, kl, kh);
    if (C <= 0) return CKKS_E_BADARG;
    if (logN < 12 || logN > 17) return CKKS_E_LOGN;
    if (!row_ok(a, as) || !row_ok(tw, tws)) return CKKS_E_ALIGN;
    const int N = 1 << logN;
    NttArgs A{a, as, tw, tws, _2q, ql, qh, kl, kh, Rs, logN, 0};
    cudaStream_t st = S(stream);
    const dim3 grid(N / TILE, C);
    if (Rs)
        ntt_fwd_colpass<true, true><<<grid, NTT_THREADS, SMEM_BYTES, st>>>(A);
    else
        ntt_fwd_colpass<true, false><<<grid, NTT_THREADS, SMEM_BYTES, st>>>(A);
    int rc = launch_status();
    if (rc) return rc;
    switch (logN - 8) {
        case 4: return launch_fwd_block<4>(A, grid, st);
        case 5: return launch_fwd_block<5>(A, grid, st);
        case 6: return launch_fwd_block<6>(A, grid, st);
        case 7: return launch_fwd_block<7>(A, grid, st);
        case 8: return launch_fwd_block<8>(A, grid, st);
        case 9: return launch_fwd_block<9>(A, grid, st);
    }
    return CKKS_E_LOGN;
}

int ckks_intt(int64_t* a, int64_t as, int C, int logN, const int64_t* tw, int64_t tws, const int64_t* Ninv,
              const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh,
              int exit_mode, void* stream) {
    CHECK_PTRS(a, tw, Ninv, _2q, ql, qh, kl, kh);
    if (C <= 0 || exit_mode < 0 || exit_mode > 3) return CKKS_E_BADARG;
    if (logN < 12 || logN > 17) return CKKS_E_LOGN;
    if (!row_ok(a, as) || !row_ok(tw, tws)) return CKKS_E_ALIGN;
    const int N = 1 << logN;
    NttArgs A{a, as, tw, tws, _2q, ql, qh, kl, kh, Ninv, logN, exit_mode};
    cudaStream_t st = S(stream);
    const dim3 grid(N / TILE, C);
    int rc = CKKS_E_LOGN;
    switch (logN - 8) {
        case 4: rc = launch_inv_block<4>(A, grid, st); break;
        case 5: rc = launch_inv_block<5>(A, grid, st); break;
        case 6: rc = launch_inv_block<6>(A, grid, st); break;
        case 7: rc = launch_inv_block<7>(A, grid, st); break;
        case 8: rc = launch_inv_block<8>(A, grid, st); break;
        case 9: rc = launch_inv_block<9>(A, grid, st); break;
    }
    if (rc) return rc;
    ntt_inv_colpass<true><<<grid, NTT_THREADS, SMEM_BYTES, st>>>(A);
    return launch_status();
}

// ---- canonical-output fast transforms --------------------------------------------------------------------
int ckks_fast_tables(const int64_t* plain, const int64_t* q, void* tw_u64, double* tw_f64, int C, int N, void* stream) {
    CHECK_PTRS(plain, q, tw_u64);
    if (C <= 0 || N <= 0) return CKKS_E_BADARG;
    fast_tables_kernel<<<dim3((N + 255) / 256, C), 256, 0, S(stream)>>>(plain, q, reinterpret_cast<ulonglong2*>(tw_u64),
                                                                        tw_f64, N);
    return launch_status();
}

int ckks_fast_pack(const void* tw_u64, const double* tw_f64, void* twp_u64, double* twp_f64, int C, int logN, void* stream) {
    if (C <= 0) return CKKS_E_BADARG;
    if (logN < 12 || logN > 17) return CKKS_E_LOGN;
    if ((twp_u64 && !tw_u64) || (twp_f64 && !tw_f64) || (!twp_u64 && !twp_f64)) return CKKS_E_BADARG;
    if ((twp_u64 && !aligned16(twp_u64)) || (twp_f64 && !aligned16(twp_f64))) return CKKS_E_ALIGN;
    const int threads = (1 << logN) >> 4;
    fast_pack_kernel<<<dim3((threads + 255) / 256, C), 256, 0, S(stream)>>>(
        reinterpret_cast<const ulonglong2*>(tw_u64), tw_f64, reinterpret_cast<ulonglong2*>(twp_u64), twp_f64, logN);
    return launch_status();
}

int ckks_perm_rows(const int64_t* in, int64_t in_stride, int64_t* out, int64_t out_stride, int rows, int N, int inverse,
                   void* stream) {
    CHECK_PTRS(in, out);
    if (rows <= 0 || N < PACK_TILE || (N & (N - 1)) || in == out) return CKKS_E_BADARG;
    fast_perm_kernel<<<dim3((N + 255) / 256, rows), 256, 0, S(stream)>>>(in, in_stride, out, out_stride, N, inverse);
    return launch_status();
}

int ckks_ntt_fast(int64_t* a, int64_t as, int rows, int period, int logN, const void* tw_u64, const double* tw_f64,
                  const void* twp_u64, const double* twp_f64, const int64_t* q, const double* qinv, const int64_t* scal,
                  const uint64_t* scal_sh, int force_int, int perm, void* stream) {
    CHECK_PTRS(a, tw_u64, q);
    if (scal && !scal_sh) return CKKS_E_BADARG;
    RC(fast_check(a, as, rows, period, logN, tw_u64, tw_f64, twp_u64, twp_f64, force_int));
    FastArgs F = fast_args(a, as, tw_u64, tw_f64, twp_u64, twp_f64, q, qinv, scal, scal_sh, period, logN);
    F.force_int = force_int ? 1 : 0;
    F.perm = perm ? 1 : 0;
    return fast_transform(true, F, rows, S(stream));
}

int ckks_intt_fast(int64_t* a, int64_t as, int rows, int period, int logN, const void* tw_u64, const double* tw_f64,
                   const void* twp_u64, const double* twp_f64, const int64_t* q, const double* qinv, const int64_t* scal,
                   const uint64_t* scal_sh, int centred, int force_int, int perm, void* stream) {
    CHECK_PTRS(a, tw_u64, q, scal, scal_sh);
    RC(fast_check(a, as, rows, period, logN, tw_u64, tw_f64, twp_u64, twp_f64, force_int));
    FastArgs F = fast_args(a, as, tw_u64, tw_f64, twp_u64, twp_f64, q, qinv, scal, scal_sh, period, logN);
    F.force_int = force_int ? 1 : 0;
    F.centred = centred;
    F.perm = perm ? 1 : 0;
    return fast_transform(false, F, rows, S(stream));
}

// ---- level 2 -----------------------------------------------------------------------------------------
int ckks_rescale(const int64_t* in, int64_t is, const int64_t* r0, int64_t* out, int64_t os, int C, int N,
                 const int64_t* scale, int64_t round_at, int canon, const int64_t* _2q, const int64_t* ql,
                 const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(in, r0, out, scale, _2q, ql, qh, kl, kh);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(in, is) || !row_ok(out, os) || !aligned16(r0)) return CKKS_E_ALIGN;
    k_rescale<<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(in, is, r0, out, os, N, scale, round_at, canon,
                                                           MontPack{_2q, ql, qh, kl, kh});
    return launch_status();
}

int ckks_rescale_scaled(const int64_t* in, int64_t is, const int64_t* r0, int64_t* out, int64_t os, int C, int N,
                        const int64_t* pre, const int64_t* scale, int64_t round_at, const int64_t* post, const int64_t* _2q,
                        const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(in, r0, out, scale, _2q, ql, qh, kl, kh);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(in, is) || !row_ok(out, os) || !aligned16(r0)) return CKKS_E_ALIGN;
    const MontPack m{_2q, ql, qh, kl, kh};
    const dim3 grid = ew_grid(N, C);
    cudaStream_t st = S(stream);
    if (pre && post) k_rescale_scaled<true, true><<<grid, EW_THREADS, 0, st>>>(in, is, r0, out, os, N, pre, scale, round_at, post, m);
    else if (pre) k_rescale_scaled<true, false><<<grid, EW_THREADS, 0, st>>>(in, is, r0, out, os, N, pre, scale, round_at, post, m);
    else if (post) k_rescale_scaled<false, true><<<grid, EW_THREADS, 0, st>>>(in, is, r0, out, os, N, pre, scale, round_at, post, m);
    else k_rescale<<<grid, EW_THREADS, 0, st>>>(in, is, r0, out, os, N, scale, round_at, 0, m);
    return launch_status();
}

int ckks_pc_product(const int64_t* p, const int64_t* c0, const int64_t* c1, int64_t is, int64_t* d0, int64_t* d1, int64_t os,
                    int C, int N, const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl,
                    const int64_t* kh, void* stream) {
    CHECK_PTRS(p, c0, c1, d0, d1, _2q, ql, qh, kl, kh);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(p, is) || !row_ok(c0, is) || !row_ok(c1, is) || !row_ok(d0, os) || !row_ok(d1, os)) return CKKS_E_ALIGN;
    k_pc_product<<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(p, c0, c1, is, d0, d1, os, N, MontPack{_2q, ql, qh, kl, kh});
    return launch_status();
}

int ckks_pc_add(const int64_t* p, int64_t ps, const int64_t* c0, int64_t cs, int64_t* out, int64_t os, int C, int N,
                const int64_t* Rs_scale, const int64_t* Rs, const int64_t* _2q, const int64_t* ql, const int64_t* qh,
                const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(p, c0, out, Rs_scale, Rs, _2q, ql, qh, kl, kh);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(p, ps) || !row_ok(c0, cs) || !row_ok(out, os)) return CKKS_E_ALIGN;
    k_pc_add<<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(p, ps, c0, cs, out, os, N, Rs_scale, Rs, MontPack{_2q, ql, qh, kl, kh});
    return launch_status();
}

int ckks_tensor_product(const int64_t* x0, const int64_t* x1, const int64_t* y0, const int64_t* y1, int64_t is,
                        int64_t* d0, int64_t* d1, int64_t* d2, int64_t os, int C, int N, const int64_t* _2q,
                        const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(x0, x1, y0, y1, d0, d1, d2, _2q, ql, qh, kl, kh);
    if (C <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(x0, is) || !row_ok(x1, is) || !row_ok(y0, is) || !row_ok(y1, is) || !row_ok(d0, os) ||
        !row_ok(d1, os) || !row_ok(d2, os))
        return CKKS_E_ALIGN;
    k_tensor<<<ew_grid(N, C), EW_THREADS, 0, S(stream)>>>(x0, x1, y0, y1, is, d0, d1, d2, os, N,
                                                          MontPack{_2q, ql, qh, kl, kh});
    return launch_status();
}

int ckks_garner_digits(const int64_t* a, int64_t as, int64_t* state, int64_t ss, int alpha, int N,
                       const int64_t* Y_scalar, const int64_t* Ltri, const int64_t* ql, const int64_t* qh,
                       const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(a, state, ql, qh, kl, kh);
    if (alpha <= 0 || alpha > MAX_ALPHA || N <= 0) return CKKS_E_BADARG;
    if (alpha > 1 && !Y_scalar) return CKKS_E_BADARG;
    if (alpha > 2 && !Ltri) return CKKS_E_BADARG;
    k_garner<<<col_grid(N), EW_THREADS, 0, S(stream)>>>(a, as, state, ss, alpha, N, Y_scalar, Ltri,
                                                        MontPack{nullptr, ql, qh, kl, kh});
    return launch_status();
}

int ckks_extend(const int64_t* state, int64_t ss, int alpha, int64_t* out, int64_t os, int E, int N, const int64_t* Rs,
                const int64_t* Lenter, int canon, const int64_t* _2q, const int64_t* ql, const int64_t* qh,
                const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(state, out, Rs, _2q, ql, qh, kl, kh);
    if (alpha <= 0 || E <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (alpha > 1 && !Lenter) return CKKS_E_BADARG;
    if (!row_ok(state, ss) || !row_ok(out, os)) return CKKS_E_ALIGN;
    k_extend<<<ew_grid(N, E), EW_THREADS, 0, S(stream)>>>(state, ss, alpha, out, os, E, N, Rs, Lenter, canon,
                                                          MontPack{_2q, ql, qh, kl, kh});
    return launch_status();
}

int ckks_ksk_accumulate(const int64_t* ext, int64_t es, const int64_t* ksk0, const int64_t* ksk1, int64_t ks,
                        int64_t* acc0, int64_t* acc1, int64_t as, int E, int N, int first, const int64_t* _2q,
                        const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(ext, ksk0, ksk1, acc0, acc1, _2q, ql, qh, kl, kh);
    if (E <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(ext, es) || !row_ok(ksk0, ks) || !row_ok(ksk1, ks) || !row_ok(acc0, as) || !row_ok(acc1, as))
        return CKKS_E_ALIGN;
    k_ksk_acc<<<ew_grid(N, E), EW_THREADS, 0, S(stream)>>>(ext, es, ksk0, ksk1, ks, acc0, acc1, as, N, first,
                                                           MontPack{_2q, ql, qh, kl, kh});
    return launch_status();
}

int ckks_ksk_inner(const int64_t* ext, int64_t es, int parts, const int64_t* const* k0_ptrs,
                   const int64_t* const* k1_ptrs, int64_t ks, int64_t* acc0, int64_t* acc1, int64_t as, int E, int N,
                   const int64_t* _2q, const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh,
                   void* stream) {
    CHECK_PTRS(ext, k0_ptrs, k1_ptrs, acc0, acc1, _2q, ql, qh, kl, kh);
    if (E <= 0 || parts <= 0 || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(ext, es) || (ks & 1) || !row_ok(acc0, as) || !row_ok(acc1, as)) return CKKS_E_ALIGN;
    k_ksk_inner<<<ew_grid(N, E), EW_THREADS, 0, S(stream)>>>(ext, es, parts, k0_ptrs, k1_ptrs, ks, acc0, acc1, as, E, N,
                                                             MontPack{_2q, ql, qh, kl, kh});
    return launch_status();
}

int ckks_moddown(int64_t* d, int64_t ds, int L, int K, int N, const int64_t* Rs, const int64_t* PiR,
                 const int64_t* add, int64_t adds, int64_t* out, int64_t os, int64_t* eff, const int64_t* _2q,
                 const int64_t* ql, const int64_t* qh, const int64_t* kl, const int64_t* kh, void* stream) {
    CHECK_PTRS(d, Rs, PiR, out, eff, _2q, ql, qh, kl, kh);
    if (L <= 0 || K <= 0 || K > MAX_ALPHA || N <= 0 || (N & 1)) return CKKS_E_BADARG;
    if (!row_ok(d, ds) || !row_ok(out, os) || (add && !row_ok(add, adds))) return CKKS_E_ALIGN;
    if (!aligned16(eff)) return CKKS_E_ALIGN;
    const MontPack m{_2q, ql, qh, kl, kh};
    k_moddown_special<<<col_grid(N), EW_THREADS, 0, S(stream)>>>(d, ds, L, K, N, PiR, eff, m);
    int rc = launch_status();
    if (rc) return rc;
    k_moddown_ordinary<<<ew_grid(N, L), EW_THREADS, 0, S(stream)>>>(d, ds, L, K, N, Rs, PiR, eff, add, adds, out, os, 0, m, 0u);
    return launch_status();
}

// ---- level 3: the fused executor (one C call = a whole stage of the hot path) ------------------------------
int ckks_exec_digits(const ckks_level_t* lv, const int64_t* a, int64_t as, int64_t* digits, int64_t ds, int64_t galois,
                     void* stream) {
    CHECK_PTRS(lv, a, digits);
    if (lv->nlocal <= 0) return 0;
    if (lv->amax > MAX_ALPHA) return CKKS_E_BADARG;   // k_garner_batched keeps at most MAX_ALPHA digits per partition
    const int N = 1 << lv->logN;
    if (galois && (!(galois & 1) || galois < 0 || galois >= 2ll * N || a == digits)) return CKKS_E_BADARG;
    return launch_k(k_garner_batched, dim3((N + EW_THREADS - 1) / EW_THREADS, lv->nlocal), dim3(EW_THREADS), 0, S(stream),
                    a, as, digits, ds, N, lv->loc_row0, lv->loc_alpha, lv->loc_Y, lv->loc_Ltri,
                    MontPack{nullptr, lv->ql, lv->qh, lv->kl, lv->kh}, galois ? galois_inverse((unsigned)galois, (unsigned)N) : 0u,
                    lv->q);
}

static FastArgs level_fast(const ckks_level_t* lv, int64_t* a, long long as, bool fwd, const int64_t* scal, const int64_t* scal_sh,
                           int period) {
    return fwd ? fast_args(a, as, lv->twf_u64, lv->twf_f64, lv->twpf_u64, lv->twpf_f64, lv->q, lv->qinv, scal,
                           (const uint64_t*)scal_sh, period, lv->logN)
               : fast_args(a, as, lv->twi_u64, lv->twi_f64, lv->twpi_u64, lv->twpi_f64, lv->q, lv->qinv, scal,
                           (const uint64_t*)scal_sh, period, lv->logN);
}

int ckks_exec_tensor_stage(const ckks_level_t* lv, const int64_t* a0, const int64_t* a1, const int64_t* b0,
                           const int64_t* b1, int64_t in_stride, const int64_t* r0a0, const int64_t* r0a1,
                           const int64_t* r0b0, const int64_t* r0b1, int64_t* x, int64_t* d, int64_t* digits,
                           void* stream) {
    CHECK_PTRS(lv, a0, a1, b0, b1, r0a0, r0a1, r0b0, r0b1, x, d, digits);
    const int L = lv->L, N = 1 << lv->logN;
    const long long LN = (long long)L * N;
    const int64_t* in[4] = {a0, a1, b0, b1};
    const int64_t* r0[4] = {r0a0, r0a1, r0b0, r0b1};
    cudaStream_t st = S(stream);
    // NTT-domain order inside this stage: warp-interleaved (the tensor product is pointwise)
    const int perm = g_perm ? 1 : 0;
    if (g_fuse_rescale && aligned16(in[0]) && aligned16(in[1]) && aligned16(in[2]) && aligned16(in[3])) {
        // rescale fused into the load of the batched column pass (no rescaled polynomial is ever written to HBM)
        RescaleIn R{};
        for (int c = 0; c < 4; ++c) { R.in[c] = in[c]; R.r0[c] = r0[c]; }
        R.in_stride = in_stride;
        R.scale = lv->rescale_scale;
        R.round_at = lv->round_at;
        R._2q = lv->_2q; R.ql = lv->ql; R.qh = lv->qh; R.kl = lv->kl; R.kh = lv->kh;
        R.L = L;
        FastArgs F = level_fast(lv, x, N, true, lv->sR, lv->sR_sh, L);
        F.prefetch = 0;
        const dim3 grid(N / TILE, 4 * L);
        RC(launch_col_rescale(F, R, grid, st));
        F.scal = nullptr;
        F.prefetch = g_prefetch;
        F.perm = perm;
        RC(launch_fast_block_any(true, F, grid, st));
    } else {
        for (int c = 0; c < 4; ++c)
            RC(ckks_rescale(in[c], in_stride, r0[c], x + c * LN, N, L, N, lv->rescale_scale, lv->round_at, 1, lv->_2q, lv->ql,
                            lv->qh, lv->kl, lv->kh, stream));
        FastArgs F = level_fast(lv, x, N, true, lv->sR, lv->sR_sh, L);
        F.perm = perm;
        RC(fast_transform(true, F, 4 * L, st));
    }
    RC(launch_k(k_tensor, ew_grid(N, L), dim3(EW_THREADS), 0, st, x, x + LN, x + 2 * LN, x + 3 * LN, N, d, d + LN, d + 2 * LN, N, N,
                MontPack{lv->_2q, lv->ql, lv->qh, lv->kl, lv->kh}));
    FastArgs Fi = level_fast(lv, d, N, false, lv->sExit, lv->sExit_sh, L);
    Fi.perm = perm;
    RC(fast_transform(false, Fi, 3 * L, st));
    return ckks_exec_digits(lv, d + 2 * LN, N, digits, N, 0, stream);
}

int ckks_exec_keyswitch_stage(const ckks_level_t* lv, const int64_t* const* digit_ptrs, int64_t digit_stride,
                              const int64_t* const* k0_ptrs, const int64_t* const* k1_ptrs, int64_t ksk_stride,
                              int keys_permuted, const int64_t* add0, const int64_t* add1, int64_t add_stride,
                              int64_t add0_galois, int64_t* out0, int64_t* out1, int64_t out_stride, int64_t* ws,
                              int phase, int part_begin, int part_end, void* stream) {
    if (phase < 1 || phase > 3) return CKKS_E_BADARG;
    const bool do_fwd = phase & 1, do_tail = phase & 2;
    CHECK_PTRS(lv, ws);
    // partition range of a forward-only call: one process per GPU transforms its own partitions while the peers' digits
    // are still on the wire, and the rest when they have arrived
    const int pb = (part_begin < 0) ? 0 : part_begin, pe = (part_end < 0) ? lv->nparts : part_end;
    if (pb < 0 || pe > lv->nparts || pb > pe || ((pb != 0 || pe != lv->nparts) && phase != 1)) return CKKS_E_BADARG;
    if (pb == pe) return 0;
    if (do_fwd) CHECK_PTRS(digit_ptrs);
    if (do_tail) CHECK_PTRS(k0_ptrs, k1_ptrs, out0, out1);
    if (add0_galois && (!(add0_galois & 1) || add0_galois < 0 || add0_galois >= (2ll << lv->logN) || !add0 || add0 == out0))
        return CKKS_E_BADARG;
    const unsigned add_ginv[2] = {add0_galois ? galois_inverse((unsigned)add0_galois, 1u << lv->logN) : 0u, 0u};
    if (lv->amax > MAX_ALPHA) return CKKS_E_BADARG;   // the extension kernels keep at most MAX_ALPHA digits per partition
    const int L = lv->L, K = lv->K, E = L + K, P = lv->nparts, N = 1 << lv->logN;
    int64_t* ext = ws;                                  // [P*E][N]
    int64_t* acc = ext + (long long)P * E * N;          // [2E][N]
    int64_t* eff = acc + 2ll * E * N;                   // [K][N]
    const MontPack m{lv->_2q, lv->ql, lv->qh, lv->kl, lv->kh};
    const bool fp64 = lv->Hm && lv->Pinv;
    // the evaluation key decides the order of the NTT domain in this stage: permuted copies <-> warp-interleaved data
    const int perm = keys_permuted ? 1 : 0;
    if (fp64) {
        // Slab pipeline over target limbs [t0, t1): extend -> column pass -> block pass -> inner product per slab, slabs
        // alternating between two internal streams so that the tail of one kernel overlaps the next slab's kernels.
        // Measured (profiles/r01_lab_notes.txt): L2-sized slabs (32-64 MB) lose more to small grids than they gain from
        // L2 residency of the extended block; two 100 MB slabs on two streams are the best setting at gold.
        ExtArgs X{};
        X.digit_ptrs = digit_ptrs ? digit_ptrs + pb : nullptr;
        X.d_stride = digit_stride;
        X.alphas = lv->part_alpha + pb;
        X.wide = lv->part_wide + pb;
        X.Hm = lv->Hm + pb;
        X.Rd = lv->Rd;
        X.C31 = lv->C31;
        X.Lenter = lv->Lenter + pb;
        X.Rs = lv->Rs;
        X.q = lv->q; X._2q = lv->_2q; X.ql = lv->ql; X.qh = lv->qh; X.kl = lv->kl; X.kh = lv->kh;
        X.out = ext + (long long)pb * E * N;
        X.E = E;
        X.N = N;
        X.raw = 1;   // scale-prime rows travel as raw doubles from the extension to the inverse transform
        X.qinv = lv->qinv;
        const long long slab_budget = (long long)g_slab_mb << 20;   // bytes of extended rows per slab
        const int Pr = pe - pb;                                      // partitions this call transforms
        const long long all_bytes = (long long)Pr * E * N * 8;
        int nslabs = (int)((all_bytes + slab_budget - 1) / slab_budget);
        if (nslabs < g_min_slabs && E >= 4 * g_min_slabs) nslabs = g_min_slabs;   // (one process per GPU: a 1/8 shard fits one slab)
        const int slab = (E + nslabs - 1) / nslabs;
        if (slab > EXT_MAX_E) return CKKS_E_BADARG;
        cudaStream_t main_st = S(stream);
        SidePipes* pipes = (nslabs > 1 && g_pipes > 1) ? side_pipes() : nullptr;
        const int npipes = pipes ? (g_pipes < nslabs ? g_pipes : nslabs) : 0;
        PipeScope scope;
        if (pipes) RC(scope.fork(pipes, main_st, npipes));
        int slab_no = 0;
        for (int t0 = 0; t0 < E; t0 += slab, ++slab_no) {
            cudaStream_t st = pipes ? pipes->s[slab_no % npipes] : main_st;
            const int t1 = (t0 + slab < E) ? t0 + slab : E;
            if (do_fwd) {
                const dim3 eg((N / 2 + 255) / 256, Pr);
                if (lv->amax <= 2) RC(launch_k(k_extend_fast<2>, eg, dim3(256), 0, st, X, t0, t1));
                else if (lv->amax <= 4) RC(launch_k(k_extend_fast<4>, eg, dim3(256), 0, st, X, t0, t1));
                else RC(launch_k(k_extend_fast<8>, eg, dim3(256), 0, st, X, t0, t1));
                FastArgs F = level_fast(lv, ext + (long long)pb * E * N, N, true, nullptr, nullptr, E);
                F.slab_rows = t1 - t0; F.group_rows = E; F.slab_t0 = t0;
                F.in_raw = 1; F.out_raw = 1;
                const dim3 grid(N / TILE, Pr * (t1 - t0));
                RC(launch_fast_col(true, F, grid, st));
                F.perm = perm;
                RC(launch_fast_block_any(true, F, grid, st));
            }
            if (!do_tail) continue;
            InnerArgs I{};
            I.ext = ext; I.k0 = k0_ptrs; I.k1 = k1_ptrs; I.k_stride = ksk_stride;
            I.acc0 = acc; I.acc1 = acc + (long long)E * N;
            I.q = lv->q; I._2q = lv->_2q; I.ql = lv->ql; I.qh = lv->qh; I.kl = lv->kl; I.kh = lv->kh;
            I.P = P; I.E = E; I.N = N; I.t0 = t0; I.raw = 1; I.qinv = lv->qinv;
            RC(launch_k(k_ksk_inner_fast, dim3((N / 2 + 255) / 256, t1 - t0), dim3(256), 0, st, I));
        }
        RC(scope.join());
    } else {
        if (perm) return CKKS_E_BADARG;   // the integer fall-back works on natural-order keys only
        if (do_fwd) {
            const int Pr = pe - pb;
            int64_t* extp = ext + (long long)pb * E * N;
            k_extend_batched<<<ew_grid(N, Pr * E), EW_THREADS, 0, S(stream)>>>(digit_ptrs + pb, digit_stride, lv->part_alpha + pb,
                                                                               extp, N, E, N, lv->Rs, lv->Lenter + pb, m);
            RC(launch_status());
            FastArgs F = level_fast(lv, extp, N, true, nullptr, nullptr, E);
            RC(fast_transform(true, F, Pr * E, S(stream)));
        }
        if (do_tail)
            RC(ckks_ksk_inner(ext, N, P, k0_ptrs, k1_ptrs, ksk_stride, acc, acc + (long long)E * N, N, E, N, lv->_2q, lv->ql,
                              lv->qh, lv->kl, lv->kh, stream));
    }
    if (!do_tail) return 0;
    // The two output polynomials are independent from here on: inverse transform + ModDown of each half run on their
    // own internal stream, so the small latency-bound ModDown kernels of one half overlap the transform of the other.
    int64_t* outs[2] = {out0, out1};
    const int64_t* adds[2] = {add0, add1};
    const int Ls = fp64 ? lv->L_small : 0;   // leading ordinary rows handled by the FP64 kernel
    SidePipes* tail = (g_pipes > 1 && g_split_tail) ? side_pipes() : nullptr;
    PipeScope tscope;
    if (tail) RC(tscope.fork(tail, S(stream), 2));
    for (int h = 0; h < 2; ++h) {
        cudaStream_t st = tail ? tail->s[h] : S(stream);
        int64_t* dh = acc + (long long)h * E * N;
        int64_t* effh = eff + (long long)h * K * N;
        if (tail || h == 0) {   // (one stream: both halves in one batched transform)
            FastArgs Fi = level_fast(lv, dh, N, false, lv->sExit, lv->sExit_sh, E);
            Fi.in_raw = fp64 ? 1 : 0;
            Fi.perm = perm;
            RC(fast_transform(false, Fi, tail ? E : 2 * E, st));
        }
        RC(launch_k(k_moddown_special, col_grid(N), dim3(EW_THREADS), 0, st, dh, N, L, K, N, lv->PiR, effh, m));
        if (Ls > 0) {
            ModDownArgs M{dh, effh, adds[h], add_stride, outs[h], out_stride, lv->Pinv, lv->C31, lv->q, L, K, E, N, lv->qinv,
                          add_ginv[h]};
            RC(launch_k(k_moddown_fast, dim3((N / 2 + 255) / 256, Ls), dim3(256), 0, st, M));
        }
        if (Ls < L) {
            RC(launch_k(k_moddown_ordinary, ew_grid(N, L - Ls), dim3(EW_THREADS), 0, st, dh, N, L, K, N, lv->Rs, lv->PiR, effh,
                        adds[h], add_stride, outs[h], out_stride, Ls, m, add_ginv[h]));
        }
    }
    RC(tscope.join());
    return 0;
}

int64_t ckks_exec_keyswitch_ws_elems(int L, int K, int nparts, int N) {
    const long long E = L + K;
    return ((long long)nparts * E + 2 * E + 2 * K) * N;   // extended block, two accumulators, ModDown scratch of both halves
}

int ckks_automorphism(const int64_t* in, int64_t is, int64_t* out, int64_t os, int C, int N, int64_t g, int canon,
                      const int64_t* _2q, void* stream) {
    CHECK_PTRS(in, out);
    if (C <= 0 || N <= 0 || (N & (N - 1)) || !(g & 1)) return CKKS_E_BADARG;
    if (canon && !_2q) return CKKS_E_BADARG;
    k_automorphism<<<dim3((N + EW_THREADS - 1) / EW_THREADS, C), EW_THREADS, 0, S(stream)>>>(
        in, is, out, os, N, (unsigned)(g & (2 * (int64_t)N - 1)), canon, _2q);
    return launch_status();
}

// ---- sampler (csprng.cuh): ChaCha20 counter mode, one block = four samples ----------------------------------------
static RngArgs rng_args(const uint32_t* key_nonce, const uint64_t* ctr_base, uint32_t* epoch, uint64_t inc, int L) {
    RngArgs A{};
    for (int i = 0; i < 10; ++i) A.key.w[i] = key_nonce[i];
    A.ctr_base = ctr_base;
    A.epoch = epoch;
    A.inc = inc;
    A.L = L;
    return A;
}
static dim3 rng_grid(int C, int L) { return dim3((L + 127) / 128, C); }

int ckks_rng_bytes(int64_t* out, int C, int L, const uint32_t* key_nonce, const uint64_t* ctr_base, uint32_t* epoch,
                   uint64_t inc, void* stream) {
    CHECK_PTRS(out, key_nonce, ctr_base, epoch);
    if (C <= 0 || L <= 0) return CKKS_E_BADARG;
    if (reinterpret_cast<uintptr_t>(out) & 31) return CKKS_E_ALIGN;
    k_rng_bytes<<<rng_grid(C, L), 128, 0, S(stream)>>>(out, rng_args(key_nonce, ctr_base, epoch, inc, L));
    return launch_status();
}

int ckks_rng_randint(int64_t* out, int C, int L, const uint64_t* q, int64_t shift, const uint32_t* key_nonce,
                     const uint64_t* ctr_base, uint32_t* epoch, uint64_t inc, void* stream) {
    CHECK_PTRS(out, q, key_nonce, ctr_base, epoch);
    if (C <= 0 || L <= 0) return CKKS_E_BADARG;
    if (reinterpret_cast<uintptr_t>(out) & 31) return CKKS_E_ALIGN;
    k_rng_randint<<<rng_grid(C, L), 128, 0, S(stream)>>>(out, q, shift, rng_args(key_nonce, ctr_base, epoch, inc, L));
    return launch_status();
}

int ckks_rng_gaussian(int64_t* out, int C, int L, const uint64_t* lut, int lut_size, int depth, const uint32_t* key_nonce,
                      const uint64_t* ctr_base, uint32_t* epoch, uint64_t inc, void* stream) {
    CHECK_PTRS(out, lut, key_nonce, ctr_base, epoch);
    if (C <= 0 || L <= 0 || lut_size <= 0 || 2 * lut_size > 128 || depth <= 0 || depth > 6) return CKKS_E_BADARG;
    if (reinterpret_cast<uintptr_t>(out) & 31) return CKKS_E_ALIGN;
    GaussLut T{};
    for (int i = 0; i < 2 * lut_size; ++i) T.v[i] = lut[i];   // HOST table: [low words | high words]
    T.size = lut_size;
    T.depth = depth;
    k_rng_gaussian<<<rng_grid(C, L), 128, 0, S(stream)>>>(out, T, rng_args(key_nonce, ctr_base, epoch, inc, L));
    return launch_status();
}

int ckks_rng_randround(const double* coef, int64_t* out, int n, const uint32_t* key_nonce, const uint64_t* ctr_base,
                       uint32_t* epoch, uint64_t inc, void* stream) {
    CHECK_PTRS(coef, out, key_nonce, ctr_base, epoch);
    if (n <= 0) return CKKS_E_BADARG;
    if (reinterpret_cast<uintptr_t>(out) & 31) return CKKS_E_ALIGN;
    const int L = (n + 15) / 16;
    k_rng_randround<<<rng_grid(1, L), 128, 0, S(stream)>>>(coef, out, n, rng_args(key_nonce, ctr_base, epoch, inc, L));
    return launch_status();
}

}  // extern "C"
