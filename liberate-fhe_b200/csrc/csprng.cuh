// csprng.cuh -- ChaCha20 counter-mode sampler kernels for sm_100a: uniform mod q, discrete Gaussian (CDT search tree),
// random rounding.  Replaces the reference's four csprng extensions (src/liberate/csprng/: chacha20_cuda_kernel.cu,
// randint_cuda_kernel.cu:23-102, discrete_gaussian_cuda_kernel.cu:27-108, randround_cuda_kernel.cu:8-37) for key
// generation and encryption.
//
// Same stream as the reference for the same (key, nonce): ChaCha20 block function (20 rounds), state words 0-3 the
// "expand 32-byte k" constants, 4-11 the key, 12-13 a 64-bit block counter, 14-15 the nonce; one block (16 x 32 bit)
// yields four 128-bit draws, each turned into one sample exactly as the reference does.  What differs is the
// machinery: the reference keeps a [blocks, 16] int64 STATE TENSOR in HBM (128 B per block, read and written by every
// call, 32-bit words held in int64, working state in shared memory).  Here a block's state is a pure function of
// (key, nonce, counter) -- the counter of block (channel c, position l) is ctr_base[c] + l + epoch * inc with a 4-byte
// epoch per block (how often that block was drawn, which is all the reference's state tensor really remembers) -- the
// 16 words live in registers, rotations are funnel shifts, and a thread's four samples leave as one 256-bit store.
#pragma once
#include <cstdint>

namespace ckks {

struct RngKey {
    uint32_t w[10];   // key[8], nonce[2]
};
struct GaussLut {
    uint64_t v[128];  // [low words of the tree nodes | high words], discrete_gaussian_sampler.py:96-116
    int size, depth;
};

#define CKKS_QR(a, b, c, d)                     \
    a += b; d ^= a; d = __funnelshift_l(d, d, 16); \
    c += d; b ^= c; b = __funnelshift_l(b, b, 12); \
    a += b; d ^= a; d = __funnelshift_l(d, d, 8);  \
    c += d; b ^= c; b = __funnelshift_l(b, b, 7);

__device__ __forceinline__ void chacha20_block(const RngKey& K, uint64_t ctr, uint32_t (&o)[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, K.w[0], K.w[1], K.w[2], K.w[3],
                      K.w[4],      K.w[5],      K.w[6],      K.w[7],      (uint32_t)ctr, (uint32_t)(ctr >> 32), K.w[8], K.w[9]};
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = s[i];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        CKKS_QR(o[0], o[4], o[8], o[12]);
        CKKS_QR(o[1], o[5], o[9], o[13]);
        CKKS_QR(o[2], o[6], o[10], o[14]);
        CKKS_QR(o[3], o[7], o[11], o[15]);
        CKKS_QR(o[0], o[5], o[10], o[15]);
        CKKS_QR(o[1], o[6], o[11], o[12]);
        CKKS_QR(o[2], o[7], o[8], o[13]);
        CKKS_QR(o[3], o[4], o[9], o[14]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] += s[i];
}

struct RngArgs {
    RngKey key;
    const uint64_t* ctr_base;   // [C] counter of position 0 of every channel of this call
    uint32_t* epoch;            // [C][L] draws so far of every block of this call (incremented)
    uint64_t inc;               // counter distance between two draws of the same block
    int L;                      // blocks per channel (N / 4)
};
__device__ __forceinline__ bool rng_block(const RngArgs& A, uint32_t (&x)[16], int& c, int& l) {
    c = blockIdx.y;
    l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= A.L) return false;
    uint32_t* ep = A.epoch + (long long)c * A.L + l;
    const uint32_t e = *ep;
    *ep = e + 1;
    chacha20_block(A.key, A.ctr_base[c] + (uint64_t)l + (uint64_t)e * A.inc, x);
    return true;
}
__device__ __forceinline__ void st4(int64_t* p, int64_t a, int64_t b, int64_t c, int64_t d) {
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// raw blocks: out[c][l][16] = the 16 words (chacha20_cuda_kernel.cu: dest)
__global__ void __launch_bounds__(128) k_rng_bytes(int64_t* __restrict__ out, const RngArgs A) {
    uint32_t x[16];
    int c, l;
    if (!rng_block(A, x, c, l)) return;
    int64_t* o = out + ((long long)c * A.L + l) * 16;
#pragma unroll
    for (int i = 0; i < 16; i += 4) st4(o + i, x[i], x[i + 1], x[i + 2], x[i + 3]);
}

// uniform in [0, q_c) + shift: floor(X q / 2^128) for the 128-bit draw X (randint_cuda_kernel.cu:60-101)
__global__ void __launch_bounds__(128) k_rng_randint(int64_t* __restrict__ out, const uint64_t* __restrict__ q, int64_t shift,
                                                     const RngArgs A) {
    uint32_t x[16];
    int c, l;
    if (!rng_block(A, x, c, l)) return;
    const uint64_t p = q[c];
    int64_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t lo = ((uint64_t)x[4 * i] << 32) | x[4 * i + 1], hi = ((uint64_t)x[4 * i + 2] << 32) | x[4 * i + 3];
        const uint64_t t = __umul64hi(p, lo);          // bits 64..127 of p * lo
        const uint64_t m = p * hi;                     // low half of p * hi (same weight)
        const uint64_t carry = (m + t < m) ? 1ull : 0ull;
        r[i] = (int64_t)(__umul64hi(p, hi) + carry) + shift;
    }
    st4(out + ((long long)c * A.L + l) * 4, r[0], r[1], r[2], r[3]);
}

// discrete Gaussian: constant-depth walk of the CDT search tree on a 127-bit draw, sign from the spare bit
// (discrete_gaussian_cuda_kernel.cu:63-107)
__global__ void __launch_bounds__(128) k_rng_gaussian(int64_t* __restrict__ out, const GaussLut T, const RngArgs A) {
    uint32_t x[16];
    int c, l;
    if (!rng_block(A, x, c, l)) return;
    int64_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t lo = ((uint64_t)x[4 * i] << 32) | x[4 * i + 1];
        uint64_t hi = ((uint64_t)x[4 * i + 2] << 32) | x[4 * i + 3];
        const int64_t sign = (int64_t)(hi & 1);
        hi >>= 1;
        int jump = 1, cur = 0, cnt = 0;
        for (int j = 0; j < T.depth; ++j) {
            const uint64_t th = T.v[cnt + cur + T.size], tl = T.v[cnt + cur];
            const int ge = (hi > th) | ((hi == th) & (lo >= tl));
            cur = 2 * cur + ge;
            cnt += jump;
            jump *= 2;
        }
        r[i] = (sign * 2 - 1) * (int64_t)cur;
    }
    st4(out + ((long long)c * A.L + l) * 4, r[0], r[1], r[2], r[3]);
}

// random rounding of n doubles with 32 random bits each: sign(x) (floor|x| + [u < frac|x| 2^32]) (randround_cuda_kernel.cu:8-37)
__global__ void __launch_bounds__(128) k_rng_randround(const double* __restrict__ coef, int64_t* __restrict__ out, int n,
                                                       const RngArgs A) {
    uint32_t x[16];
    int c, l;
    if (!rng_block(A, x, c, l)) return;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        int64_t r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long idx = (long long)l * 16 + i + k;
            const double v = (idx < n) ? coef[idx] : 0.0;
            const double a = fabs(v), fl = floor(a);
            const int64_t ifrac = __double2ll_rn((a - fl) * 4294967296.0);
            const int64_t rr = (int64_t)fl + (((int64_t)x[i + k] < ifrac) ? 1 : 0);
            r[k] = signbit(v) ? -rr : rr;
        }
        if ((long long)l * 16 + i + 3 < n) st4(out + (long long)l * 16 + i, r[0], r[1], r[2], r[3]);
        else
            for (int k = 0; k < 4; ++k)
                if ((long long)l * 16 + i + k < n) out[(long long)l * 16 + i + k] = r[k];
    }
}

}  // namespace ckks
