// pipe_mix_bench.cu -- register-only ceilings for the round-2 NTT design question: how fast is an INTEGER butterfly
// for the scale primes (q < 2^42) written for the IMAD pipe, and how much do integer warps and FP64 warps gain when
// they share an SM?  Each thread keeps 16 coefficients + 15 twiddles in registers and loops over radix-16 rounds.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/pipe_mix_bench.bin scripts/pipe_mix_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define MAGIC 6755399441055744.0

// ---- FP64 error-free butterfly (8 FP64 instructions), "vertical" issue order as in ntt_fast.cuh ----
__device__ __forceinline__ void f64_round(double (&e)[16], const double (&w)[15], double q, double qinv) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int d = 8 >> s;
        double p[8], r[8], m[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) { const int kk = ((b / d) * 2 * d) + (b % d); p[b] = __dmul_rn(e[kk + d], w[(1 << s) - 1 + (kk >> (4 - s))]); }
#pragma unroll
        for (int b = 0; b < 8; ++b) { const int kk = ((b / d) * 2 * d) + (b % d); r[b] = __fma_rn(e[kk + d], w[(1 << s) - 1 + (kk >> (4 - s))], -p[b]); }
#pragma unroll
        for (int b = 0; b < 8; ++b) m[b] = __fma_rn(p[b], qinv, MAGIC);
#pragma unroll
        for (int b = 0; b < 8; ++b) m[b] = __dadd_rn(m[b], -MAGIC);
#pragma unroll
        for (int b = 0; b < 8; ++b) p[b] = __fma_rn(-m[b], q, p[b]);
#pragma unroll
        for (int b = 0; b < 8; ++b) p[b] = __dadd_rn(p[b], r[b]);
#pragma unroll
        for (int b = 0; b < 8; ++b) { const int kk = ((b / d) * 2 * d) + (b % d); const double u = e[kk]; e[kk] = __dadd_rn(u, p[b]); e[kk + d] = __dadd_rn(u, -p[b]); }
    }
}

// ---- integer butterfly for q < 2^42, values lazily growing (no conditional corrections) ----
// Shoup: w' = floor(w 2^64 / q).  t ~ floor(v w' / 2^64) from three 32x32 products (the lo x lo product is dropped:
// t in [exact-1, exact]), r = v w - t q (low 64 bits) in [0, 3q).
struct IC { uint32_t nq0, nq1; uint64_t q3; };
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ uint64_t imulmod(uint64_t v, uint64_t w, uint64_t wp, const IC& c) {
    const uint32_t v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32);
    const uint32_t w0 = (uint32_t)w, w1 = (uint32_t)(w >> 32);
    const uint32_t p0 = (uint32_t)wp, p1 = (uint32_t)(wp >> 32);
    const uint32_t a = __umulhi(v0, p1);
    uint64_t S, t, r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(S) : "r"(v1), "r"(p0), "l"((uint64_t)a));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(v1), "r"(p1), "l"(S >> 32));
    const uint32_t t0 = (uint32_t)t, t1 = (uint32_t)(t >> 32);
    // r = v*w + t*(-q)  (mod 2^64)
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(v0), "r"(w0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(t0), "r"(c.nq0));
    uint32_t hi = (uint32_t)(r >> 32);
    hi = v0 * w1 + hi;
    hi = v1 * w0 + hi;
    hi = t0 * c.nq1 + hi;
    hi = t1 * c.nq0 + hi;
    return pack((uint32_t)r, hi);
}
__device__ __forceinline__ void int_round(uint64_t (&e)[16], const uint64_t (&w)[15], const uint64_t (&wp)[15], const IC& c) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int d = 8 >> s;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int kk = ((b / d) * 2 * d) + (b % d);
            const int wi = (1 << s) - 1 + (kk >> (4 - s));
            const uint64_t r = imulmod(e[kk + d], w[wi], wp[wi], c);
            const uint64_t u = e[kk];
            e[kk] = u + r;
            e[kk + d] = u + c.q3 - r;
        }
    }
}
__device__ __forceinline__ uint64_t ireduce(uint64_t v, uint64_t q, uint64_t qp, const IC& c) {   // -> [0, 3q)
    return imulmod(v, 1ull, qp, c);
}

// mode: 0 = all FP64, 1 = all integer, m >= 2: warp w is integer iff w % m == m - 1
__global__ void __launch_bounds__(256) k(double* out, int iters, uint64_t q, int mode) {
    const int warp = threadIdx.x >> 5;
    const bool is_int = mode == 1 || (mode >= 2 && (warp % mode) == mode - 1);
    if (!is_int) {
        double e[16], w[15];
        const double qd = (double)q, qinv = 1.0 / qd;
        for (int i = 0; i < 16; ++i) e[i] = (double)((threadIdx.x * 16 + i) % 1000003);
        for (int i = 0; i < 15; ++i) w[i] = (double)((blockIdx.x * 15 + i * 7919 + 12345) % 1099511);
        for (int it = 0; it < iters; ++it) {
            f64_round(e, w, qd, qinv);
#pragma unroll
            for (int i = 0; i < 16; ++i) { const double m = __dadd_rn(__fma_rn(e[i], qinv, MAGIC), -MAGIC); e[i] = __fma_rn(-m, qd, e[i]); }
        }
        double s = 0;
        for (int i = 0; i < 16; ++i) s += e[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        uint64_t e[16], w[15], wp[15];
        const uint64_t nq = 0ull - q;
        const IC c{(uint32_t)nq, (uint32_t)(nq >> 32), 3 * q};
        const uint64_t qp = (uint64_t)(((unsigned __int128)1 << 64) / q);
        for (int i = 0; i < 16; ++i) e[i] = (uint64_t)((threadIdx.x * 16 + i) % 1000003);
        for (int i = 0; i < 15; ++i) {
            w[i] = (uint64_t)((blockIdx.x * 15 + i * 7919 + 12345) % 1099511);
            wp[i] = (uint64_t)((((unsigned __int128)w[i]) << 64) / q);
        }
        for (int it = 0; it < iters; ++it) {
            int_round(e, w, wp, c);
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = ireduce(e[i], q, qp, c);
        }
        uint64_t s = 0;
        for (int i = 0; i < 16; ++i) s += e[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (double)s;
    }
}

// correctness of the integer butterfly against 128-bit arithmetic (host check of a device result)
__global__ void check_k(uint64_t* out, uint64_t q) {
    const uint64_t nq = 0ull - q;
    const IC c{(uint32_t)nq, (uint32_t)(nq >> 32), 3 * q};
    uint64_t v = (threadIdx.x + 1) * 0x9E3779B97F4A7C15ull >> 14;   // < 2^50
    uint64_t w = ((threadIdx.x + 7) * 0xD1B54A32D192ED03ull) % q;
    uint64_t wp = (uint64_t)((((unsigned __int128)w) << 64) / q);
    out[3 * threadIdx.x] = v;
    out[3 * threadIdx.x + 1] = w;
    out[3 * threadIdx.x + 2] = imulmod(v, w, wp, c);
}

void run(int mode, int ctas_per_sm, int sms, double* out, double ghz) {
    const int blocks = sms * ctas_per_sm, iters = 400;
    const uint64_t q = 1099511799809ull;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<blocks, 256>>>(out, 10, q, mode);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<blocks, 256>>>(out, iters, q, mode);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bfly = (double)blocks * 256 * iters * 32;
    printf("mode %d  %d CTAs/SM (%2d warps/SM): %7.3f ms  %7.2f G bfly/s  -> 2^16-point limb NTT compute floor %.3f us  (%.2f SM-clk per warp-butterfly)\n",
           mode, ctas_per_sm, ctas_per_sm * 8, ms, bfly / ms * 1e-6, 524288.0 / (bfly / ms * 1e-3),
           (ms * 1e-3) * ghz * 1e9 * sms / (bfly / 32));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    {
        uint64_t* d; cudaMalloc(&d, 3 * 256 * 8);
        const uint64_t q = 1099511799809ull;
        check_k<<<1, 256>>>(d, q);
        uint64_t h[3 * 256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0; uint64_t maxk = 0;
        for (int i = 0; i < 256; ++i) {
            const unsigned __int128 prod = (unsigned __int128)h[3 * i] * h[3 * i + 1];
            const uint64_t want = (uint64_t)(prod % q), got = h[3 * i + 2];
            if (got % q != want || got >= 3 * q) ++bad;
            if (got / q > maxk) maxk = got / q;
        }
        printf("integer butterfly check: %d bad of 256, max multiple of q in result %llu\n", bad, (unsigned long long)maxk);
    }
    for (int mode : {0, 1, 2, 3, 4})
        for (int c : {2, 4, 6, 8}) run(mode, c, p.multiProcessorCount, out, khz * 1e-6);
    return 0;
}
