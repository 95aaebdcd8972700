// bfly_bench.cu -- register-only ceiling of the FP64 error-free butterfly (no memory, no barriers): each thread keeps
// 16 doubles and 15 twiddles in registers and runs radix-16 rounds (4 stages x 8 butterflies, 8 FP64 instr each) in a
// loop.  Reports butterflies/s and FP64 warp-instr/clk/SM for several occupancies, to separate "FP64 pipe behaviour"
// from "memory / synchronisation" in the NTT kernels.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a ...
#include <cstdio>
#include <cuda_runtime.h>
#define MAGIC 6755399441055744.0

__device__ __forceinline__ double mulmod(double v, double w, double q, double qinv) {
    const double p = __dmul_rn(v, w);
    const double e = __fma_rn(v, w, -p);
    const double m = __dadd_rn(__fma_rn(p, qinv, MAGIC), -MAGIC);
    return __dadd_rn(__fma_rn(-m, q, p), e);
}

template <int VERT>
__global__ void __launch_bounds__(256) k(double* out, int iters, double q) {
    double e[16], w[15];
    const double qinv = 1.0 / q;
    for (int i = 0; i < 16; ++i) e[i] = (double)((threadIdx.x * 16 + i) % 1000003);
    for (int i = 0; i < 15; ++i) w[i] = (double)((blockIdx.x * 15 + i * 7919 + 12345) % 1099511);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int d = 8 >> s;
            if (VERT) {
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
            } else {
#pragma unroll
                for (int kk = 0; kk < 16; ++kk)
                    if (!(kk & d)) {
                        const double r = mulmod(e[kk + d], w[(1 << s) - 1 + (kk >> (4 - s))], q, qinv);
                        const double u = e[kk];
                        e[kk] = __dadd_rn(u, r);
                        e[kk + d] = __dadd_rn(u, -r);
                    }
            }
        }
        // keep magnitudes bounded like the real kernel does between passes
#pragma unroll
        for (int i = 0; i < 16; ++i) { const double m = __dadd_rn(__fma_rn(e[i], qinv, MAGIC), -MAGIC); e[i] = __fma_rn(-m, q, e[i]); }
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int VERT>
void run(int ctas_per_sm, int sms, double* out, double ghz) {
    const int blocks = sms * ctas_per_sm, iters = 400;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<VERT><<<blocks, 256>>>(out, 10, 1099511799809.0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<VERT><<<blocks, 256>>>(out, iters, 1099511799809.0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bfly = (double)blocks * 256 * iters * 32;
    const double fp64 = (double)blocks * 8 * iters * (32 * 8 + 16 * 3);
    printf("%s  %d CTAs/SM (%2d warps/SM): %7.3f ms  %6.2f G bfly/s  %5.2f FP64 warp-inst/clk/SM  -> 2^16-point limb NTT compute floor %.3f us\n",
           VERT ? "vertical  " : "sequential", ctas_per_sm, ctas_per_sm * 8, ms, bfly / ms * 1e-6, fp64 / (ms * 1e-3) / sms / (ghz * 1e9),
           524288.0 / (bfly / ms * 1e-3));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    for (int c : {1, 2, 3, 4, 6, 8}) { run<0>(c, p.multiProcessorCount, out, khz * 1e-6); }
    for (int c : {1, 2, 3, 4, 6, 8}) { run<1>(c, p.multiProcessorCount, out, khz * 1e-6); }
    return 0;
}
