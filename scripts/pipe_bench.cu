// pipe_bench.cu -- instruction-throughput microbenchmark for the pipes the NTT butterflies live on
// (sm_100a): IMAD.WIDE.U32, IMAD (lo), IADD3, LOP3, DFMA/DADD/DMUL and mixes.  Prints warp-instructions
// per clock per SM.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/pipe_bench.bin scripts/pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int KIND>
__global__ void k(uint64_t* out, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u;
    uint64_t acc[ILP];
    double d[ILP];
    uint32_t r[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = a + i; d[i] = 1.0 + i * 1e-3; r[i] = b + i; }
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(r[i]));
            if (KIND == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (KIND == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a));
            if (KIND == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (KIND == 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
            if (KIND == 5) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c));
            if (KIND == 6) {  // 1 IMAD.WIDE + 1 IADD
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(r[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            }
            if (KIND == 7) {  // 1 IMAD.WIDE + 1 DFMA
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(b));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
            }
            if (KIND == 8) {  // 1 IMAD.WIDE + 1 IADD + 1 DFMA
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
            }
            if (KIND == 9) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"((uint64_t)a * 0x9e3779b97f4a7c15ull));
            if (KIND == 10) asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"((uint64_t)b));
            if (KIND == 11) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(a));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i] + (uint64_t)d[i] + r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, int per_iter, uint64_t* out, int sms, double ghz_hint) {
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<KIND><<<blocks, threads>>>(out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<KIND><<<blocks, threads>>>(out, 2);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)blocks * (threads / 32) * ITERS * ILP * per_iter;
    printf("%-28s %8.3f ms  %7.2f G warp-inst/s  %5.2f warp-inst/clk/SM @%.2f GHz\n", name, ms, winst / ms * 1e-6,
           winst / (ms * 1e-3) / sms / (ghz_hint * 1e9), ghz_hint);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    printf("%s, %d SMs, max clock %.3f GHz (rates assume max clock; real clock may be lower)\n", p.name, sms, ghz);
    uint64_t* out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    run<0>("IMAD.WIDE.U32", 1, out, sms, ghz);
    run<1>("IMAD (lo)", 1, out, sms, ghz);
    run<2>("IADD", 1, out, sms, ghz);
    run<3>("LOP3", 1, out, sms, ghz);
    run<11>("SHF", 1, out, sms, ghz);
    run<4>("DFMA", 1, out, sms, ghz);
    run<5>("DADD", 1, out, sms, ghz);
    run<6>("IMAD.WIDE + IADD", 2, out, sms, ghz);
    run<7>("IMAD.WIDE + DFMA", 2, out, sms, ghz);
    run<8>("IMAD.WIDE + IADD + DFMA", 3, out, sms, ghz);
    run<9>("mul.hi.u64 (multi-inst)", 1, out, sms, ghz);
    run<10>("add.u64 (2 inst)", 1, out, sms, ghz);
    return 0;
}
