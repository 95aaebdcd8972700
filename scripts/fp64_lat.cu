// fp64_lat.cu -- how much parallelism the B200 FP64 pipe needs: DFMA throughput (warp-inst/clk/SM) as a function of
// independent chains per warp (ILP) and warps per SM.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_lat.bin fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters) {
    double d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(double* out, int sms, int warps_per_sm, double ghz) {
    const int iters = 2048;
    const int threads = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
    const int blocks = sms * (warps_per_sm * 32 / threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<blocks, threads>>>(out, 16);
    cudaEventRecord(e0);
    k<ILP><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)blocks * (threads / 32) * iters * 8.0 * ILP;
    const double per_clk_sm = inst / (ms * 1e-3) / (ghz * 1e9) / sms;
    const double lat = (double)(ms * 1e-3) * ghz * 1e9 / (iters * 8.0);   // clk per dependent step of one warp
    printf("ILP %d  warps/SM %2d : %.2f warp-DFMA/clk/SM   (%.1f clk per dependent step)\n", ILP, warps_per_sm, per_clk_sm, lat);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate * 1e-6;
    double* out; cudaMalloc(&out, 148 * 2048 * 8);
    for (int w : {4, 8, 16, 24, 32}) { run<1>(out, p.multiProcessorCount, w, ghz); run<2>(out, p.multiProcessorCount, w, ghz); run<4>(out, p.multiProcessorCount, w, ghz); run<8>(out, p.multiProcessorCount, w, ghz); }
    return 0;
}
