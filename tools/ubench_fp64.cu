// fp64 micro-benchmark for B200: DFMA and DMMA (mma.sync m8n8k4 f64) throughput and dependent-issue latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench_fp64.cu -o tools/ubench_fp64
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dmma(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = 0.5 * i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s  SMs %d  clock %d kHz\n", p.name, sms, clk);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = warps * 32, blocks = sms * 2;
        float ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 8 * iters * (double)threads * blocks;
        printf("DFMA ILP8  %2d warps/blk x2 blk/SM : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 256 * 4 * iters * (double)warps * blocks;
        printf("DMMA ILP4  %2d warps/blk x2 blk/SM : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
    // dependent-issue latency: 1 warp, ILP 1
    {
        float ms = time_ms([&] { k_dfma<1><<<1, 32>>>(out, 200000, 1.0000001, 1e-9); });
        printf("DFMA dependent chain: %.2f ns/op\n", ms * 1e6 / 200000);
        ms = time_ms([&] { k_dmma<1><<<1, 32>>>(out, 200000, 1.0000001, 1e-9); });
        printf("DMMA dependent chain: %.2f ns/op\n", ms * 1e6 / 200000);
    }
    return 0;
}
