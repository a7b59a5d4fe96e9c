// DMMA.8x8x4 throughput vs. number of independent accumulator chains per warp and warps per SM.
// Answers: how much ILP/TLP does the FP64 tensor pipe need on B200?
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k(double* out, int iters, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
template <int NACC>
void run(int warps) {
    double* out; cudaMalloc(&out, 8);
    int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 40000 / NACC;
    k<NACC><<<nsm, warps * 32>>>(out, iters / 10, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<NACC><<<nsm, warps * 32>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dmma_per_smsp = (double)iters * NACC * warps / 4.0;   // warps spread over 4 SMSPs
    double cyc = ms * 1e-3 * 1.965e9;
    double tf = 2.0 * 256 * (double)iters * NACC * warps * nsm / ms * 1e-9;
    printf("acc/warp=%d warps/SM=%2d : %6.2f TFLOP/s  (%5.1f cycles per DMMA per SMSP; per-chain issue interval %6.1f cycles)\n",
           NACC, warps, tf, cyc / dmma_per_smsp, cyc / iters);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
