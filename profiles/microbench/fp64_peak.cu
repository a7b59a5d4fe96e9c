// FP64 peak microbenchmark: DFMA vs DMMA (m8n8k4, m16n8k4, m16n8k8, m16n8k16) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b){
    double c[16];
    #pragma unroll
    for(int i=0;i<16;i++) c[i]=threadIdx.x*1e-3+i;
    for(int it=0; it<iters; it++){
        #pragma unroll
        for(int i=0;i<16;i++) c[i]=fma(c[i],a,b);
    }
    double s=0; 
    #pragma unroll
    for(int i=0;i<16;i++) s+=c[i];
    if(s==12345.678) out[0]=s;
}
// m8n8k4: A 1 reg, B 1 reg, C 2 regs
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b){
    double c[8][2];
    #pragma unroll
    for(int i=0;i<8;i++){c[i][0]=threadIdx.x; c[i][1]=i;}
    for(int it=0; it<iters; it++){
        #pragma unroll
        for(int i=0;i<8;i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s=0;
    #pragma unroll
    for(int i=0;i<8;i++) s+=c[i][0]+c[i][1];
    if(s==12345.678) out[0]=s;
}
// m16n8k4: A 2 regs, B 1, C 4
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double a, double b){
    double c[4][4];
    #pragma unroll
    for(int i=0;i<4;i++) for(int j=0;j<4;j++) c[i][j]=threadIdx.x+i+j;
    for(int it=0; it<iters; it++){
        #pragma unroll
        for(int i=0;i<4;i++)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
    }
    double s=0;
    #pragma unroll
    for(int i=0;i<4;i++) for(int j=0;j<4;j++) s+=c[i][j];
    if(s==12345.678) out[0]=s;
}
// m16n8k8: A 4 regs, B 2, C 4
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b){
    double c[4][4];
    #pragma unroll
    for(int i=0;i<4;i++) for(int j=0;j<4;j++) c[i][j]=threadIdx.x+i+j;
    for(int it=0; it<iters; it++){
        #pragma unroll
        for(int i=0;i<4;i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
    double s=0;
    #pragma unroll
    for(int i=0;i<4;i++) for(int j=0;j<4;j++) s+=c[i][j];
    if(s==12345.678) out[0]=s;
}
// m16n8k16: A 8 regs, B 4, C 4
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b){
    double c[4][4];
    #pragma unroll
    for(int i=0;i<4;i++) for(int j=0;j<4;j++) c[i][j]=threadIdx.x+i+j;
    for(int it=0; it<iters; it++){
        #pragma unroll
        for(int i=0;i<4;i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s=0;
    #pragma unroll
    for(int i=0;i<4;i++) for(int j=0;j<4;j++) s+=c[i][j];
    if(s==12345.678) out[0]=s;
}
// mixed: DFMA + DMMA interleaved in the same warp
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b){
    double c[4][4]; double f[8];
    #pragma unroll
    for(int i=0;i<4;i++) for(int j=0;j<4;j++) c[i][j]=threadIdx.x+i+j;
    #pragma unroll
    for(int i=0;i<8;i++) f[i]=i+threadIdx.x;
    for(int it=0; it<iters; it++){
        #pragma unroll
        for(int i=0;i<4;i++){
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
            f[2*i]=fma(f[2*i],a,b); f[2*i+1]=fma(f[2*i+1],a,b);
        }
    }
    double s=0;
    #pragma unroll
    for(int i=0;i<4;i++) for(int j=0;j<4;j++) s+=c[i][j];
    #pragma unroll
    for(int i=0;i<8;i++) s+=f[i];
    if(s==12345.678) out[0]=s;
}
template<typename K> int run(const char* name, K kern, double fma_per_thread_iter, int iters, int blocks_per_sm){
    double* out; CK(cudaMalloc(&out, 8));
    int nsm=148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<nsm*blocks_per_sm,256>>>(out, iters/10, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    float best=1e30, tot=0; int reps=8;
    for(int r=0;r<reps;r++){
        cudaEventRecord(e0);
        kern<<<nsm*blocks_per_sm,256>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; tot+=ms;
    }
    double flops = 2.0*fma_per_thread_iter*iters*256.0*nsm*blocks_per_sm;
    printf("%-14s blocks/SM=%d  best %.3f ms  %.2f TFLOP/s (best)  %.2f TFLOP/s (avg of %d back-to-back)\n", name, blocks_per_sm, best, flops/best*1e-9, flops/(tot/reps)*1e-9, reps);
    cudaFree(out); return 0;
}
int main(){
    int iters=20000;
    for(int bps=1;bps<=4;bps*=2){
        run("dfma", k_dfma, 16, iters, bps);
        run("dmma m8n8k4", k_dmma884, 8*256/32.0, iters, bps);
        run("dmma m16n8k4", k_dmma1684, 4*512/32.0, iters, bps);
        run("dmma m16n8k8", k_dmma1688, 4*1024/32.0, iters, bps);
        run("dmma m16n8k16", k_dmma16816, 4*2048/32.0, iters/2, bps);
        run("mixed k8+dfma", k_mixed, 4*1024/32.0+8, iters, bps);
    }
    return 0;
}
