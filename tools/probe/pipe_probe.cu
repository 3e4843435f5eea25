// Do the FP64 tensor pipe (DMMA) and the FP64 FMA pipe (DFMA) run concurrently on sm_100a?
// Register-only loops: W_D warps per CTA issue DMMA.8x8x4, W_F warps issue DFMA; one CTA of 8 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1) probe(int wd, int iters, double* out) {
    const int warp = threadIdx.x >> 5;
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i][0] = acc[i][1] = threadIdx.x * 1e-9 + i;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    if (warp < wd) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) dmma(acc[i][0], acc[i][1], a, b);
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                acc[i][0] = fma(acc[i][0], a, b);
                acc[i][1] = fma(acc[i][1], b, a);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * nsm * 256);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("SMs %d\n", nsm);
    for (int wd = 0; wd <= 8; wd++) {
        probe<<<nsm, 256>>>(wd, iters, out);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        probe<<<nsm, 256>>>(wd, iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        // per warp: DMMA: iters*16 instr * 512 flop; DFMA: iters*32 instr * 64 flop
        const double fl_d = (double)nsm * wd * iters * 16 * 512.0, fl_f = (double)nsm * (8 - wd) * iters * 32 * 64.0;
        printf("dmma warps %d dfma warps %d : %.3f ms  DMMA %.2f TF/s  DFMA %.2f TF/s  total %.2f TF/s\n", wd, 8 - wd, ms,
               fl_d / ms / 1e9, fl_f / ms / 1e9, (fl_d + fl_f) / ms / 1e9);
    }
    return 0;
}
