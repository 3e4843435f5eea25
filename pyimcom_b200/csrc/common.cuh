// Shared device helpers for the pyimcom_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b200 {

// ---- error plumbing: every C-ABI entry point returns 0 or a negative/CUDA error code ----------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what, const char* file, int line);

#define B200_CUDA(call)                                                   \
    do {                                                                  \
        int _rc = ::b200::check_cuda((call), #call, __FILE__, __LINE__);  \
        if (_rc) return _rc;                                              \
    } while (0)

// kernel-launch accounting (b200_launch_count): every launcher reports how many kernels it enqueued
extern long long g_launches;
#define B200_LAUNCH_CHECK()              \
    do {                                 \
        ::b200::g_launches += 1;         \
        B200_CUDA(cudaGetLastError());   \
    } while (0)
#define B200_LAUNCHED(k) (::b200::g_launches += (k))

#define B200_REQUIRE(cond, msg)                                           \
    do {                                                                  \
        if (!(cond)) {                                                    \
            ::b200::set_error("%s:%d: requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
            return -1;                                                    \
        }                                                                 \
    } while (0)

// per-launch device timing (bench.py's roofline): CUDA events around selected launches when enabled
enum ProfKind {
    PROF_CHOL_SUPER = 0, PROF_POTRF_DIAG, PROF_CHOL_PANEL, PROF_CHOL_INNER, PROF_BACK_SUPER, PROF_BACK_DIAG,
    PROF_BACK_INNER, PROF_BUILD_A, PROF_BUILD_B, PROF_FINALIZE, PROF_GEMM, PROF_ITER_CG, PROF_LAKERNEL1,
    PROF_EIGH, PROF_ASSEMBLE_A, PROF_OZ_SLICE, PROF_OZ_GEMM, PROF_NKINDS
};
void prof_begin(int kind, cudaStream_t st);
void prof_end(double work, cudaStream_t st);  // work: flops (tensor kinds) or bytes (HBM kinds) of the launch

// grow-only device scratch used by the host-pointer entry points (function seam)
int scratch(int slot, size_t bytes, void** out);
void scratch_release();

// ---- FP64 tensor-core step and asynchronous global -> shared copies (shared by the tile kernels) -------------
// D (8x8) += A (8x4, row) * B (4x8, col): lane (g = lane/4, q = lane%4) supplies A[g][q] and B[q][g] and owns
// D[g][2q], D[g][2q+1].  SASS: DMMA.8x8x4.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(double* smem, const double* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- small reductions --------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum (blockDim.x multiple of 32, <= 1024); result valid in every thread
__device__ __forceinline__ double block_sum(double v, double* red /* >= 33 doubles of smem */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = (lane < nw) ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

}  // namespace b200
