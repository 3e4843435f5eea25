// IterKernel: per-output-pixel conjugate gradient on the sub-system selected by
// the acceptance radius.  Replaces IterKernel._iterative_wrapper / _extract_submatrix /
// conjugate_gradient (lakernel.py:397-443, 446-530, 548-590) and the relevant_matrix construction
// (lakernel.py:615-619).
//
// One CTA per output pixel a:
//   1. sel = { i : hypot(y_out[a]-y_in[i], x_out[a]-x_in[i]) < rho }  (ordered compaction, ascending i)
//   2. CG exactly as coded in the reference: x0 = 0, atol = |b| rtol, loop it < maxiter:
//        rho = r.r ; if sqrt(rho) < atol break ; if it>0: p = p*(rho/rho_prev) + r ; q = A_sel p ;
//        alpha = rho/(p.q) ; x += alpha p ; r -= alpha q.          No final residual check.
//      A_sel is never materialised: rows are gathered straight from the (L2-resident) n x n matrix,
//      one warp per row, lanes over the selected columns.
//   3. Ti[a, sel] = float32(x) (the reference's Ti is float32 from the start, lakernel.py:577), zeros elsewhere;
//      written as f64 holding the float32-rounded value so the downstream D/N/T kernels are shared with Cholesky.
// All reductions use a fixed thread->element mapping and fixed trees: iteration counts are reproducible.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int CT = 256;

__global__ void __launch_bounds__(CT) k_iter_cg(const double* __restrict__ AA, int lda, double diag_add,
                                                const double* __restrict__ mB, int ldb, int m, int n,
                                                const double* __restrict__ inx, const double* __restrict__ iny,
                                                const double* __restrict__ outx, const double* __restrict__ outy,
                                                double rho_acc, double rtol,
                                                int maxiter, double* __restrict__ Ti, int ldt,
                                                int* __restrict__ niter, int* __restrict__ nsel) {
    extern __shared__ __align__(16) double sm[];
    double* red = sm;           // 40
    double* r = sm + 40;        // n
    double* p = r + n;          // n
    double* q = p + n;          // n
    double* x = q + n;          // n
    int* sel = reinterpret_cast<int*>(x + n);  // n
    __shared__ int wcount[CT / 32];
    __shared__ int sbase;
    const int a = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double yo = outy[a], xo = outx[a];
    // ---- ordered compaction of the accepted input pixels ----
    if (tid == 0) sbase = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += CT) {
        const int i = i0 + tid;
        bool ok = false;
        if (i < n) ok = hypot(yo - iny[i], xo - inx[i]) < rho_acc;
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) wcount[warp] = __popc(bal);
        __syncthreads();
        int off = sbase;
        for (int w = 0; w < warp; w++) off += wcount[w];
        if (ok) sel[off + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < CT / 32; w++) t += wcount[w];
            sbase += t;
        }
        __syncthreads();
    }
    const int na = sbase;
    // ---- CG ----
    const double* brow = mB + (size_t)a * ldb;
    double nb2 = 0.0;
    for (int j = tid; j < na; j += CT) {
        const double bj = brow[sel[j]];
        r[j] = bj;
        p[j] = bj;
        x[j] = 0.0;
        nb2 += bj * bj;
    }
    nb2 = block_sum(nb2, red);
    const double atol = sqrt(nb2) * rtol;
    double rho_prev = 0.0;
    int nprod = 0;
    for (int it = 0; it < maxiter; it++) {
        double rho = 0.0;
        for (int j = tid; j < na; j += CT) rho += r[j] * r[j];
        rho = block_sum(rho, red);
        if (sqrt(rho) < atol) break;
        if (it > 0) {
            const double beta = rho / rho_prev;
            for (int j = tid; j < na; j += CT) p[j] = p[j] * beta + r[j];
        }
        __syncthreads();
        // q = A_sel p : warp per row
        double pq = 0.0;
        for (int i = warp; i < na; i += CT / 32) {
            const int gi = sel[i];
            const double* Arow = AA + (size_t)gi * lda;
            double s = 0.0;
            for (int j = lane; j < na; j += 32) {
                const int gj = sel[j];
                double v = Arow[gj];
                if (gj == gi) v += diag_add;
                s += v * p[j];
            }
            s = warp_sum(s);
            if (lane == 0) {
                q[i] = s;
                pq += p[i] * s;
            }
        }
        pq = block_sum(pq, red);
        nprod++;
        const double alpha = rho / pq;
        for (int j = tid; j < na; j += CT) {
            x[j] += alpha * p[j];
            r[j] -= alpha * q[j];
        }
        rho_prev = rho;
        __syncthreads();
    }
    __syncthreads();
    // ---- scatter (float32 rounding as in the reference) ----
    double* Trow = Ti + (size_t)a * ldt;
    for (int i = tid; i < n; i += CT) Trow[i] = 0.0;
    __syncthreads();
    for (int j = tid; j < na; j += CT) Trow[sel[j]] = (double)(float)x[j];
    if (tid == 0) {
        if (niter) niter[a] = nprod;
        if (nsel) nsel[a] = na;
    }
}

}  // namespace

int launch_iter_cg(const double* AA, int lda, double diag_add, const double* mB, int ldb, int m, int n,
                   const double* inx, const double* iny, const double* outx, const double* outy, double rho_acc,
                   double rtol, int maxiter, double* Ti, int ldt, int* niter, int* nsel, cudaStream_t s) {
    if (m <= 0 || n <= 0) return 0;
    const size_t smem = sizeof(double) * (40 + 4 * (size_t)n) + sizeof(int) * (size_t)n + 16;
    B200_REQUIRE(smem <= 220 * 1024, "iter_cg: n too large for the shared-memory CG vectors (n <= ~6200)");
    static bool done = false;
    if (!done) {
        B200_CUDA(cudaFuncSetAttribute(k_iter_cg, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        done = true;
    }
    prof_begin(PROF_ITER_CG, s);
    k_iter_cg<<<m, CT, smem, s>>>(AA, lda, diag_add, mB, ldb, m, n, inx, iny, outx, outy, rho_acc, rtol,
                                  maxiter, Ti, ldt, niter, nsel);
    prof_end(8.0 * m * (double)n * 2.0, s);  // bytes: mBhalf rows read + Ti rows written (A_sel gathers hit L2)
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace b200
