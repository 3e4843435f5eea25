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
// All reductions use a fixed thread->element mapping and fixed trees: iteration counts are reproducible.  The
// element-wise updates round the product before the sum, as NumPy / Numba do (no FMA contraction).
#include <stdlib.h>

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
                                                int* __restrict__ niter, int* __restrict__ nsel,
                                                const int* __restrict__ only /* nullable: process a only if only[a] */,
                                                int ncap /* capacity of the per-pixel vectors (accepted inputs) */) {
    extern __shared__ __align__(16) double sm[];
    if (only && !only[blockIdx.x]) return;
    double* red = sm;           // 40
    double* r = sm + 40;        // ncap
    double* p = r + ncap;       // ncap
    double* q = p + ncap;       // ncap
    double* x = q + ncap;       // ncap
    int* sel = reinterpret_cast<int*>(x + ncap);  // ncap
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
        if (ok) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            if (pos < ncap) sel[pos] = i;
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < CT / 32; w++) t += wcount[w];
            sbase += t;
        }
        __syncthreads();
    }
    const int na = sbase;
    if (na > ncap) {  // more accepted input pixels than the shared-memory vectors hold: reported, never truncated
        if (tid == 0 && niter) niter[a] = -1;
        return;
    }
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
            for (int j = tid; j < na; j += CT) p[j] = __dadd_rn(__dmul_rn(p[j], beta), r[j]);  // p *= beta; p += r (two roundings)
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
            x[j] = __dadd_rn(x[j], __dmul_rn(alpha, p[j]));  // x += alpha * p, r -= alpha * q: NumPy rounds the
            r[j] = __dsub_rn(r[j], __dmul_rn(alpha, q[j]));  // product before the sum (no FMA contraction)
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


// ---- tile form: TP adjacent output pixels per CTA share one gather of A ----------------------------------------
// Neighbouring output pixels accept almost the same input pixels, so the per-pixel kernel above streams nearly the
// same n_a x n_a sub-matrix through L2 once per pixel and iteration.  Here a CTA takes TP consecutive output pixels,
// forms the UNION U of their accepted sets once, and runs the TP conjugate-gradient recurrences side by side on
// vectors that live on U and are zero outside each pixel's own set: q_c = mask_c .* (A_U p_c).  Every row of A_U is
// then gathered once per iteration for all TP systems.  Padding a sum with exact zeros does not change it, so each
// recurrence is the reference's recurrence on its own sub-system (only the lane each term lands in differs from
// the per-pixel kernel); a pixel that has converged (sqrt(rho) < atol, lakernel.py:428) is frozen.
constexpr int TP = 4;

__device__ __forceinline__ void block_sum4(double (&v)[TP], double* red /* >= 8 * TP doubles */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < TP; c++) v[c] = warp_sum(v[c]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < TP; c++) red[wid * TP + c] = v[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < TP; c++) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < CT / 32; w++) t += red[w * TP + c];
        v[c] = t;
    }
}

__global__ void __launch_bounds__(CT) k_iter_cg_tile(const double* __restrict__ AA, int lda, double diag_add,
                                                     const double* __restrict__ mB, int ldb, int m, int n,
                                                     const double* __restrict__ inx, const double* __restrict__ iny,
                                                     const double* __restrict__ outx, const double* __restrict__ outy,
                                                     double rho_acc, double rtol, int maxiter, double* __restrict__ Ti,
                                                     int ldt, int* __restrict__ niter, int* __restrict__ nsel, int namax,
                                                     int* __restrict__ redo /* [m] out: 1 = union too large */) {
    extern __shared__ __align__(16) double sm[];
    double* red = sm;                        // 40
    double* r = sm + 40;                     // [namax][TP]
    double* p = r + (size_t)namax * TP;      // [namax][TP]
    double* q = p + (size_t)namax * TP;      // [namax][TP]
    double* x = q + (size_t)namax * TP;      // [namax][TP]
    int* sel = reinterpret_cast<int*>(x + (size_t)namax * TP);          // [namax]
    unsigned char* mk = reinterpret_cast<unsigned char*>(sel + namax);  // [namax] bit c: pixel c accepts this input
    __shared__ int wcount[CT / 32];
    __shared__ int sbase;
    const int a0 = blockIdx.x * TP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double yo[TP], xo[TP];
#pragma unroll
    for (int c = 0; c < TP; c++) {
        const int a = min(a0 + c, m - 1);
        yo[c] = outy[a];
        xo[c] = outx[a];
    }
    // ---- ordered compaction of the union of the accepted sets ----
    if (tid == 0) sbase = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += CT) {
        const int i = i0 + tid;
        unsigned mask = 0;
        if (i < n) {
            const double yi = iny[i], xi = inx[i];
#pragma unroll
            for (int c = 0; c < TP; c++)
                if (a0 + c < m && hypot(yo[c] - yi, xo[c] - xi) < rho_acc) mask |= 1u << c;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, mask != 0);
        if (lane == 0) wcount[warp] = __popc(bal);
        __syncthreads();
        int off = sbase;
        for (int w = 0; w < warp; w++) off += wcount[w];
        if (mask) {
            const int k = off + __popc(bal & ((1u << lane) - 1u));
            if (k < namax) {
                sel[k] = i;
                mk[k] = (unsigned char)mask;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < CT / 32; w++) t += wcount[w];
            sbase += t;
        }
        __syncthreads();
    }
    const int na = sbase;
    if (tid < TP && a0 + tid < m) redo[a0 + tid] = na > namax;
    if (na > namax) return;  // the union does not fit: these pixels go through the per-pixel kernel afterwards
    // ---- TP conjugate-gradient recurrences ----
    double nb2[TP], atol[TP], rho_prev[TP];
    int nprod[TP], ncnt[TP];
    bool done[TP];
#pragma unroll
    for (int c = 0; c < TP; c++) {
        nb2[c] = 0.0;
        rho_prev[c] = 0.0;
        nprod[c] = 0;
        ncnt[c] = 0;
        done[c] = a0 + c >= m;
    }
    for (int j = tid; j < na; j += CT) {
        const int gj = sel[j];
        const unsigned mj = mk[j];
#pragma unroll
        for (int c = 0; c < TP; c++) {
            const double bj = (mj >> c & 1u) ? mB[(size_t)min(a0 + c, m - 1) * ldb + gj] : 0.0;
            r[j * TP + c] = bj;
            p[j * TP + c] = bj;
            x[j * TP + c] = 0.0;
            nb2[c] += bj * bj;
            ncnt[c] += (int)(mj >> c & 1u);
        }
    }
    block_sum4(nb2, red);
#pragma unroll
    for (int c = 0; c < TP; c++) atol[c] = sqrt(nb2[c]) * rtol;
    for (int it = 0; it < maxiter; it++) {
        double rho[TP];
#pragma unroll
        for (int c = 0; c < TP; c++) rho[c] = 0.0;
        for (int j = tid; j < na; j += CT) {
#pragma unroll
            for (int c = 0; c < TP; c++) rho[c] += r[j * TP + c] * r[j * TP + c];
        }
        block_sum4(rho, red);
        bool any = false;
#pragma unroll
        for (int c = 0; c < TP; c++) {
            if (!done[c] && sqrt(rho[c]) < atol[c]) done[c] = true;  // lakernel.py:428
            any = any || !done[c];
        }
        if (!any) break;
        if (it > 0) {
            double beta[TP];
#pragma unroll
            for (int c = 0; c < TP; c++) beta[c] = done[c] ? 0.0 : rho[c] / rho_prev[c];
            for (int j = tid; j < na; j += CT) {
#pragma unroll
                for (int c = 0; c < TP; c++)
                    if (!done[c]) p[j * TP + c] = __dadd_rn(__dmul_rn(p[j * TP + c], beta[c]), r[j * TP + c]);
            }
        }
        __syncthreads();
        // q = mask .* (A_U p): one warp per row, every gathered entry of A serves all TP systems
        double pq[TP];
#pragma unroll
        for (int c = 0; c < TP; c++) pq[c] = 0.0;
        for (int i = warp; i < na; i += CT / 32) {
            const int gi = sel[i];
            const double* Arow = AA + (size_t)gi * lda;
            double s[TP];
#pragma unroll
            for (int c = 0; c < TP; c++) s[c] = 0.0;
            for (int j = lane; j < na; j += 32) {
                const int gj = sel[j];
                double v = __ldg(Arow + gj);
                if (gj == gi) v += diag_add;
                const double2 p01 = *reinterpret_cast<const double2*>(p + j * TP);
                const double2 p23 = *reinterpret_cast<const double2*>(p + j * TP + 2);
                s[0] += v * p01.x;
                s[1] += v * p01.y;
                s[2] += v * p23.x;
                s[3] += v * p23.y;
            }
#pragma unroll
            for (int c = 0; c < TP; c++) s[c] = warp_sum(s[c]);
            if (lane == 0) {
                const unsigned mi = mk[i];
#pragma unroll
                for (int c = 0; c < TP; c++) {
                    const double qc = (mi >> c & 1u) ? s[c] : 0.0;
                    q[i * TP + c] = qc;
                    pq[c] += p[i * TP + c] * qc;
                }
            }
        }
        block_sum4(pq, red);
        double alpha[TP];
#pragma unroll
        for (int c = 0; c < TP; c++) {
            alpha[c] = done[c] ? 0.0 : rho[c] / pq[c];
            if (!done[c]) {
                nprod[c]++;
                rho_prev[c] = rho[c];
            }
        }
        for (int j = tid; j < na; j += CT) {
#pragma unroll
            for (int c = 0; c < TP; c++)
                if (!done[c]) {
                    x[j * TP + c] = __dadd_rn(x[j * TP + c], __dmul_rn(alpha[c], p[j * TP + c]));
                    r[j * TP + c] = __dsub_rn(r[j * TP + c], __dmul_rn(alpha[c], q[j * TP + c]));
                }
        }
        __syncthreads();
    }
    __syncthreads();
    // ---- scatter (float32 rounding as in the reference, lakernel.py:577) ----
#pragma unroll
    for (int c = 0; c < TP; c++) {
        if (a0 + c >= m) continue;
        double* Trow = Ti + (size_t)(a0 + c) * ldt;
        for (int i = tid; i < n; i += CT) Trow[i] = 0.0;
    }
    __syncthreads();
    for (int j = tid; j < na; j += CT) {
        const unsigned mj = mk[j];
#pragma unroll
        for (int c = 0; c < TP; c++)
            if (a0 + c < m && (mj >> c & 1u)) Ti[(size_t)(a0 + c) * ldt + sel[j]] = (double)(float)x[j * TP + c];
    }
    double cnt[TP];
#pragma unroll
    for (int c = 0; c < TP; c++) cnt[c] = (double)ncnt[c];
    block_sum4(cnt, red);
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < TP; c++)
            if (a0 + c < m) {
                if (niter) niter[a0 + c] = nprod[c];
                if (nsel) nsel[a0 + c] = (int)(cnt[c] + 0.5);
            }
    }
}

}  // namespace

constexpr int ITER_NCAP = 6100;

int launch_iter_cg(const double* AA, int lda, double diag_add, const double* mB, int ldb, int m, int n,
                   const double* inx, const double* iny, const double* outx, const double* outy, double rho_acc,
                   double rtol, int maxiter, double* Ti, int ldt, int* niter, int* nsel, cudaStream_t s) {
    if (m <= 0 || n <= 0) return 0;
    // per-pixel kernel: four CG vectors + the index list of the ACCEPTED input pixels (those within rho_acc of the output
    // pixel) in shared memory; at most ITER_NCAP of them (a pixel that accepts more gets niter = -1 and the host raises)
    const int ncap = n < ITER_NCAP ? n : ITER_NCAP;
    const size_t smem1 = sizeof(double) * (40 + 4 * (size_t)ncap) + sizeof(int) * (size_t)ncap + 16;
    // tile kernel: union of TP accepted sets, at most namax entries (4 vectors x TP doubles + index + mask per entry)
    int namax = (n + 7) / 8 * 8;
    if (namax > 1400) namax = 1400;
    const size_t smem_t = sizeof(double) * (40 + 4 * (size_t)namax * TP) + (sizeof(int) + 1) * (size_t)namax + 64;
    static bool done = false;
    if (!done) {
        B200_CUDA(cudaFuncSetAttribute(k_iter_cg, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        B200_CUDA(cudaFuncSetAttribute(k_iter_cg_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        done = true;
    }
    void* flg = nullptr;
    if (int rc = scratch(8, sizeof(int) * (size_t)m, &flg)) return rc;
    int* redo = (int*)flg;
    const bool tiled = getenv("B200_ITER_PER_PIXEL") == nullptr;
    prof_begin(PROF_ITER_CG, s);
    if (tiled) {
        k_iter_cg_tile<<<(m + TP - 1) / TP, CT, smem_t, s>>>(AA, lda, diag_add, mB, ldb, m, n, inx, iny, outx, outy,
                                                             rho_acc, rtol, maxiter, Ti, ldt, niter, nsel, namax, redo);
        B200_LAUNCHED(1);
    }
    // per-pixel kernel: every pixel (tiling off) or only those whose tile overflowed
    k_iter_cg<<<m, CT, smem1, s>>>(AA, lda, diag_add, mB, ldb, m, n, inx, iny, outx, outy, rho_acc, rtol, maxiter, Ti,
                                   ldt, niter, nsel, tiled ? redo : nullptr, ncap);
    prof_end(8.0 * m * (double)n * 2.0, s);  // bytes: mBhalf rows read + Ti rows written (A_sel gathers hit L2)
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace b200
