// Per-output-pixel Lagrange-multiplier kernels:
//   lakernel1            routine.py:341-430   kappa bisection over the eigen-spectrum (EigenKernel multi-kappa)
//   eigen_single         lakernel.py:165-170  closed forms at one kappa (EigenKernel single-kappa)
//   lsolve_sps           routine.py:433-484   small in-place Cholesky solve
//   build_reduced_T      routine.py:487-588   bracket + 12 bisections in the space of nv Cholesky nodes
//   node_stats           lakernel.py:361-368  D_p, N_pq, E_pq of the nv node solutions
// All are streaming / latency kernels: one CTA (or one thread) per output pixel, deterministic
// reductions (fixed thread -> element mapping, fixed tree), so repeated runs are bit-identical.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int KT_THREADS = 256;

// ---- lakernel1 ---------------------------------------------------------------------------------
// One CTA per output pixel a.  The row P[a,:] is staged once in shared memory (when it fits) and
// re-used by the nbis+1 passes, so HBM traffic is 8*m*n read + 8*m*n written; lam stays in L1/L2.
template <bool ROW_IN_SMEM>
__global__ void __launch_bounds__(KT_THREADS) k_lakernel1(const double* __restrict__ lam,
                                                          const double* __restrict__ mPhalf, int ldp, int m, int n,
                                                          double C, double targetleak, double kCmin, double kCmax,
                                                          int nbis, double* __restrict__ kappa,
                                                          double* __restrict__ Sigma, double* __restrict__ UC,
                                                          double* __restrict__ T, int ldt, double smax) {
    extern __shared__ __align__(16) double sm[];
    double* red = sm;        // 33 doubles
    double* row = sm + 40;   // n doubles when ROW_IN_SMEM
    const int a = blockIdx.x;
    const double* P = mPhalf + (size_t)a * ldp;
    if (ROW_IN_SMEM) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) row[i] = P[i];
        __syncthreads();
        P = row;
    }
    double factor = sqrt(kCmax / kCmin);
    double kap = sqrt(kCmax * kCmin);
    for (int b = 0; b < nbis; b++) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double l = lam[i];
            const double var = P[i] / (l + kap);
            s2 += var * var;
            s1 += (l + 2.0 * kap) * var * var;
        }
        s1 = block_sum(s1, red);
        s2 = block_sum(s2, red);
        const double udc = 1.0 - s1 / C;
        factor = sqrt(factor);
        kap *= (udc > targetleak && s2 < smax) ? 1.0 / factor : factor;
    }
    double s1 = 0.0, s2 = 0.0;
    double* Trow = T + (size_t)a * ldt;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double l = lam[i];
        const double var = P[i] / (l + kap);
        Trow[i] = var;
        s2 += var * var;
        s1 += (l + 2.0 * kap) * var * var;
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        Sigma[a] = s2;
        kappa[a] = kap;
        UC[a] = 1.0 - s1 / C;
    }
}

// ---- EigenKernel single kappa: tt = P/(lam+kappa); Sigma = sum tt^2; UC = 1 - sum (lam+2k)/(lam+k)^2 P^2 / C
__global__ void __launch_bounds__(KT_THREADS) k_eigen_single(const double* __restrict__ lam,
                                                             const double* __restrict__ mPhalf, int ldp, int m, int n,
                                                             double C, double kap, double* __restrict__ Sigma,
                                                             double* __restrict__ UC, double* __restrict__ T, int ldt) {
    __shared__ double red[40];
    const int a = blockIdx.x;
    const double* P = mPhalf + (size_t)a * ldp;
    double* Trow = T + (size_t)a * ldt;
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double l = lam[i], p = P[i];
        const double d = l + kap;
        const double tt = p / d;
        Trow[i] = tt;
        s2 += tt * tt;
        s1 += (l + 2.0 * kap) / (d * d) * (p * p);
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        Sigma[a] = s2;
        UC[a] = 1.0 - s1 / C;
    }
}

// ---- EmpirKernel (lakernel.py:747-805): Ti[a, i] = max(rho - hypot(dy, dx), 0) / sum_i(...) -----------------------
// One CTA per output pixel; the row is formed twice (sum, then normalised store) so nothing is staged.  The sum runs
// over i in ascending order per thread and a fixed tree across threads (np.sum is pairwise: both are deterministic
// float64 sums of non-negative terms, equal to a few ulp).  Columns n .. ldt-1 are zero-filled (solver padding).
__global__ void __launch_bounds__(KT_THREADS) k_empir_T(const double* __restrict__ inx, const double* __restrict__ iny,
                                                        const double* __restrict__ outx, const double* __restrict__ outy,
                                                        int m, int n, int npad, double rho, double* __restrict__ T, int ldt) {
    __shared__ double red[40];
    const int a = blockIdx.x;
    double* Trow = T + (size_t)a * ldt;
    if (a >= m) {
        for (int i = threadIdx.x; i < npad; i += blockDim.x) Trow[i] = 0.0;
        return;
    }
    const double yo = outy[a], xo = outx[a];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += fmax(rho - hypot(yo - iny[i], xo - inx[i]), 0.0);
    s = block_sum(s, red);
    for (int i = threadIdx.x; i < npad; i += blockDim.x)
        Trow[i] = i < n ? fmax(rho - hypot(yo - iny[i], xo - inx[i]), 0.0) / s : 0.0;
}

// ---- lsolve_sps: unblocked in-place lower Cholesky + two substitutions (destroys A) -----------------
// Device helper used per thread for the tiny nv x nv systems of build_reduced_T.
__device__ __forceinline__ void lsolve_small(int N, double* A, double* x, const double* b, double* p1) {
    for (int i = 0; i < N; i++) {
        for (int j = 0; j < i; j++) {
            double s = 0.0;
            for (int k = 0; k < j; k++) s += A[i * N + k] * A[j * N + k];
            A[i * N + j] = (A[i * N + j] - s) / A[j * N + j];
        }
        double s = 0.0;
        for (int k = 0; k < i; k++) s += A[i * N + k] * A[i * N + k];
        A[i * N + i] = sqrt(A[i * N + i] - s);
    }
    for (int i = 0; i < N; i++) {
        double s = 0.0;
        for (int j = 0; j < i; j++) s += A[i * N + j] * p1[j];
        p1[i] = (b[i] - s) / A[i * N + i];
    }
    for (int i = N - 1; i >= 0; i--) {
        double s = 0.0;
        for (int j = i + 1; j < N; j++) s += A[j * N + i] * x[j];
        x[i] = (p1[i] - s) / A[i * N + i];
    }
}

// General-N lsolve_sps for the function seam (tested at N = 1089 by the reference, test_routine.py:148-156).
// One CTA; row i of the factor is produced by the whole block: thread j < i computes
// L[i][j] needs L[i][k<j] -> left-looking by columns instead: column j of L from columns < j.
__global__ void __launch_bounds__(1024) k_lsolve_sps(int N, double* __restrict__ A, double* __restrict__ x,
                                                     const double* __restrict__ b, double* __restrict__ p1) {
    __shared__ double red[40];
    __shared__ double sdiag;
    const int tid = threadIdx.x, nt = blockDim.x;
    // left-looking Cholesky by columns (same arithmetic per entry as the reference's row-ordered loops:
    // L[i][j] = (A[i][j] - sum_{k<j} L[i][k] L[j][k]) / L[j][j], sums taken in increasing k)
    for (int j = 0; j < N; j++) {
        if (tid == 0) {
            double s = 0.0;
            for (int k = 0; k < j; k++) s += A[(size_t)j * N + k] * A[(size_t)j * N + k];
            sdiag = sqrt(A[(size_t)j * N + j] - s);
            A[(size_t)j * N + j] = sdiag;
        }
        __syncthreads();
        const double d = sdiag;
        for (int i = j + 1 + tid; i < N; i += nt) {
            double s = 0.0;
            const double* Ai = A + (size_t)i * N;
            const double* Aj = A + (size_t)j * N;
            for (int k = 0; k < j; k++) s += Ai[k] * Aj[k];
            A[(size_t)i * N + j] = (Ai[j] - s) / d;
        }
        __syncthreads();
    }
    // forward substitution L p1 = b
    for (int i = 0; i < N; i++) {
        double s = 0.0;
        for (int j = tid; j < i; j += nt) s += A[(size_t)i * N + j] * p1[j];
        s = block_sum(s, red);
        if (tid == 0) p1[i] = (b[i] - s) / A[(size_t)i * N + i];
        __syncthreads();
    }
    // backward substitution L^T x = p1
    for (int i = N - 1; i >= 0; i--) {
        double s = 0.0;
        for (int j = i + 1 + tid; j < N; j += nt) s += A[(size_t)j * N + i] * x[j];
        s = block_sum(s, red);
        if (tid == 0) x[i] = (p1[i] - s) / A[(size_t)i * N + i];
        __syncthreads();
    }
}

// ---- build_reduced_T_wrap ---------------------------------------------------------------------------
// One thread per output pixel (nv <= 16 nodes; the nv x nv system lives in local memory).
constexpr int NVMAX = 16;
__global__ void __launch_bounds__(128) k_build_reduced_T(const double* __restrict__ Nflat,
                                                         const double* __restrict__ Dflat,
                                                         const double* __restrict__ Eflat,
                                                         const double* __restrict__ kappa, int nv, int m, double ucmin,
                                                         double smax, double* __restrict__ out_kappa,
                                                         double* __restrict__ out_Sigma, double* __restrict__ out_UC,
                                                         double* __restrict__ out_w, int* __restrict__ out_iv,
                                                         int* __restrict__ out_branch) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m) return;
    const int nv2 = nv * nv;
    const double* Na = Nflat + (size_t)a * nv2;
    const double* Ea = Eflat + (size_t)a * nv2;
    const double* Da = Dflat + (size_t)a * nv;
    double M2d[NVMAX * NVMAX], w[NVMAX], p1[NVMAX];
    // bracket: walk down from the top node while the leakage target is missed and the noise cap holds
    int iv = nv - 1;
    double UCv = ucmin * 10, S = smax / 10;
    while (iv > 0 && ucmin < UCv && smax > S) {
        iv -= 1;
        S = Na[iv * (nv + 1)];
        UCv = 1.0 - 2.0 * Da[iv] + Ea[iv * (nv + 1)];
    }
    double kappamid = sqrt(kappa[iv] * kappa[iv + 1]);
    double factor = pow(kappa[iv + 1] / kappa[iv], 0.25);
    int branch = 0;
    for (int it = 0; it < 12; it++) {
        for (int p = 0; p < nv; p++)
            for (int q = 0; q <= p; q++)
                M2d[p * nv + q] = __dadd_rn(Ea[p + nv * q], __dmul_rn(kappamid, Na[p + nv * q]));
        lsolve_small(nv, M2d, w, Da, p1);
        S = 0.0;
        for (int p = 0; p < nv; p++) {
            double s = 0.0;
            for (int q = 0; q < nv; q++) s += Na[p + nv * q] * w[q];
            S += s * w[p];
        }
        UCv = 1.0 - kappamid * S;
        for (int p = 0; p < nv; p++) UCv -= Da[p] * w[p];
        const bool down = (ucmin < UCv && smax > S);
        if (down) branch |= (1 << it);
        kappamid *= down ? 1.0 / factor : factor;
        factor = sqrt(factor);
    }
    for (int p = 0; p < nv; p++) out_w[(size_t)a * nv + p] = w[p];
    out_kappa[a] = kappamid;
    out_Sigma[a] = S;
    out_UC[a] = UCv;
    if (out_iv) out_iv[a] = iv;
    if (out_branch) out_branch[a] = branch;
}

// ---- node statistics ---------------------------------------------------------------------------------
// One CTA per output pixel: D_p = sum_i mB[a,i] Tpi[p,a,i];  N_pq = sum_i Tpi[p,a,i] Tpi[q,a,i];
// E_pq = D_q - kappa_p N_pq (q <= p), symmetrised; also D/C and E/C as handed to build_reduced_T.
// Epq_in != nullptr (IterKernel exact_UC, lakernel.py:709-716) supplies E instead of the closed form.
__global__ void __launch_bounds__(KT_THREADS) k_node_stats(const double* __restrict__ mB, int ldb,
                                                           const double* __restrict__ Tpi, int ldt, size_t strideT,
                                                           int nv, int m, int n, DiagNodes kn, double Cnorm,
                                                           double* __restrict__ Dp, double* __restrict__ Npq,
                                                           double* __restrict__ Epq, double* __restrict__ DpC,
                                                           double* __restrict__ EpqC, const double* __restrict__ Epq_in) {
    __shared__ double red[40];
    __shared__ double sD[NVMAX], sN[NVMAX * NVMAX];
    const int a = blockIdx.x;
    const double* b = mB + (size_t)a * ldb;
    for (int p = 0; p < nv; p++) {
        const double* tp = Tpi + p * strideT + (size_t)a * ldt;
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s += b[i] * tp[i];
        s = block_sum(s, red);
        if (threadIdx.x == 0) sD[p] = s;
        for (int q = 0; q <= p; q++) {
            const double* tq = Tpi + q * strideT + (size_t)a * ldt;
            double s2 = 0.0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) s2 += tp[i] * tq[i];
            s2 = block_sum(s2, red);
            if (threadIdx.x == 0) sN[p * nv + q] = sN[q * nv + p] = s2;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nv * nv; t += blockDim.x) {
        const int p = t / nv, q = t - p * nv;
        const int hi = p > q ? p : q, lo = p > q ? q : p;
        double e;
        if (Epq_in)
            e = Epq_in[(size_t)a * nv * nv + t];
        else
            e = __dadd_rn(sD[lo], -__dmul_rn(kn.v[hi], sN[hi * nv + lo]));
        Npq[(size_t)a * nv * nv + t] = sN[t];
        Epq[(size_t)a * nv * nv + t] = e;
        EpqC[(size_t)a * nv * nv + t] = e / Cnorm;
    }
    for (int p = threadIdx.x; p < nv; p += blockDim.x) {
        Dp[(size_t)a * nv + p] = sD[p];
        DpC[(size_t)a * nv + p] = sD[p] / Cnorm;
    }
}

// rowdot[a] = sum_i X[a,i] * Y[a,i]   (E_pq of the exact-UC path after the DMMA product Tpi[p] @ A)
__global__ void __launch_bounds__(KT_THREADS) k_rowdot(const double* __restrict__ X, int ldx,
                                                       const double* __restrict__ Y, int ldy, int n,
                                                       double* __restrict__ out, int ostride) {
    __shared__ double red[40];
    const int a = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += X[(size_t)a * ldx + i] * Y[(size_t)a * ldy + i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[(size_t)a * ostride] = s;
}

// single-kappa outputs: Sigma = N; UC = 1 - (kappa N + D)/C (lakernel.py:312-316) or, with the exact
// E = T A T^T, UC = 1 + (E - 2 D)/C (lakernel.py:643-645); the kappa map is the scalar.
__global__ void k_single_kappa_maps(const double* __restrict__ D, const double* __restrict__ N,
                                    const double* __restrict__ E, int m, double kap, double C,
                                    double* __restrict__ ko, double* __restrict__ So, double* __restrict__ Uo) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m) return;
    const double Nv = N[a], Dv = D[a];
    ko[a] = kap;
    So[a] = Nv;
    Uo[a] = E ? 1.0 + (E[a] - 2.0 * Dv) / C : 1.0 - (kap * Nv + Dv) / C;
}

__global__ void k_scale(const double* __restrict__ in, double scale, int m, double* __restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < m) out[a] = in[a] * scale;
}

}  // namespace

int launch_single_kappa_maps(const double* D, const double* N, const double* E, int m, double kappa, double C,
                             double* kappa_out, double* Sigma_out, double* UC_out, cudaStream_t s) {
    if (m <= 0) return 0;
    k_single_kappa_maps<<<(m + 255) / 256, 256, 0, s>>>(D, N, E, m, kappa, C, kappa_out, Sigma_out, UC_out);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_empir_T(const double* inx, const double* iny, const double* outx, const double* outy, int m, int mpad, int n,
                   int npad, double rho, double* T, int ldt, cudaStream_t s) {
    if (mpad <= 0 || npad <= 0) return 0;
    k_empir_T<<<mpad, KT_THREADS, 0, s>>>(inx, iny, outx, outy, m, n, npad, rho, T, ldt);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_scale(const double* in, double scale, int m, double* out, cudaStream_t s) {
    if (m <= 0) return 0;
    k_scale<<<(m + 255) / 256, 256, 0, s>>>(in, scale, m, out);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_lakernel1(const double* lam, const double* mPhalf, int ldp, int m, int n, double C, double targetleak,
                     double kCmin, double kCmax, int nbis, double* kappa, double* Sigma, double* UC, double* T, int ldt,
                     double smax, cudaStream_t s) {
    if (m <= 0) return 0;
    prof_begin(PROF_LAKERNEL1, s);
    const size_t smem_row = (40 + (size_t)n) * sizeof(double);
    if (smem_row <= 200 * 1024) {
        static bool done = false;
        if (!done) {
            B200_CUDA(cudaFuncSetAttribute(k_lakernel1<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            done = true;
        }
        k_lakernel1<true><<<m, KT_THREADS, smem_row, s>>>(lam, mPhalf, ldp, m, n, C, targetleak, kCmin, kCmax, nbis,
                                                           kappa, Sigma, UC, T, ldt, smax);
    } else {
        k_lakernel1<false><<<m, KT_THREADS, 40 * sizeof(double), s>>>(lam, mPhalf, ldp, m, n, C, targetleak, kCmin,
                                                                       kCmax, nbis, kappa, Sigma, UC, T, ldt, smax);
    }
    prof_end(8.0 * m * (double)n * 2.0, s);  // bytes: mPhalf read once (row kept on chip), T written once
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_eigen_single(const double* lam, const double* mPhalf, int ldp, int m, int n, double C, double kappa,
                        double* Sigma, double* UC, double* T, int ldt, cudaStream_t s) {
    if (m <= 0) return 0;
    k_eigen_single<<<m, KT_THREADS, 0, s>>>(lam, mPhalf, ldp, m, n, C, kappa, Sigma, UC, T, ldt);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_lsolve_sps(int N, double* A, double* x, const double* b, double* work, cudaStream_t s) {
    if (N <= 0) return 0;
    k_lsolve_sps<<<1, 1024, 0, s>>>(N, A, x, b, work);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_build_reduced_T(const double* Nflat, const double* Dflat, const double* Eflat, const double* kappa, int nv,
                           int m, double ucmin, double smax, double* out_kappa, double* out_Sigma, double* out_UC,
                           double* out_w, int* out_iv, int* out_branch, cudaStream_t s) {
    if (m <= 0) return 0;
    B200_REQUIRE(nv >= 2 && nv <= NVMAX, "build_reduced_T supports 2..16 kappa nodes");
    k_build_reduced_T<<<(m + 127) / 128, 128, 0, s>>>(Nflat, Dflat, Eflat, kappa, nv, m, ucmin, smax, out_kappa,
                                                      out_Sigma, out_UC, out_w, out_iv, out_branch);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_node_stats(const double* mB, int ldb, const double* Tpi, int ldt, size_t strideT, int nv, int m, int n,
                      const double* kappa_nodes, double Cnorm, double* Dp, double* Npq, double* Epq, double* DpC,
                      double* EpqC, const double* Epq_in, cudaStream_t s) {
    if (m <= 0) return 0;
    B200_REQUIRE(nv >= 1 && nv <= NVMAX, "node_stats supports up to 16 kappa nodes");
    DiagNodes kn;
    for (int i = 0; i < NVMAX; i++) kn.v[i] = i < nv ? kappa_nodes[i] : 0.0;
    k_node_stats<<<m, KT_THREADS, 0, s>>>(mB, ldb, Tpi, ldt, strideT, nv, m, n, kn, Cnorm, Dp, Npq, Epq, DpC, EpqC,
                                          Epq_in);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_rowdot(const double* X, int ldx, const double* Y, int ldy, int m, int n, double* out, int ostride,
                  cudaStream_t s) {
    if (m <= 0) return 0;
    k_rowdot<<<m, KT_THREADS, 0, s>>>(X, ldx, Y, ldy, n, out, ostride);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace b200
