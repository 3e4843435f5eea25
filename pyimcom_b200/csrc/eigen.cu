// Symmetric eigendecomposition for EigenKernel and the Cholesky repair branch:
// replaces np.linalg.eigh at lakernel.py:162, 201, 266.
//
// Parallel two-sided (classical) Jacobi in implicit form: the kernel keeps G = V^T A and V^T (V = I
// initially) and never stores H = V^T A V; the three entries a rotation of the pair (p, q) needs are dot
// products of rows, h_pp = g_p . v_p, h_qq = g_q . v_q, h_pq = g_p . v_q, and the rotation that zeroes h_pq
// is applied to rows p, q of both G and V^T (which is H <- J^T H J).  Working on H rather than on the Gram
// matrix G G^T = V^T A^2 V (one-sided Hestenes) keeps the absolute accuracy at eps |A|_F for the small
// eigenvalues too, which is what LAPACK's eigh delivers and what 1/(lam + kappa) in EigenKernel needs.
// Pairs are scheduled by the round-robin ("circle") tournament: npl/2 disjoint pairs per round, npl - 1
// rounds per sweep, one kernel launch per round, one CTA per pair.
// The pair (p,q) is skipped when |h_pq| <= tol sqrt|h_pp h_qq| (tol = max(1e-15, sqrt(n) eps)) or when
// |h_pq| <= 2 sqrt(n) eps |A|_F (the rounding floor of the implicit dot products): any orthonormal basis of a
// numerically degenerate subspace serves the callers (T, Sigma and U/C are invariant to it).
// Every reduction uses a fixed thread -> element mapping and a fixed tree: the result is deterministic.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int JT = 256;

// three simultaneous deterministic block sums
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* red /* >= 3*9 doubles */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    __syncthreads();
    if (lane == 0) {
        red[wid] = a;
        red[8 + wid] = b;
        red[16 + wid] = c;
    }
    __syncthreads();
    double ta = 0.0, tb = 0.0, tc = 0.0;
#pragma unroll
    for (int w = 0; w < JT / 32; w++) {
        ta += red[w];
        tb += red[8 + w];
        tc += red[16 + w];
    }
    a = ta;
    b = tb;
    c = tc;
}

__global__ void k_jacobi_init(double* __restrict__ Vt, int ldv, int n, const double* __restrict__ G, int lda,
                              double* __restrict__ fro2) {
    __shared__ double red[40];
    const int r = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        Vt[(size_t)r * ldv + c] = (r == c) ? 1.0 : 0.0;
        const double v = G[(size_t)r * lda + c];
        s += v * v;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) fro2[r] = s;
}

// sum of the per-row squared norms -> state[0] = |A|_F^2 (single CTA, deterministic)
__global__ void k_jacobi_fro(const double* __restrict__ fro2, int n, double* __restrict__ state) {
    __shared__ double red[40];
    double s = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) s += fro2[c];
    s = block_sum(s, red);
    if (threadIdx.x == 0) state[0] = s;
}

__global__ void __launch_bounds__(JT) k_jacobi_round(double* __restrict__ G, int lda, double* __restrict__ Vt, int ldv,
                                                     int n, int npl, int round, double tol,
                                                     const double* __restrict__ state, int* __restrict__ nrot) {
    __shared__ double red[32];
    // circle method: player npl-1 is fixed, the others rotate
    const int k = blockIdx.x, mod = npl - 1;
    int p, q;
    if (k == 0) {
        p = npl - 1;
        q = round % mod;
    } else {
        p = (round + k) % mod;
        q = (round - k + mod) % mod;
    }
    if (p > q) {
        const int t = p;
        p = q;
        q = t;
    }
    if (q >= n) return;  // dummy player of an odd-sized problem
    double* gp = G + (size_t)p * lda;
    double* gq = G + (size_t)q * lda;
    double* vp = Vt + (size_t)p * ldv;
    double* vq = Vt + (size_t)q * ldv;
    double a = 0.0, b = 0.0, g = 0.0;
    for (int c = threadIdx.x; c < n; c += JT) {
        const double x = gp[c], y = gq[c], u = vp[c], w = vq[c];
        a += x * u;
        b += y * w;
        g += x * w;
    }
    block_sum3(a, b, g, red);
    const double ab = sqrt(fabs(a)) * sqrt(fabs(b));
    const double floor1 = 2.0 * 2.220446049250313e-16 * sqrt((double)n * state[0]);  // 2 sqrt(n) eps |A|_F
    if (!(fabs(g) > tol * ab) || fabs(g) <= floor1) return;
    const double zeta = (b - a) / (2.0 * g);
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
    double *op = gp, *oq = gq, *wp = vp, *wq = vq;
    for (int c = threadIdx.x; c < n; c += JT) {
        const double x = gp[c], y = gq[c];
        op[c] = cs * x - sn * y;
        oq[c] = sn * x + cs * y;
        const double u = vp[c], w = vq[c];
        wp[c] = cs * u - sn * w;
        wq[c] = sn * u + cs * w;
    }
    if (threadIdx.x == 0) atomicAdd(nrot, 1);
}

__global__ void __launch_bounds__(JT) k_jacobi_finish(const double* __restrict__ G, int lda,
                                                      const double* __restrict__ Vt, int ldv, int n,
                                                      double* __restrict__ lam) {
    __shared__ double red[40];
    const int r = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < n; c += JT) s += G[(size_t)r * lda + c] * Vt[(size_t)r * ldv + c];
    s = block_sum(s, red);
    if (threadIdx.x == 0) lam[r] = s;
}

}  // namespace

int launch_jacobi_eigh(double* A, int lda, int n, double* Vt, int ldv, double* lam, int max_sweeps, int* sweeps_done,
                       cudaStream_t st) {
    if (sweeps_done) *sweeps_done = 0;
    if (n <= 0) return 0;
    B200_REQUIRE(lda >= n && ldv >= n, "eigh: leading dimensions too small");
    // device state: [0] |A|_F^2 ; then n per-row norms ; then the rotation counter
    void* wsv = nullptr;
    if (int rc = scratch(7, sizeof(double) * (size_t)(n + 2) + 64, &wsv)) return rc;
    double* state = (double*)wsv;
    double* fro2 = state + 1;
    int* nrot = (int*)(state + n + 2);
    k_jacobi_init<<<n, 256, 0, st>>>(Vt, ldv, n, A, lda, fro2);
    B200_LAUNCH_CHECK();
    k_jacobi_fro<<<1, 256, 0, st>>>(fro2, n, state);
    B200_LAUNCH_CHECK();
    if (n > 1) {
        const int npl = n + (n & 1);
        double tol = sqrt((double)n) * 2.220446049250313e-16;
        if (tol < 1e-15) tol = 1e-15;
        int sweep = 0;
        for (; sweep < max_sweeps; sweep++) {
            B200_CUDA(cudaMemsetAsync(nrot, 0, sizeof(int), st));
            for (int r = 0; r < npl - 1; r++)
                k_jacobi_round<<<npl / 2, JT, 0, st>>>(A, lda, Vt, ldv, n, npl, r, tol, state, nrot);
            B200_LAUNCHED(npl - 1);
            B200_CUDA(cudaGetLastError());
            int h = 0;
            B200_CUDA(cudaMemcpyAsync(&h, nrot, sizeof(int), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            if (h == 0) {
                sweep++;
                break;
            }
        }
        if (sweeps_done) *sweeps_done = sweep;
    }
    k_jacobi_finish<<<n, JT, 0, st>>>(A, lda, Vt, ldv, n, lam);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace b200
