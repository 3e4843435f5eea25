// Symmetric eigendecomposition by tridiagonalisation (replaces np.linalg.eigh at lakernel.py:162, 201, 266 for the
// EigenKernel and the Cholesky repair branch).
//
// The matrices of this path have spectra graded over ten decades with ~1 % relative gaps everywhere, so the ABSOLUTE gaps
// between neighbours are far below the rounding level of the large eigenvalues: two-sided Jacobi (eigen.cu) then only
// converges linearly, ~37 sweeps of 12 n^3 flops.  This solver does what LAPACK-class solvers do, arranged for the GPU
// and batched over the OutStamps of a batch (one grid dimension = the system):
//
//   0. a matrix whose entries sit near the ends of the float64 range is scaled by a power of two first (dsyev's rule);
//   1. Householder tridiagonalisation  A = Q T Q^T, blocked in panels of 64 columns (dlatrd / dsytrd arrangement): the
//      trailing matrix is only read inside a panel and receives the panel's 128 rank-one updates as one DMMA GEMM
//      (k_trib_reflect / k_trib_symv / k_trib_p, three launches per column, 8 n^3 / 3 bytes of traffic per system);
//      B200_TRIDIAG=unblocked keeps the first version (k_tri_reflect + k_tri_update, one read-write pass per column);
//   2. eigenvalues of T by bisection on Sturm counts to the last bit, one thread per eigenvalue (k_tri_bisect);
//   3. eigenvectors of T by inverse iteration, one thread per eigenvalue, every eigenvalue independently
//      (k_tri_invit: LU with partial pivoting of T - lambda_j I, the computed eigenvalue itself as the shift, overflow-
//      safe back substitution): one solve from a pseudo-random start, orthonormalisation, one more solve from the
//      orthonormal vectors, orthonormalisation;
//   4. orthonormalisation of ALL vectors at once by Cholesky-QR on the n x n matrix of vectors: G = Z Z^T (DMMA GEMM),
//      G = L L^T and Z <- L^-1 Z through the batched Cholesky / triangular solve of linalg.cu.  A round whose Gram matrix
//      is already within 0.1 / n of the identity is skipped, a second final round runs only after an ill-conditioned
//      first.  Inside numerically degenerate clusters the pseudo-random starts make G a well-conditioned Gram matrix of
//      generic vectors of the cluster's subspace (any orthonormal basis of it serves the callers);
//   5. back-transformation Z <- Q Z with the reflectors in compact-WY panels of 128: three batched DMMA GEMMs per panel;
//   6. a system whose Gram matrix fails to factorise (linearly dependent vectors) is solved again, from a saved copy of
//      its matrix, by the block-Jacobi solver of eigen.cu.
//
// Checked on the device against NumPy (tests/test_gpu_parity.py::test_eigh_*, tests/test_gpu_fullsize.py at n = 1532 and
// n = 6248, tools/trieig_check.py, tools/eig_p4_debug.py): orthogonality and residual at the 1e-15 |A| level,
// eigenvalues to a few eps |A|.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

struct TriSys {
    double* A;     // (n, lda): in the matrix; out: reflector k in row k, columns k+1.. (v[k+1] = 1 stored)
    double* d;     // [n] diagonal of T
    double* e;     // [n] sub-diagonal of T (e[k] couples k and k+1; e[n-1] = 0)
    double* tau;   // [n]
    double* vbuf;  // [2][n] reflector of the current / previous column (zeros up to and including the column index)
    double* wbuf;  // [2][n]
    double* p;     // [n]
    int lda, n;
};
struct TriBatch {
    TriSys s[MAXB];
};

// Column k, single-CTA part: finish w_{k-1} = p - (tau/2)(p.v) v, bring row k up to date with the pending rank-2
// update of column k-1, then generate the reflector of column k (LAPACK dlarfg convention) from it.
// last = 1 (k = n-2): no reflector any more; rows n-2 and n-1 are completed and d, e written.
__global__ void __launch_bounds__(256) k_tri_reflect(TriBatch bt, int k) {
    __shared__ double red[40];
    const TriSys& s = bt.s[blockIdx.x];
    const int n = s.n, tid = threadIdx.x;
    if (k >= n || (n > 1 && k == n - 1)) return;  // (column n-1 is completed together with column n-2)
    const double* vp = s.vbuf + (size_t)((k + 1) & 1) * n;  // reflector of column k-1
    double* wp = s.wbuf + (size_t)((k + 1) & 1) * n;
    if (k > 0) {
        double dot = 0.0;
        for (int i = k + tid; i < n; i += 256) dot += s.p[i] * vp[i];
        dot = block_sum(dot, red);
        const double alpha = -0.5 * s.tau[k - 1] * dot;
        for (int i = k + tid; i < n; i += 256) wp[i] = s.p[i] + alpha * vp[i];
        __syncthreads();
    }
    double* row = s.A + (size_t)k * s.lda;
    if (k > 0) {
        const double vk = vp[k], wk = wp[k];
        for (int j = k + tid; j < n; j += 256) row[j] -= vk * wp[j] + wk * vp[j];
        __syncthreads();
    }
    if (k >= n - 2) {  // tail: no reflector for the last two columns
        if (k == n - 2) {
            double* row1 = s.A + (size_t)(n - 1) * s.lda;
            if (tid == 0) {
                if (k > 0) row1[n - 1] -= 2.0 * vp[n - 1] * wp[n - 1];
                s.d[n - 2] = row[n - 2];
                s.e[n - 2] = row[n - 1];
                s.d[n - 1] = row1[n - 1];
                s.e[n - 1] = 0.0;
                s.tau[n - 2] = 0.0;
                s.tau[n - 1] = 0.0;
            }
        } else if (tid == 0) {  // n == 1
            s.d[n - 1] = row[n - 1];
            s.e[n - 1] = 0.0;
            s.tau[n - 1] = 0.0;
        }
        return;
    }
    double* v = s.vbuf + (size_t)(k & 1) * n;
    double ss = 0.0;
    for (int j = k + 2 + tid; j < n; j += 256) ss += row[j] * row[j];
    ss = block_sum(ss, red);
    const double alpha0 = row[k + 1];
    double beta = alpha0, tau = 0.0, scale = 0.0;
    if (ss > 0.0) {
        const double nrm = sqrt(alpha0 * alpha0 + ss);
        beta = alpha0 >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha0) / beta;
        scale = 1.0 / (alpha0 - beta);
    }
    __syncthreads();
    for (int j = tid; j < n; j += 256) {
        double vj = 0.0;
        if (j == k + 1)
            vj = 1.0;
        else if (j > k + 1)
            vj = row[j] * scale;
        v[j] = vj;
        if (j > k) row[j] = vj;  // the reflector lives in the dead row k from now on
    }
    if (tid == 0) {
        s.d[k] = row[k];
        s.e[k] = beta;
        s.tau[k] = tau;
    }
}

// Column k, parallel part: every row i > k of the trailing matrix receives the pending rank-2 update of column k-1
// (columns > k) and contributes p_i = tau_k * (a_i . v_k).  One warp per row, 8 rows per CTA in flight.
__global__ void __launch_bounds__(256) k_tri_update(TriBatch bt, int k) {
    const TriSys& s = bt.s[blockIdx.y];
    const int n = s.n;
    if (k >= n - 2) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* vp = s.vbuf + (size_t)((k + 1) & 1) * n;
    const double* wp = s.wbuf + (size_t)((k + 1) & 1) * n;
    const double* v = s.vbuf + (size_t)(k & 1) * n;
    const double tau = s.tau[k];
    const bool pend = k > 0;
    for (int i = k + 1 + blockIdx.x * 8 + warp; i < n; i += gridDim.x * 8) {
        double* row = s.A + (size_t)i * s.lda;
        const double vi = pend ? vp[i] : 0.0, wi = pend ? wp[i] : 0.0;
        double acc = 0.0;
        for (int j = k + 1 + lane; j < n; j += 32) {
            double a = row[j];
            if (pend) {
                a -= vi * wp[j] + wi * vp[j];
                row[j] = a;
            }
            acc += a * v[j];
        }
        acc = warp_sum(acc);
        if (lane == 0) s.p[i] = tau * acc;
    }
}

// 1 / q to a couple of ulps for normal q of any magnitude: the hardware's 20-bit estimate (MUFU.RCP64H) and two Newton
// steps - a third of the instructions of an IEEE division, which is what the Sturm sweeps are bound by (|q| >= 1e-300 by
// construction, so neither the estimate nor the steps meet denormals or infinities).
__device__ __forceinline__ double fast_rcp(double q) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(q));
    r = fma(fma(-q, r, 1.0), r, r);
    r = fma(fma(-q, r, 1.0), r, r);
    return r;
}

// ---- eigenvalues of the tridiagonal matrix: bisection on Sturm counts, one thread per eigenvalue ----------------
__global__ void __launch_bounds__(256) k_tri_bisect(TriBatch bt, double* const* lam_out) {
    extern __shared__ double sm[];  // d[n], e2[n]
    const TriSys& s = bt.s[blockIdx.y];
    const int n = s.n;
    double* d = sm;
    double* e2 = sm + n;
    double lo = 1e300, hi = -1e300, emax = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double di = s.d[i], ei = i + 1 < n ? s.e[i] : 0.0, em = i > 0 ? s.e[i - 1] : 0.0;
        d[i] = di;
        e2[i] = ei * ei;
        const double r = fabs(ei) + fabs(em);
        lo = fmin(lo, di - r);
        hi = fmax(hi, di + r);
        emax = fmax(emax, fabs(ei));
    }
    __shared__ double slo[256], shi[256], sem[256];
    slo[threadIdx.x] = lo;
    shi[threadIdx.x] = hi;
    sem[threadIdx.x] = emax;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            slo[threadIdx.x] = fmin(slo[threadIdx.x], slo[threadIdx.x + o]);
            shi[threadIdx.x] = fmax(shi[threadIdx.x], shi[threadIdx.x + o]);
            sem[threadIdx.x] = fmax(sem[threadIdx.x], sem[threadIdx.x + o]);
        }
        __syncthreads();
    }
    const double gl = slo[0], gu = shi[0];
    const double tnorm = fmax(fabs(gl), fabs(gu));
    const double pivmin = fmax(2.2250738585072014e-308 * fmax(1.0, sem[0] * sem[0]), 1e-300);
    const double gl2 = gl - 2.0 * tnorm * 2.220446049250313e-16 * n - 2.0 * pivmin;
    const double gu2 = gu + 2.0 * tnorm * 2.220446049250313e-16 * n + 2.0 * pivmin;
    double* lam = lam_out[blockIdx.y];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        double a = gl2, b = gu2;  // eigenvalue j (ascending, 0-based): count(x) = #eigenvalues < x
        for (int it = 0; it < 200; it++) {
            const double x = 0.5 * (a + b);
            if (!(x > a) || !(x < b)) break;
            int cnt = 0;
            double q = d[0] - x;
            if (fabs(q) < pivmin) q = -pivmin;
            cnt += q < 0.0;
            for (int i = 1; i < n; i++) {
                q = d[i] - x - e2[i - 1] * fast_rcp(q);
                if (fabs(q) < pivmin) q = -pivmin;
                cnt += q < 0.0;
            }
            if (cnt > j)
                b = x;
            else
                a = x;
        }
        lam[j] = 0.5 * (a + b);
    }
}

// ---- eigenvectors of the tridiagonal matrix: inverse iteration, one thread per eigenvalue -----------------------
// (T - x I) = P L U with partial pivoting (U has two super-diagonals); three solves from a pseudo-random start.
// Per-thread arrays live in global scratch with the thread index fastest ([i][thread]: coalesced).
__device__ __forceinline__ double hash_unit(unsigned a, unsigned b) {
    unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0xC2B2AE3Du;
    h ^= h >> 13;
    h *= 0x27D4EB2Fu;
    h ^= h >> 16;
    return ((double)h + 0.5) * (2.0 / 4294967296.0) - 1.0;  // (-1, 1)
}

struct InvitSys {
    const double* d;
    const double* e;
    const double* lam;
    double* Zt;       // (ntot, ldz): row j <- eigenvector j of T
    double* scratch;  // 5 * n * nthr doubles: 1 / u0, u1, u2, multipliers, x  (element i of thread j at [i * nthr + j])
    unsigned char* piv;  // n * nthr
    int n, ldz, nthr;
};
struct InvitBatch {
    InvitSys s[MAXB];
};

// One step of inverse iteration per eigenvalue, one thread per eigenvalue, every eigenvalue independently.
// stage 0: factor T - shift_j I = P L U (partial pivoting; the factors stay in the scratch arrays), solve from a
//          pseudo-random start;  stage 1: solve again from row j of Zt (the orthonormalised vectors of stage 0).
//
// Shifts.  Bisection delivers every eigenvalue to the last bit, and the graded matrices of this path determine their
// small eigenvalues to high RELATIVE accuracy: neighbours whose gap is far below ulp(|T|) are still resolved, and a
// shift equal to the computed eigenvalue sits next to its own eigenvalue.  (Spreading the shifts of close eigenvalues
// by multiples of ulp(|T|), as a first version did, moves them next to OTHER eigenvalues: several threads then converge
// to the same vector and the Gram matrix of the orthogonalisation stage is singular - seen on the paper-4 stamp, whose
// 3000 eigenvalues inside +-1.5e-11 |T| have gaps of 0.3 - 3 ulp(|T|).)  Only eigenvalues that agree to a few ulps of
// THEMSELVES are pushed apart, by that amount (dstein's rule), so that no two threads factor the same matrix; such
// a pair is numerically degenerate and the two solves return generic vectors of its invariant subspace.
__global__ void __launch_bounds__(128) k_tri_invit(InvitBatch bt, int stage) {
    const InvitSys& s = bt.s[blockIdx.y];
    const int n = s.n, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const size_t st = (size_t)s.nthr;  // stride between consecutive i
    double* u0 = s.scratch + j;
    double* u1 = u0 + (size_t)n * st;
    double* u2 = u1 + (size_t)n * st;
    double* mm = u2 + (size_t)n * st;
    double* x = mm + (size_t)n * st;
    unsigned char* pv = s.piv + j;
    double* zrow = s.Zt + (size_t)j * s.ldz;
    const double eps = 2.220446049250313e-16;
    const double tn = fmax(fmax(fabs(s.lam[0]), fabs(s.lam[n - 1])), 1e-300);
    if ((stage & 1) == 0) {
        const double floor_ = 1e-290;  // (denormal eigenvalues: push by something representable)
        double xs = s.lam[j];
        {
            const double sep = 8.0 * eps * fmax(fabs(xs), floor_);
            int back = 0;
            while (j - back - 1 >= 0 && s.lam[j - back] - s.lam[j - back - 1] < sep && back < 4096) back++;
            if (back > 0) xs = s.lam[j - back] + back * sep;
        }
        if (stage & 2) xs = s.lam[n / 2];  // test hook (B200_EIGH_TEST_FAIL): every thread takes the same shift
        const double tiny = fmax(eps * tn, 1e-290);  // (normal: the reciprocals below flush denormals)
        // ---- factorisation (row i of the working pair is (a, b, c) = (diag, super1, super2)) ----
        double a = s.d[0] - xs, b = n > 1 ? s.e[0] : 0.0, c = 0.0;
        for (int i = 0; i < n - 1; i++) {
            const double sub = s.e[i];                                  // element (i+1, i)
            const double dn = s.d[i + 1] - xs, en = i + 2 < n ? s.e[i + 1] : 0.0;  // row i+1: (dn, en)
            double m;
            if (fabs(a) >= fabs(sub)) {  // no interchange
                if (fabs(a) < 1e-290) a = tiny;
                const double ra = fast_rcp(a);
                m = sub * ra;
                u0[i * st] = ra;  // (the pivots are stored as reciprocals: the solves multiply)
                u1[i * st] = b;
                u2[i * st] = c;
                pv[i * st] = 0;
                a = dn - m * b;
                b = en - m * c;
                c = 0.0;
            } else {  // rows i and i+1 swap
                const double rs_ = fast_rcp(sub);
                m = a * rs_;
                u0[i * st] = rs_;
                u1[i * st] = dn;
                u2[i * st] = en;
                pv[i * st] = 1;
                a = b - m * dn;
                b = c - m * en;
                c = 0.0;
            }
            mm[i * st] = m;
        }
        if (fabs(a) < 1e-290) a = tiny;
        u0[(size_t)(n - 1) * st] = fast_rcp(a);
        u1[(size_t)(n - 1) * st] = 0.0;
        u2[(size_t)(n - 1) * st] = 0.0;
        pv[(size_t)(n - 1) * st] = 0;
        for (int i = 0; i < n; i++) x[i * st] = hash_unit((unsigned)j, (unsigned)i);
    } else {
        for (int i = 0; i < n; i++) x[i * st] = zrow[i];
    }
    // ---- forward: apply P and L ----
    for (int i = 0; i < n - 1; i++) {
        const double m = mm[i * st];
        double xi = x[i * st], xn = x[(size_t)(i + 1) * st];
        if (pv[i * st] & 1) {
            const double t = xi;
            xi = xn;
            xn = t - m * xi;
        } else {
            xn -= m * xi;
        }
        x[i * st] = xi;
        x[(size_t)(i + 1) * st] = xn;
    }
    // ---- backward with U.  An eigenvector of a graded tridiagonal matrix is localised and decays by hundreds of decades
    // away from its centre, so the substitution can grow past the float64 range on its way in: whenever an entry
    // passes 1e150 everything (entries already computed and the rest of the right-hand side) is scaled by 1e-150.
    // The entries already stored are not revisited: each carries the number of rescalings that preceded it (bits
    // 1-7 of its pivot byte) and is brought to the final scale in the normalisation pass (dstein rescales likewise).
    double x1 = 0.0, x2 = 0.0, big = 0.0, rs = 1.0;
    int epoch = 0;
    for (int i = n - 1; i >= 0; i--) {
        double v = (x[i * st] * rs - u1[i * st] * x1 - u2[i * st] * x2) * u0[i * st];
        if (fabs(v) > 1e150 && epoch < 127) {
            v *= 1e-150;
            x1 *= 1e-150;
            big *= 1e-150;
            rs *= 1e-150;
            epoch++;
        }
        x[i * st] = v;
        pv[i * st] = (unsigned char)((pv[i * st] & 1) | (epoch << 1));
        x2 = x1;
        x1 = v;
        big = fmax(big, fabs(v));
    }
    // ---- normalise (scaled by the largest entry first, so that the squares neither overflow nor all underflow) ----
    const double sc = big > 0.0 ? 1.0 / big : 1.0;
    double nrm2 = 0.0;
    for (int i = 0; i < n; i++) {
        const int behind = epoch - (pv[i * st] >> 1);
        const double f = behind == 0 ? 1.0 : (behind == 1 ? 1e-150 : 0.0);
        const double v = x[i * st] * f * sc;
        x[i * st] = v;
        nrm2 += v * v;
    }
    const double inv = 1.0 / sqrt(fmax(nrm2, 1e-300));
    for (int i = 0; i < n; i++) zrow[i] = x[i * st] * inv;
}

// ---- compact-WY panels of the reflectors ------------------------------------------------------------------------
struct WySys {
    const double* A;    // reflector store (row k: v_k at columns > k)
    const double* tau;
    double *Vt, *V, *S, *T;  // (128, ntot), (ntot, 128), (128, 128), (128, 128)
    int lda, n, ntot, active;
};
struct WyBatch {
    WySys s[MAXB];
};

// Vt (128, ntot) <- rows p0 .. p0+127 of the reflector store (v_k[k+1] = 1 is stored), zero elsewhere; V its transpose.
__global__ void __launch_bounds__(256) k_wy_extract(WyBatch wb, int p0) {
    const WySys& w = wb.s[blockIdx.z];
    if (!w.active) return;
    const int r = blockIdx.y;  // reflector inside the panel
    const int k = p0 + r;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.ntot; i += gridDim.x * blockDim.x) {
        double v = 0.0;
        if (k < w.n - 2 && i > k && i < w.n) v = w.A[(size_t)k * w.lda + i];
        w.Vt[(size_t)r * w.ntot + i] = v;
        w.V[(size_t)i * NB + r] = v;
    }
}

// T (128 x 128, upper triangular, row-major) of H_p0 ... H_p0+127 = I - V T V^T from the Gram matrix S = V^T V
// (LAPACK dlarft, forward / columnwise): T[j][j] = tau_j, T[0:j, j] = -tau_j T[0:j, 0:j] S[0:j, j].  Row i of T only
// depends on row i of T and on S: four lanes of one warp own a row (quarters of the inner sum, two shuffles); the upper
// triangles of S and T are held packed in shared memory (2 x 66 KB).
constexpr int WY_PACKED = NB * (NB + 1) / 2;
__device__ __forceinline__ int wy_pk(int r, int c) { return r * NB - r * (r - 1) / 2 + (c - r); }  // r <= c
__global__ void __launch_bounds__(512) k_wy_tfactor(WyBatch wb, int p0) {
    extern __shared__ double wy_sm[];  // S packed, then T packed
    double* Ss = wy_sm;
    double* Ts = wy_sm + WY_PACKED;
    const WySys& w = wb.s[blockIdx.x];
    if (!w.active) return;
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * NB; e += 512) {
        const int r = e / NB, c = e - r * NB;
        if (r <= c) Ss[wy_pk(r, c)] = w.S[e];
    }
    __syncthreads();
    const int i = tid >> 2, part = tid & 3;  // row, quarter
    for (int j = (tid >> 5) * 8; j < NB; j++) {  // (warp-uniform trip count: the eight rows of a warp start at 8 w)
        const int k = p0 + j;
        const double tj = (k < w.n - 2) ? w.tau[k] : 0.0;
        double acc = 0.0;
        if (j > i)
            for (int l = i + part; l < j; l += 4) acc += Ts[wy_pk(i, l)] * Ss[wy_pk(l, j)];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (part == 0 && j >= i) Ts[wy_pk(i, j)] = (j == i) ? tj : -tj * acc;
        __syncwarp();
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += 512) {
        const int r = e / NB, c = e - r * NB;
        w.T[e] = r <= c ? Ts[wy_pk(r, c)] : 0.0;
    }
}

// ---- scaling (what dsyev does): a matrix whose entries sit near the ends of the float64 range is brought to max |a| ~ 1
// by a power of two before the reduction (squares of its entries would under- or overflow), and the eigenvalues are
// scaled back at the end.  All on the device: k_eigh_absmax, k_eigh_scale (a no-op pass unless max |a| is outside
// [1e-100, 1e100]), k_eigh_unscale.
struct ScaleBatch {
    double* A[MAXB];
    double* lam[MAXB];
    int lda[MAXB], n[MAXB];
};
__global__ void __launch_bounds__(256) k_eigh_absmax(ScaleBatch sb, double* amax) {
    __shared__ double red[8];
    const int q = blockIdx.y, n = sb.n[q];
    const double* A = sb.A[q];
    double m = 0.0;
    const size_t tot = (size_t)n * n;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / n), j = (int)(t - (size_t)i * n);
        const double v = fabs(A[(size_t)i * sb.lda[q] + j]);
        if (v <= 1.7e308) m = fmax(m, v);  // (non-finite entries do not steer the scaling)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) m = fmax(m, red[k]);
        atomicMax(reinterpret_cast<unsigned long long*>(amax + q), (unsigned long long)__double_as_longlong(m));
    }
}
__device__ __forceinline__ double eigh_scale_factor(double amax) {
    if (amax == 0.0 || (amax >= 1e-100 && amax <= 1e100)) return 1.0;
    int ex;
    frexp(amax, &ex);       // amax = f * 2^ex, 0.5 <= f < 1
    return ldexp(1.0, -ex);  // exact power of two: the scaled matrix has the same significands
}
__global__ void __launch_bounds__(256) k_eigh_scale(ScaleBatch sb, const double* amax) {
    const int q = blockIdx.y, n = sb.n[q];
    const double sc = eigh_scale_factor(amax[q]);
    if (sc == 1.0) return;
    double* A = sb.A[q];
    const size_t tot = (size_t)n * n;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / n), j = (int)(t - (size_t)i * n);
        A[(size_t)i * sb.lda[q] + j] *= sc;
    }
}
__global__ void __launch_bounds__(256) k_eigh_unscale(ScaleBatch sb, const double* amax) {
    const int q = blockIdx.y, n = sb.n[q];
    const double sc = eigh_scale_factor(amax[q]);
    if (sc == 1.0) return;
    const double inv = 1.0 / sc;  // (a power of two as well)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) sb.lam[q][i] *= inv;
}

// max |G - I| of each system's Gram matrix (non-finite entries count as +inf); dev[] zeroed by the caller
struct GramDev {
    const double* G[MAXB];
    int ntot[MAXB];
};
__global__ void __launch_bounds__(256) k_gram_dev(GramDev gd, double* dev) {
    __shared__ double red[40];
    const int q = blockIdx.y, ntot = gd.ntot[q];
    const double* G = gd.G[q];
    double m = 0.0;
    const size_t tot = (size_t)ntot * ntot;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / ntot), j = (int)(t - (size_t)i * ntot);
        double v = fabs(G[t] - (i == j ? 1.0 : 0.0));
        if (!(v <= 1.7e308)) v = INFINITY;
        m = fmax(m, v);
    }
    // block maximum, then one atomic per CTA (the bit patterns of non-negative doubles are ordered like the values)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[w] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) m = fmax(m, red[k]);
        atomicMax(reinterpret_cast<unsigned long long*>(dev + q), (unsigned long long)__double_as_longlong(m));
    }
}

// Zt rows n .. ntot-1 = unit vectors, columns n .. of the real rows = 0 (the identity padding of the caller)
__global__ void k_pad_vectors(double* __restrict__ Zt, int ldz, int n, int ntot) {
    const int r = blockIdx.y;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ntot; c += gridDim.x * blockDim.x) {
        if (r >= n)
            Zt[(size_t)r * ldz + c] = (r == c) ? 1.0 : 0.0;
        else if (c >= n)
            Zt[(size_t)r * ldz + c] = 0.0;
    }
}

// ---- blocked tridiagonalisation (LAPACK dlatrd / dsytrd arrangement, rows in the role of columns) ----------------
// Inside a panel of NBT columns the trailing matrix is NOT updated: column k is formed on the fly from the panel's
// reflectors V and companions W (a_k = A[k, :] - V W[k, :]^T - W V[k, :]^T), the symmetric matrix-vector product only
// READS the trailing matrix and is corrected with the same skinny products, and the 2 NBT rank-one updates of a panel
// reach the matrix as ONE DMMA GEMM  A -= [V | W] [W | V]^T  (K = 128).  Per column the matrix is read once instead of
// read and written once (8 n^3 / 3 instead of 16 n^3 / 3 bytes per system).  Three launches per column:
//   k_trib_reflect (one CTA per system): finish w_{k-1} = p + alpha v (alpha from the partial dot products), column
//                  a_k = c - 2 alpha v_{k-1}, reflector v_k (dlarfg), its copies (row k of A, VW / WV panels, vbuf);
//   k_trib_symv    (grid-wide): y = A[k+1:, k+1:] v, one warp per row, 16-byte loads; a few extra CTAs form partial
//                  sums of V^T v and W^T v;
//   k_trib_p       (grid-wide, one warp per row): p_i = tau (y_i - V[i, :] (W^T v) - W[i, :] (V^T v)), the partial dot
//                  products p . v, and the NEXT column with everything but the alpha terms: c_i = A[k+1, i] - (panel
//                  corrections of columns < j) - (v_i p_{k+1} + p_i)  (row i of V / W is in registers already).
constexpr int NBT = 64;      // panel width (the trailing update is a K = 2 NBT = 128 GEMM)
constexpr int TRIB_ZC = 16;  // row chunks of the V^T v / W^T v partial sums
constexpr int TRIB_GX = 1024;

struct TribSys {
    double* A;
    double *d, *e, *tau;
    double* VW;   // (ntot, 2 NBT): [V | W] of the current panel, row-major
    double* WV;   // (ntot, 2 NBT): [W | V]
    double *vbuf, *p, *c, *y, *abuf;  // [n]
    double* zpart;  // [TRIB_ZC][2 NBT]
    double* dpart;  // [TRIB_GX]
    int lda, n, ntot;
};
struct TribBatch {
    TribSys s[MAXB];
};

constexpr int RC = 8;  // elements of a column one thread of k_trib_reflect keeps in registers (columns up to 8192 long)
__global__ void __launch_bounds__(1024) k_trib_reflect(TribBatch bt, int k, int k0_prev, int k0, int finish, int reflect,
                                                      int fresh, int ndpart) {
    __shared__ double red[40];
    __shared__ double sh_a[2];
    const TribSys& s = bt.s[blockIdx.x];
    const int n = s.n, tid = threadIdx.x;
    const bool do_finish = finish && k >= 1 && k - 1 <= n - 3;
    const bool do_reflect = reflect && k <= n - 3;
    if (!do_finish && !do_reflect) return;
    double* row = s.A + (size_t)k * s.lda;
    if (s.ntot - k <= RC * 1024) {
        // every global load is issued before the first reduction: one round trip to memory, then registers only
        double pv[RC], cv[RC], vv[RC];
#pragma unroll
        for (int u = 0; u < RC; u++) {
            const int i = k + tid + u * 1024;
            pv[u] = cv[u] = vv[u] = 0.0;
            if (i < n) {
                if (!fresh || do_finish) vv[u] = s.vbuf[i];
                if (do_finish) pv[u] = s.p[i];
                if (do_reflect) cv[u] = fresh ? row[i] : s.c[i];
            }
        }
        double alpha = 0.0;
        if (do_finish) {
            double acc = 0.0;
            for (int i = tid; i < ndpart; i += 1024) acc += s.dpart[i];
            const double tprev = s.tau[k - 1];
            acc = block_sum(acc, red);
            alpha = -0.5 * tprev * acc;
            const int jp = k - 1 - k0_prev;
#pragma unroll
            for (int u = 0; u < RC; u++) {
                const int i = k + tid + u * 1024;
                if (i < s.ntot) {  // (rows above k are dead: nobody reads them again)
                    const double w = i < n ? pv[u] + alpha * vv[u] : 0.0;
                    s.VW[(size_t)i * (2 * NBT) + NBT + jp] = w;
                    s.WV[(size_t)i * (2 * NBT) + jp] = w;
                }
            }
        }
        if (!do_reflect) return;
        const int j = k - k0;
        double ss = 0.0;
#pragma unroll
        for (int u = 0; u < RC; u++) {
            const int i = k + tid + u * 1024;
            if (!fresh) cv[u] -= 2.0 * alpha * vv[u];  // cv = column k of the current matrix from here on
            if (i >= k + 2 && i < n) ss += cv[u] * cv[u];
        }
        if (tid < 2) sh_a[tid] = cv[0];  // a[k], a[k+1]
        ss = block_sum(ss, red);
        const double alpha0 = sh_a[1];
        double beta = alpha0, tau = 0.0, scale = 0.0;
        if (ss > 0.0) {
            const double nrm = sqrt(alpha0 * alpha0 + ss);
            beta = alpha0 >= 0.0 ? -nrm : nrm;
            tau = (beta - alpha0) / beta;
            scale = 1.0 / (alpha0 - beta);
        }
#pragma unroll
        for (int u = 0; u < RC; u++) {
            const int i = k + tid + u * 1024;
            if (i < s.ntot) {
                double vi = 0.0;
                if (i == k + 1)
                    vi = 1.0;
                else if (i > k + 1 && i < n)
                    vi = cv[u] * scale;
                s.VW[(size_t)i * (2 * NBT) + j] = vi;
                s.WV[(size_t)i * (2 * NBT) + NBT + j] = vi;
                if (i < n) {
                    s.vbuf[i] = vi;
                    if (i > k) row[i] = vi;  // the reflector lives in the dead row k from now on
                }
            }
        }
        if (tid == 0) {
            s.d[k] = sh_a[0];
            s.e[k] = beta;
            s.tau[k] = tau;
        }
        return;
    }
    // ---- general path (columns longer than RC * 1024): the same steps through the global work vector abuf ----
    double alpha = 0.0;
    if (do_finish) {
        const int jp = k - 1 - k0_prev;
        double acc = 0.0;
        for (int i = tid; i < ndpart; i += 1024) acc += s.dpart[i];
        acc = block_sum(acc, red);
        alpha = -0.5 * s.tau[k - 1] * acc;
        for (int i = k + tid; i < s.ntot; i += 1024) {
            const double w = i < n ? s.p[i] + alpha * s.vbuf[i] : 0.0;
            s.VW[(size_t)i * (2 * NBT) + NBT + jp] = w;
            s.WV[(size_t)i * (2 * NBT) + jp] = w;
        }
    }
    if (!do_reflect) return;
    const int j = k - k0;
    for (int i = k + tid; i < n; i += 1024) s.abuf[i] = fresh ? row[i] : s.c[i] - 2.0 * alpha * s.vbuf[i];
    __syncthreads();
    double ss = 0.0;
    for (int i = k + 2 + tid; i < n; i += 1024) ss += s.abuf[i] * s.abuf[i];
    ss = block_sum(ss, red);
    const double alpha0 = s.abuf[k + 1];
    double beta = alpha0, tau = 0.0, scale = 0.0;
    if (ss > 0.0) {
        const double nrm = sqrt(alpha0 * alpha0 + ss);
        beta = alpha0 >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha0) / beta;
        scale = 1.0 / (alpha0 - beta);
    }
    for (int i = k + tid; i < s.ntot; i += 1024) {
        double vi = 0.0;
        if (i == k + 1)
            vi = 1.0;
        else if (i > k + 1 && i < n)
            vi = s.abuf[i] * scale;
        s.VW[(size_t)i * (2 * NBT) + j] = vi;
        s.WV[(size_t)i * (2 * NBT) + NBT + j] = vi;
        if (i < n) {
            if (i >= k) s.vbuf[i] = vi;
            if (i > k) row[i] = vi;
        }
    }
    if (tid == 0) {
        s.d[k] = s.abuf[k];
        s.e[k] = beta;
        s.tau[k] = tau;
    }
}

__global__ void __launch_bounds__(256) k_trib_symv(TribBatch bt, int k, int gx_rows, int zc) {
    const TribSys& s = bt.s[blockIdx.y];
    const int n = s.n;
    if (k > n - 3) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* __restrict__ v = s.vbuf;
    if ((int)blockIdx.x < gx_rows) {
        int c0 = k + 1;
        const bool head = c0 & 1;
        if (head) c0++;
        for (int i = k + 1 + blockIdx.x * 8 + warp; i < n; i += gx_rows * 8) {
            const double* __restrict__ row = s.A + (size_t)i * s.lda;
            double acc = (head && lane == 0) ? row[k + 1] * v[k + 1] : 0.0;
            int c = c0 + 2 * lane;
#pragma unroll 4
            for (; c + 1 < n; c += 64) {
                const double2 a = *reinterpret_cast<const double2*>(row + c);
                const double2 b = *reinterpret_cast<const double2*>(v + c);
                acc += a.x * b.x;
                acc += a.y * b.y;
            }
            if (c < n) acc += row[c] * v[c];
            acc = warp_sum(acc);
            if (lane == 0) s.y[i] = acc;
        }
    } else {  // partial sums of [V | W]^T v over a chunk of rows: thread = (column t, row group h)
        __shared__ double part[2][128];
        const int chunk = blockIdx.x - gx_rows, t = threadIdx.x & 127, h = threadIdx.x >> 7;
        const int per = (n - k - 1 + zc - 1) / zc;
        const int i0 = k + 1 + chunk * per, i1 = min(n, i0 + per);
        double acc = 0.0;
#pragma unroll 4
        for (int i = i0 + h; i < i1; i += 2) acc += s.VW[(size_t)i * (2 * NBT) + t] * v[i];
        part[h][t] = acc;
        __syncthreads();
        if (h == 0) {
            double z = 0.0;
#pragma unroll
            for (int q = 0; q < 2; q++) z += part[q][t];
            s.zpart[chunk * (2 * NBT) + t] = z;
        }
    }
}

__global__ void __launch_bounds__(256) k_trib_p(TribBatch bt, int k, int k0, int gx, int zc) {
    __shared__ double zz[2 * NBT], rr[2 * NBT], wsum[8];
    const TribSys& s = bt.s[blockIdx.y];
    const int n = s.n;
    if (k > n - 3) return;
    const int j = k - k0, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double tau = s.tau[k];
    if (tid < 2 * NBT) {
        double z = 0.0, r = 0.0;
        if ((tid & (NBT - 1)) < j) {
            const int src = tid < NBT ? tid + NBT : tid - NBT;  // V columns pair with W^T v and vice versa
#pragma unroll
            for (int q = 0; q < TRIB_ZC; q++) z += s.zpart[q * (2 * NBT) + src];
            r = s.WV[(size_t)(k + 1) * (2 * NBT) + tid];
        }
        zz[tid] = z;
        rr[tid] = r;
    }
    __syncthreads();
    // p_{k+1} (every warp for itself: no block-wide phase for it)
    double pk1;
    {
        const double* vw = s.VW + (size_t)(k + 1) * (2 * NBT);
        const double yk1 = s.y[k + 1];
        double sp = 0.0;
        if (j > 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) sp += vw[lane + 32 * q] * zz[lane + 32 * q];
            sp = warp_sum(sp);
        }
        pk1 = tau * (yk1 - sp);
    }
    const double* rowk1 = s.A + (size_t)(k + 1) * s.lda;
    double dot = 0.0;
    for (int i = k + 1 + blockIdx.x * 8 + warp; i < n; i += gx * 8) {
        const double* vw = s.VW + (size_t)i * (2 * NBT);
        const double yi = s.y[i], vi = s.vbuf[i], ai = rowk1[i];  // (same address in every lane: one broadcast load each)
        double sp = 0.0, sc = 0.0;
        if (j > 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double x = vw[lane + 32 * q];
                sp += x * zz[lane + 32 * q];
                sc += x * rr[lane + 32 * q];
            }
            sp = warp_sum(sp);
            sc = warp_sum(sc);
        }
        if (lane == 0) {
            const double pi = tau * (yi - sp);
            s.p[i] = pi;
            s.c[i] = ai - sc - (vi * pk1 + pi);
            dot += pi * vi;
        }
    }
    if (lane == 0) wsum[warp] = dot;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int q = 0; q < 8; q++) t += wsum[q];
        s.dpart[blockIdx.x] = t;
    }
}

// last two rows: d[n-2], e[n-2], d[n-1] from the matrix and the reflectors of the last (unfinished) panel
__global__ void __launch_bounds__(32) k_trib_tail(TribBatch bt) {
    const TribSys& s = bt.s[blockIdx.x];
    const int n = s.n, lane = threadIdx.x;
    if (n == 1) {
        if (lane == 0) {
            s.d[0] = s.A[0];
            s.e[0] = 0.0;
            s.tau[0] = 0.0;
        }
        return;
    }
    int J = 0;
    if (n >= 3) {
        const int kl = n - 3;
        J = kl - (kl / NBT) * NBT + 1;
    }
    const double* r0 = s.VW + (size_t)(n - 2) * (2 * NBT);
    const double* r1 = s.VW + (size_t)(n - 1) * (2 * NBT);
    double c00 = 0.0, c01 = 0.0, c11 = 0.0;
    for (int l = lane; l < J; l += 32) {
        const double v0 = r0[l], w0 = r0[NBT + l], v1 = r1[l], w1 = r1[NBT + l];
        c00 += 2.0 * v0 * w0;
        c01 += v0 * w1 + w0 * v1;
        c11 += 2.0 * v1 * w1;
    }
    c00 = warp_sum(c00);
    c01 = warp_sum(c01);
    c11 = warp_sum(c11);
    if (lane == 0) {
        const double* a0 = s.A + (size_t)(n - 2) * s.lda;
        const double* a1 = s.A + (size_t)(n - 1) * s.lda;
        s.d[n - 2] = a0[n - 2] - c00;
        s.e[n - 2] = a0[n - 1] - c01;
        s.d[n - 1] = a1[n - 1] - c11;
        s.e[n - 1] = 0.0;
        s.tau[n - 2] = 0.0;
        s.tau[n - 1] = 0.0;
    }
}

}  // namespace

// Tridiagonalisation of every problem (debug / test entry as well): d, e, tau are device arrays of n doubles.
static int tridiagonalise_unblocked(const TriBatch& bt, int nsys, int nmax, cudaStream_t st) {
    for (int k = 0; k < nmax; k++) {
        k_tri_reflect<<<nsys, 256, 0, st>>>(bt, k);
        if (k < nmax - 2) {
            const int rows = nmax - k - 1;
            int gx = (rows + 7) / 8;
            if (gx > 4 * 148) gx = 4 * 148;
            k_tri_update<<<dim3(gx, nsys), 256, 0, st>>>(bt, k);
        }
    }
    B200_LAUNCHED(2 * nmax);
    B200_CUDA(cudaGetLastError());
    return 0;
}

static int tridiagonalise(const TriBatch& bt, int nsys, int nmax, cudaStream_t st) {
    static const bool unblocked = [] {
        const char* e = getenv("B200_TRIDIAG");
        return e && strcmp(e, "unblocked") == 0;
    }();
    if (unblocked) return tridiagonalise_unblocked(bt, nsys, nmax, st);
    int ntot_max = 0;
    for (int q = 0; q < nsys; q++) {
        const int ntot = (bt.s[q].n + NB - 1) / NB * NB;
        ntot_max = ntot > ntot_max ? ntot : ntot_max;
    }
    void* ws = nullptr;
    const size_t nv = ((size_t)nmax + 1) & ~(size_t)1;  // (even: the vectors are read with 16-byte loads)
    const size_t per = 2 * (size_t)ntot_max * (2 * NBT) + 5 * nv + TRIB_ZC * 2 * NBT + TRIB_GX + 64;
    if (int rc = scratch(13, sizeof(double) * per * nsys, &ws)) return rc;
    // stale panel entries are multiplied by zero masks: they must be finite
    B200_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * per * nsys, st));
    TribBatch tb;
    for (int q = 0; q < nsys; q++) {
        TribSys& t = tb.s[q];
        double* b = static_cast<double*>(ws) + per * q;
        t.A = bt.s[q].A;
        t.d = bt.s[q].d;
        t.e = bt.s[q].e;
        t.tau = bt.s[q].tau;
        t.lda = bt.s[q].lda;
        t.n = bt.s[q].n;
        t.ntot = (t.n + NB - 1) / NB * NB;
        t.VW = b;
        t.WV = t.VW + (size_t)ntot_max * (2 * NBT);
        t.vbuf = t.WV + (size_t)ntot_max * (2 * NBT);
        t.p = t.vbuf + nv;
        t.c = t.p + nv;
        t.y = t.c + nv;
        t.abuf = t.y + nv;
        t.zpart = t.abuf + nv;
        t.dpart = t.zpart + TRIB_ZC * 2 * NBT;
    }
    for (int q = nsys; q < MAXB; q++) tb.s[q] = tb.s[0];
    int launches = 0, gx_prev = 0;
    const int klast = nmax - 3;  // last column with a reflector
    for (int k = 0; k <= klast; k++) {
        const int k0 = (k / NBT) * NBT, fresh = k == k0;
        k_trib_reflect<<<nsys, 1024, 0, st>>>(tb, k, k0, k0, !fresh, 1, fresh, gx_prev);
        const int rows = nmax - k - 1;
        int gs = (rows + 7) / 8;  // symv: one warp per row, 8 rows per CTA
        if (gs > TRIB_GX) gs = TRIB_GX;
        // p: a few CTAs per SM over all systems (each pays a fixed preamble), several rows per warp
        int gx = (rows + 7) / 8;
        const int cap = (4 * 148 + nsys - 1) / nsys;
        if (gx > cap) gx = cap;
        const int zc = fresh ? 0 : TRIB_ZC;
        k_trib_symv<<<dim3(gs + zc, nsys), 256, 0, st>>>(tb, k, gs, zc);
        k_trib_p<<<dim3(gx, nsys), 256, 0, st>>>(tb, k, k0, gx, zc);
        launches += 3;
        gx_prev = gx;
        if (k - k0 == NBT - 1 || k == klast) {  // end of the panel: finish its last w, then the trailing update
            k_trib_reflect<<<nsys, 1024, 0, st>>>(tb, k + 1, k0, k0, 1, 0, 0, gx_prev);
            launches++;
            if (k < klast) {
                const int k1 = k0 + NBT;
                GemmProb g[MAXB];
                int ng = 0;
                for (int q = 0; q < nsys; q++) {
                    const TribSys& t = tb.s[q];
                    if (t.n - 3 < k1) continue;  // no further column: the tail kernel reads the panel directly
                    g[ng++] = GemmProb{t.VW + (size_t)k1 * (2 * NBT), t.WV + (size_t)k1 * (2 * NBT),
                                       t.A + (size_t)k1 * t.lda + k1, 2 * NBT, 2 * NBT, t.lda, t.ntot - k1, t.ntot - k1,
                                       2 * NBT};
                }
                if (int rc = launch_gemm_nt_batch(g, ng, -1, st)) return rc;
            }
        }
    }
    k_trib_tail<<<nsys, 32, 0, st>>>(tb);
    B200_LAUNCHED(launches + 1);
    B200_CUDA(cudaGetLastError());
    return 0;
}

// B200_EIGH_TIMING=1: wall time of each stage (synchronises the stream; experiments only)
struct StageTimer {
    cudaStream_t st;
    bool on;
    cudaEvent_t a, b;
    explicit StageTimer(cudaStream_t s) : st(s), on(getenv("B200_EIGH_TIMING") != nullptr) {
        if (on) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
        }
    }
    void lap(const char* what) {
        if (!on) return;
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        fprintf(stderr, "  eigh stage %-22s %8.2f ms\n", what, ms);
        cudaEventRecord(a, st);
    }
    ~StageTimer() {
        if (on) {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    }
};

constexpr int QR_ROUNDS = 3;
static long long g_eigh_fallbacks = 0;

int launch_tri_eigh_batch(const EighProblem* pr, int nsys, cudaStream_t st) {
    if (nsys <= 0) return 0;
    B200_REQUIRE(nsys <= MAXB, "at most MAXB eigenproblems per batched call");
    int nmax = 0, ntot_max = 0;
    for (int q = 0; q < nsys; q++) {
        const int ntot = (pr[q].n + NB - 1) / NB * NB;
        B200_REQUIRE(pr[q].n > 0 && pr[q].lda >= ntot && pr[q].ldv >= ntot && pr[q].lda % 2 == 0 && pr[q].ldv % 2 == 0,
                     "eigh: A and Vt must be padded to a multiple of 128 rows / columns, even leading dimensions");
        nmax = pr[q].n > nmax ? pr[q].n : nmax;
        ntot_max = ntot > ntot_max ? ntot : ntot_max;
    }
    prof_begin(PROF_EIGH, st);
    StageTimer tm(st);
    int* qr_info = nullptr;
    // scaling of badly scaled matrices (no-op passes for everything this path produces: max |a| = 1)
    void* sc_ws = nullptr;
    if (int rc = scratch(15, sizeof(double) * MAXB, &sc_ws)) return rc;
    double* d_amax = static_cast<double*>(sc_ws);
    ScaleBatch sbt;
    for (int q = 0; q < MAXB; q++) {
        const int qq = q < nsys ? q : 0;
        sbt.A[q] = pr[qq].A;
        sbt.lam[q] = pr[qq].lam;
        sbt.lda[q] = pr[qq].lda;
        sbt.n[q] = pr[qq].n;
    }
    B200_CUDA(cudaMemsetAsync(d_amax, 0, sizeof(double) * MAXB, st));
    k_eigh_absmax<<<dim3(64, nsys), 256, 0, st>>>(sbt, d_amax);
    k_eigh_scale<<<dim3(64, nsys), 256, 0, st>>>(sbt, d_amax);
    B200_LAUNCHED(2);
    B200_CUDA(cudaGetLastError());
    // copies of the matrices (the tridiagonalisation destroys them): only read again if the orthogonalisation fails
    void* acopy = nullptr;
    if (int rc = scratch(14, sizeof(double) * (size_t)ntot_max * ntot_max * nsys, &acopy)) return rc;
    for (int q = 0; q < nsys; q++) {
        const int ntot = (pr[q].n + NB - 1) / NB * NB;
        B200_CUDA(cudaMemcpy2DAsync(static_cast<double*>(acopy) + (size_t)ntot_max * ntot_max * q, sizeof(double) * ntot,
                                    pr[q].A, sizeof(double) * pr[q].lda, sizeof(double) * ntot, ntot,
                                    cudaMemcpyDeviceToDevice, st));
    }
    const int test_fail = getenv("B200_EIGH_TEST_FAIL") != nullptr ? 2 : 0;
    // ---- scratch: vectors of the tridiagonalisation, inverse-iteration arrays, WY panels, Gram matrices ----
    void* ws = nullptr;
    const size_t per_vec = 8 * (size_t)nmax;
    if (int rc = scratch(9, sizeof(double) * per_vec * nsys + sizeof(double*) * MAXB + 256, &ws)) return rc;
    double* vec = static_cast<double*>(ws);
    double** lam_ptrs = reinterpret_cast<double**>(vec + per_vec * nsys);
    TriBatch bt;
    double* h_lam[MAXB];
    for (int q = 0; q < nsys; q++) {
        TriSys& s = bt.s[q];
        double* b = vec + per_vec * q;
        s.A = pr[q].A;
        s.lda = pr[q].lda;
        s.n = pr[q].n;
        s.d = b;
        s.e = b + nmax;
        s.tau = b + 2 * (size_t)nmax;
        s.vbuf = b + 3 * (size_t)nmax;
        s.wbuf = b + 5 * (size_t)nmax;
        s.p = b + 7 * (size_t)nmax;
        h_lam[q] = pr[q].lam;
    }
    for (int q = nsys; q < MAXB; q++) {
        bt.s[q] = bt.s[0];
        h_lam[q] = h_lam[0];
    }
    B200_CUDA(cudaMemcpyAsync(lam_ptrs, h_lam, sizeof(double*) * MAXB, cudaMemcpyHostToDevice, st));
    if (int rc = tridiagonalise(bt, nsys, nmax, st)) return rc;
    tm.lap("tridiagonalisation");
    // ---- eigenvalues ----
    {
        const size_t smem = sizeof(double) * 2 * (size_t)nmax;
        B200_REQUIRE(smem <= 200 * 1024, "eigh: matrix too large for the bisection kernel's shared memory");
        static bool attr = false;
        if (!attr) {
            B200_CUDA(cudaFuncSetAttribute(k_tri_bisect, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr = true;
        }
        k_tri_bisect<<<dim3((nmax + 255) / 256, nsys), 256, smem, st>>>(bt, lam_ptrs);
        B200_LAUNCH_CHECK();
    }
    tm.lap("bisection");
    // ---- eigenvectors of T (rows of Vt): solve, orthonormalise, solve again from the orthonormal vectors, orthonormalise
    // twice.  Orthonormalisation = Cholesky-QR on the rows of Vt:  G = Z Z^T = L L^T,  Z <- L^-1 Z ----
    InvitBatch ib;
    const int nthr = (nmax + 127) / 128 * 128;
    {
        void* is = nullptr;
        const size_t per = 5 * (size_t)nmax * nthr;
        if (int rc = scratch(10, sizeof(double) * per * nsys + (size_t)nmax * nthr * nsys + 256, &is)) return rc;
        double* base = static_cast<double*>(is);
        unsigned char* pbase = reinterpret_cast<unsigned char*>(base + per * nsys);
        for (int q = 0; q < nsys; q++) {
            ib.s[q] = InvitSys{bt.s[q].d, bt.s[q].e, pr[q].lam, pr[q].Vt, base + per * q,
                               pbase + (size_t)nmax * nthr * q, pr[q].n, pr[q].ldv, nthr};
        }
        for (int q = nsys; q < MAXB; q++) ib.s[q] = ib.s[0];
    }
    void* gs = nullptr;
    const size_t per_g = 2 * (size_t)ntot_max * ntot_max + 2 * (size_t)(ntot_max / NB) * NB * NB;
    if (int rc = scratch(11, sizeof(double) * per_g * nsys + sizeof(int) * QR_ROUNDS * MAXB + sizeof(double) * MAXB + 256, &gs))
        return rc;
    double* gb = static_cast<double*>(gs);
    qr_info = reinterpret_cast<int*>(gb + per_g * nsys);
    B200_CUDA(cudaMemsetAsync(qr_info, 0, sizeof(int) * QR_ROUNDS * MAXB, st));
    double* d_dev = reinterpret_cast<double*>(qr_info + QR_ROUNDS * MAXB);
    static const bool always_qr = getenv("B200_EIGH_ALWAYS_QR") != nullptr;  // (experiments: never skip a round)
    // One round.  The Gram matrix is always formed and its distance from the identity read back (one stream
    // synchronisation); when that is below skip_below the vectors are orthonormal enough for what follows and the
    // factorisation + triangular solve are skipped.  *dev_out receives the distance.
    auto cholqr = [&](int round, double skip_below, double* dev_out) -> int {
        SolveSys sys[MAXB];
        GemmProb gram[MAXB];
        GramDev gd;
        int* info = qr_info + round * MAXB;
        for (int q = 0; q < nsys; q++) {
            const int ntot = (pr[q].n + NB - 1) / NB * NB;
            double* G = gb + per_g * q;
            gram[q] = GemmProb{pr[q].Vt, pr[q].Vt, G, pr[q].ldv, pr[q].ldv, ntot, ntot, ntot, ntot};
            gd.G[q] = G;
            gd.ntot[q] = ntot;
        }
        for (int q = nsys; q < MAXB; q++) {
            gd.G[q] = gd.G[0];
            gd.ntot[q] = 0;
        }
        if (int rc = launch_gemm_nt_batch(gram, nsys, 0, st)) return rc;
        B200_CUDA(cudaMemsetAsync(d_dev, 0, sizeof(double) * MAXB, st));
        k_gram_dev<<<dim3(64, nsys), 256, 0, st>>>(gd, d_dev);
        B200_LAUNCH_CHECK();
        double h_dev[MAXB];
        B200_CUDA(cudaMemcpyAsync(h_dev, d_dev, sizeof(double) * MAXB, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        double dev = 0.0;
        for (int q = 0; q < nsys; q++) dev = h_dev[q] > dev ? h_dev[q] : dev;
        if (dev_out) *dev_out = dev;
        if (tm.on) fprintf(stderr, "  eigh round %d: max |G - I| = %.2e\n", round, dev);
        if (dev < skip_below && !always_qr) return 0;
        for (int q = 0; q < nsys; q++) {
            const int ntot = (pr[q].n + NB - 1) / NB * NB;
            double* G = gb + per_g * q;
            double* Xt = G + (size_t)ntot_max * ntot_max;  // Z^T (components x vectors)
            double* Dinv = Xt + (size_t)ntot_max * ntot_max;
            if (int rc = launch_transpose(pr[q].Vt, pr[q].ldv, Xt, ntot, ntot, ntot, st)) return rc;
            SolveSys& s = sys[q];
            s.W = G;
            s.X = Xt;
            s.Dinv = Dinv;
            s.info = info + q;
            s.npad = ntot;
            s.mpad = ntot;
            s.ldw = ntot;
            s.ldx = ntot;
            s.mrows = 0;
            s.pad_ = 0;
            s.work = nullptr;
            s.work_bytes = 0;
        }
        if (int rc = launch_chol_solve(sys, nsys, 1, 2, st)) return rc;  // factor + FORWARD solve only: X <- X L^-T
        for (int q = 0; q < nsys; q++) {
            const int ntot = (pr[q].n + NB - 1) / NB * NB;
            double* Xt = gb + per_g * q + (size_t)ntot_max * ntot_max;
            if (int rc = launch_transpose(Xt, ntot, pr[q].Vt, pr[q].ldv, ntot, ntot, st)) return rc;
        }
        return 0;
    };
    k_tri_invit<<<dim3(nthr / 128, nsys), 128, 0, st>>>(ib, 0 | test_fail);
    B200_LAUNCH_CHECK();
    for (int q = 0; q < nsys; q++) {
        const int ntot = (pr[q].n + NB - 1) / NB * NB;
        k_pad_vectors<<<dim3((ntot + 255) / 256, ntot), 256, 0, st>>>(pr[q].Vt, pr[q].ldv, pr[q].n, ntot);
        B200_LAUNCHED(1);
    }
    B200_CUDA(cudaGetLastError());
    tm.lap("inverse iteration 1");
    // max |G - I| * ntot bounds the 2-norm of G - I: below 0.1 the Gram matrix has a condition number < 1.25.
    // Between the solves the vectors only have to be far from collapsing: such nearly orthonormal ones go on as they are
    const double well = 0.1 / ntot_max;
    if (int rc = cholqr(0, well, nullptr)) return rc;
    tm.lap("Cholesky-QR");
    k_tri_invit<<<dim3(nthr / 128, nsys), 128, 0, st>>>(ib, 1);
    B200_LAUNCH_CHECK();
    tm.lap("inverse iteration 2");
    // one round leaves an orthogonality error of ~ eps * cond(G): a second one only after an ill-conditioned first
    double dev1 = 0.0;
    if (int rc = cholqr(1, 0.0, &dev1)) return rc;
    if (dev1 >= well || always_qr)
        if (int rc = cholqr(2, 0.0, nullptr)) return rc;
    tm.lap("Cholesky-QR x 1-2");
    // ---- back-transformation: rows of Vt <- eigenvectors of A;  Z <- (I - V T V^T) Z panel by panel, last panel first,
    // every launch over all systems ----
    {
        static bool attr2 = false;
        if (!attr2) {
            B200_CUDA(cudaFuncSetAttribute(k_wy_tfactor, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(sizeof(double) * NB * (NB + 1))));
            attr2 = true;
        }
        void* wsy = nullptr;
        const size_t per = 2 * (size_t)NB * ntot_max + 2 * (size_t)NB * NB + 2 * (size_t)ntot_max * NB;
        if (int rc = scratch(12, sizeof(double) * per * nsys, &wsy)) return rc;
        WyBatch wb;
        double *Wt[MAXB], *Yt[MAXB];
        int npanel_max = 0;
        for (int q = 0; q < nsys; q++) {
            const int n = pr[q].n, ntot = (n + NB - 1) / NB * NB;
            double* b = static_cast<double*>(wsy) + per * q;
            WySys& w = wb.s[q];
            w.A = bt.s[q].A;
            w.tau = bt.s[q].tau;
            w.lda = bt.s[q].lda;
            w.n = n;
            w.ntot = ntot;
            w.Vt = b;                              // (128, ntot)
            w.V = w.Vt + (size_t)NB * ntot_max;    // (ntot, 128)
            w.S = w.V + (size_t)ntot_max * NB;     // (128, 128)
            w.T = w.S + (size_t)NB * NB;           // (128, 128)
            Wt[q] = w.T + (size_t)NB * NB;         // (ntot, 128)
            Yt[q] = Wt[q] + (size_t)ntot_max * NB; // (ntot, 128)
            w.active = 0;
            const int np = n >= 3 ? (n - 2 + NB - 1) / NB : 0;
            npanel_max = np > npanel_max ? np : npanel_max;
        }
        for (int q = nsys; q < MAXB; q++) {
            wb.s[q] = wb.s[0];
            wb.s[q].active = 0;
        }
        for (int pnl = npanel_max - 1; pnl >= 0; pnl--) {
            const int p0 = pnl * NB;
            GemmProb gS[MAXB], gW[MAXB], gY[MAXB], gZ[MAXB];
            int na = 0;
            for (int q = 0; q < nsys; q++) {
                WySys& w = wb.s[q];
                w.active = (w.n >= 3 && p0 < w.n - 2) ? 1 : 0;
                if (!w.active) continue;
                const int ntot = w.ntot;
                gS[na] = GemmProb{w.Vt, w.Vt, w.S, ntot, ntot, NB, NB, NB, ntot};          // S = V^T V
                gW[na] = GemmProb{pr[q].Vt, w.Vt, Wt[q], pr[q].ldv, ntot, NB, ntot, NB, ntot};  // Wt = Zt V
                gY[na] = GemmProb{Wt[q], w.T, Yt[q], NB, NB, NB, ntot, NB, NB};            // Yt = Wt T^T
                gZ[na] = GemmProb{Yt[q], w.V, pr[q].Vt, NB, NB, pr[q].ldv, ntot, ntot, NB};  // Zt -= Yt V^T
                na++;
            }
            if (na == 0) continue;
            k_wy_extract<<<dim3((ntot_max + 255) / 256, NB, nsys), 256, 0, st>>>(wb, p0);
            B200_LAUNCHED(1);
            if (int rc = launch_gemm_nt_batch(gS, na, 0, st)) return rc;
            k_wy_tfactor<<<nsys, 512, sizeof(double) * 2 * WY_PACKED, st>>>(wb, p0);
            B200_LAUNCHED(1);
            if (int rc = launch_gemm_nt_batch(gW, na, 0, st)) return rc;
            if (int rc = launch_gemm_nt_batch(gY, na, 0, st)) return rc;
            if (int rc = launch_gemm_nt_batch(gZ, na, -1, st)) return rc;
        }
        B200_CUDA(cudaGetLastError());
    }
    tm.lap("back-transformation");
    prof_end(0.0, st);
    // The Gram matrices of all Cholesky-QR rounds must have been positive definite: a failed factorisation means the
    // inverse-iteration vectors were linearly dependent, and everything after it is garbage.  Such a system is solved
    // again from the saved copy of its matrix with the block-Jacobi solver of eigen.cu (slow, but it has no such failure
    // mode); b200_eigh_fallback_count() reports how often that happened.
    int h_info[QR_ROUNDS * MAXB];
    B200_CUDA(cudaMemcpyAsync(h_info, qr_info, sizeof(h_info), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    EighProblem redo[MAXB];
    int nredo = 0;
    for (int q = 0; q < nsys; q++) {
        bool bad = false;
        for (int r = 0; r < QR_ROUNDS; r++) bad = bad || h_info[r * MAXB + q] != 0;
        if (!bad) continue;
        const int ntot = (pr[q].n + NB - 1) / NB * NB;
        B200_CUDA(cudaMemcpy2DAsync(pr[q].A, sizeof(double) * pr[q].lda,
                                    static_cast<double*>(acopy) + (size_t)ntot_max * ntot_max * q, sizeof(double) * ntot,
                                    sizeof(double) * ntot, ntot, cudaMemcpyDeviceToDevice, st));
        redo[nredo++] = pr[q];
    }
    if (nredo > 0) {
        g_eigh_fallbacks += nredo;
        int sweeps = 0;
        if (int rc = launch_jacobi_eigh_batch(redo, nredo, 80, &sweeps, st)) return rc;
        for (int q = 0; q < nredo; q++) {  // (the Jacobi solver owns a smaller padding: restore ours)
            const int ntot = (redo[q].n + NB - 1) / NB * NB;
            k_pad_vectors<<<dim3((ntot + 255) / 256, ntot), 256, 0, st>>>(redo[q].Vt, redo[q].ldv, redo[q].n, ntot);
            B200_LAUNCHED(1);
        }
        B200_CUDA(cudaGetLastError());
    }
    k_eigh_unscale<<<dim3(8, nsys), 256, 0, st>>>(sbt, d_amax);
    B200_LAUNCH_CHECK();
    return 0;
}

long long eigh_fallback_count() { return g_eigh_fallbacks; }

// Householder tridiagonalisation alone (tests): A (n x n, lda) is overwritten by the reflectors, d / e / tau receive
// T's diagonal, sub-diagonal and the reflector scales (device arrays of n doubles).
int launch_tridiag(double* A, int lda, int n, double* d, double* e, double* tau, cudaStream_t st) {
    if (n <= 0) return 0;
    void* ws = nullptr;
    if (int rc = scratch(9, sizeof(double) * 8 * (size_t)n + 1024, &ws)) return rc;
    double* b = static_cast<double*>(ws);
    TriBatch bt;
    TriSys& s = bt.s[0];
    s.A = A;
    s.lda = lda;
    s.n = n;
    s.d = d;
    s.e = e;
    s.tau = tau;
    s.vbuf = b;
    s.wbuf = b + 2 * (size_t)n;
    s.p = b + 4 * (size_t)n;
    for (int q = 1; q < MAXB; q++) bt.s[q] = bt.s[0];
    return tridiagonalise(bt, 1, n, st);
}

}  // namespace b200
