// Symmetric eigendecomposition for EigenKernel and the Cholesky repair branch:
// replaces np.linalg.eigh at lakernel.py:162, 201, 266.
//
// Two-sided (classical) Jacobi in implicit, blocked form: the kernels keep G = V^T A and V^T (V = I initially) and
// never store H = V^T A V; the entries a rotation needs are dot products of rows, h_pq = g_p . v_q, and a rotation
// of the pair (p, q) is applied to rows p, q of both G and V^T (which is H <- J^T H J).  Working on H rather than on
// the Gram matrix G G^T = V^T A^2 V (one-sided Hestenes) keeps the absolute accuracy at eps |A|_F for the small
// eigenvalues too, which is what LAPACK's eigh delivers and what 1/(lam + kappa) in EigenKernel needs.
// Rotations are skipped when |h_pq| <= 2 sqrt(n) eps |A|_F (the rounding floor of the implicit dot products): any
// orthonormal basis of a numerically degenerate subspace serves the callers (T, Sigma and U/C are invariant to it).
// Every reduction uses a fixed thread -> element mapping and a fixed tree: the result is deterministic.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int JT = 256;
constexpr int EIGH_INNER_DEFAULT = 1;

// V^T = I on the whole padded range (ntot rows/columns); per-row squared norms of the real n x n part of A
__global__ void k_jacobi_init(double* __restrict__ Vt, int ldv, int n, int ntot, const double* __restrict__ G, int lda,
                              double* __restrict__ fro2) {
    __shared__ double red[40];
    const int r = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < ntot; c += blockDim.x) {
        Vt[(size_t)r * ldv + c] = (r == c) ? 1.0 : 0.0;
        if (r < n && c < n) {
            const double v = G[(size_t)r * lda + c];
            s += v * v;
        }
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0 && r < n) fro2[r] = s;
}

// sum of the per-row squared norms -> state[0] = |A|_F^2 (single CTA, deterministic)
__global__ void k_jacobi_fro(const double* __restrict__ fro2, int n, double* __restrict__ state) {
    __shared__ double red[40];
    double s = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) s += fro2[c];
    s = block_sum(s, red);
    if (threadIdx.x == 0) state[0] = s;
}

// ---- block Jacobi round ---------------------------------------------------------------------------------------
// The row-pair kernel above moves 8 rows of length n through L2 for ONE rotation.  Here a CTA takes a PAIR OF
// 16-ROW BLOCKS (P, Q): it forms the 32x32 matrix H_sub = G_PQ V_PQ^T (= V_PQ^T A V_PQ) with one pass over the 64
// rows, runs one cyclic sweep of two-sided Jacobi on it in shared memory (496 rotations, accumulated in J), and
// applies J^T to the 32 rows of G and of V^T with a second pass: 496 row pairs per visit for the traffic of 16.
// Block pairs follow the same round-robin tournament (nblk/2 disjoint pairs per round = one launch, nblk - 1 rounds
// per sweep).  All reductions have a fixed order: the result is deterministic.
constexpr int BJ = 16;           // rows per block
constexpr int B2 = 2 * BJ;       // rows per block pair
constexpr int CK1 = 64;          // phase 1: columns per chunk (G and V chunks, double buffered)
constexpr int LDH = CK1 + 4;     // 68 = 4 mod 16: conflict-free DMMA fragment loads
constexpr int CK4 = 128;         // phase 4: columns per chunk (one array, double buffered)
constexpr int LDX = CK4 + 4;     // 132 = 4 mod 16
constexpr int LDS_ = B2 + 1;     // S and J: [B2][LDS_]
constexpr int LDJ = B2 + 4;      // J^T: [B2][LDJ], 36 = 4 mod 16
constexpr int BUF_DOUBLES = 4 * B2 * LDH;  // 8704 >= 2 * B2 * LDX (8448) and >= 8 * B2 * B2 (8192, partial H)
constexpr size_t BJ_SMEM = sizeof(double) * ((size_t)BUF_DOUBLES + 2 * B2 * LDS_ + B2 * LDJ + 64);

struct EighSys {
    double* G;
    double* Vt;
    int ldg, ldv, nblk;  // nblk even (>= 2): number of 16-row blocks, padding blocks carry the identity
    double floor1;       // 2 sqrt(n) eps |A|_F
    int* nrot;
    int inner;           // cyclic sweeps of the 32 x 32 sub-problem per visit (stops early once nothing rotates)
};
struct EighBatch {
    EighSys s[MAXB];
};

__device__ __forceinline__ void circle_pair(int npl, int round, int k, int& p, int& q) {
    const int mod = npl - 1;
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
}

__global__ void __launch_bounds__(256, 2) k_jacobi_block_round(EighBatch bt, int round) {
    extern __shared__ __align__(16) double sm[];
    const EighSys& sy = bt.s[blockIdx.y];
    const int npl = sy.nblk;
    if (npl < 2 || (int)blockIdx.x >= npl / 2 || round >= npl - 1) return;
    int P, Q;
    circle_pair(npl, round, blockIdx.x, P, Q);
    const int ncol = npl * BJ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* buf = sm;                      // chunk stages, later the 8 partial H
    double* S = buf + BUF_DOUBLES;         // [32][33]
    double* J = S + B2 * LDS_;             // [32][33]
    double* JT = J + B2 * LDS_;            // [32][36]  J^T, the A operand of phase 4
    double* cs = JT + B2 * LDJ;            // [16][2] rotation of each pair of the inner round (c, s); s = 0: skip
    __shared__ int s_any;
    auto grow = [&](int r) { return (r < BJ ? P * BJ + r : Q * BJ + (r - BJ)); };

    // ---- phase 1: H = G_PQ V_PQ^T on the FP64 tensor pipe.  64-column chunks of the 32 G rows and the 32 V rows are
    // double buffered with cp.async; warp w takes columns 8w .. 8w+7 of every chunk and accumulates the full 32x32
    // (4 x 4 m8n8 tiles); the eight partial H are added in warp order afterwards.
    {
        auto load1 = [&](int stage, int c0) {
            double* Gs = buf + (2 * stage) * B2 * LDH;
            double* Vs = Gs + B2 * LDH;
            for (int e = tid; e < B2 * (CK1 / 2); e += 256) {
                const int r = e / (CK1 / 2), k2 = (e - r * (CK1 / 2)) * 2;
                const int c = c0 + k2;
                const size_t gr = (size_t)grow(r);
                if (c < ncol) {
                    cp_async16(Gs + r * LDH + k2, sy.G + gr * sy.ldg + c);
                    cp_async16(Vs + r * LDH + k2, sy.Vt + gr * sy.ldv + c);
                } else {
                    *reinterpret_cast<double2*>(Gs + r * LDH + k2) = make_double2(0.0, 0.0);
                    *reinterpret_cast<double2*>(Vs + r * LDH + k2) = make_double2(0.0, 0.0);
                }
            }
        };
        double acc[4][4][2];
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        const int nch = (ncol + CK1 - 1) / CK1;
        load1(0, 0);
        cp_async_commit();
        for (int ch = 0; ch < nch; ch++) {
            if (ch + 1 < nch) load1((ch + 1) & 1, (ch + 1) * CK1);
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            const double* Gs = buf + (2 * (ch & 1)) * B2 * LDH + g * LDH + 8 * warp + q;
            const double* Vs = Gs + B2 * LDH;
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; mi++) a[mi] = Gs[mi * 8 * LDH + kk * 4];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) b[ni] = Vs[ni * 8 * LDH + kk * 4];
#pragma unroll
                for (int mi = 0; mi < 4; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
            __syncthreads();
        }
        cp_async_wait<0>();
        double* Hw = buf + warp * B2 * B2;
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                Hw[(8 * mi + g) * B2 + 8 * ni + 2 * q] = acc[mi][ni][0];
                Hw[(8 * mi + g) * B2 + 8 * ni + 2 * q + 1] = acc[mi][ni][1];
            }
    }
    __syncthreads();
    // ---- phase 2: S = (H + H^T) / 2, J = I; anything to do?
    if (tid == 0) s_any = 0;
    for (int e = tid; e < B2 * B2; e += 256) {
        const int a = e >> 5, b = e & 31;
        double h = 0.0, ht = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            h += buf[(w * B2 + a) * B2 + b];
            ht += buf[(w * B2 + b) * B2 + a];
        }
        S[a * LDS_ + b] = 0.5 * (h + ht);
        J[a * LDS_ + b] = (a == b) ? 1.0 : 0.0;
    }
    __syncthreads();
    {
        int act = 0;
        for (int e = tid; e < B2 * B2; e += 256) {
            const int a = e >> 5, b = e & 31;
            if (a != b && fabs(S[a * LDS_ + b]) > sy.floor1) act = 1;
        }
        if (act) s_any = 1;  // benign race: every writer stores 1
    }
    __syncthreads();
    if (!s_any) return;
    // ---- phase 3: one cyclic sweep of two-sided Jacobi on S (32 x 32), rotations accumulated in J.
    // 16 disjoint pairs per inner round, 16 threads per pair.
    const int pr = tid >> 4, sub = tid & 15;
    int napplied = 0;
    for (int isw = 0; isw < sy.inner; isw++) {
    int applied_now = 0;
    for (int rnd = 0; rnd < B2 - 1; rnd++) {
        int p, q;
        circle_pair(B2, rnd, pr, p, q);
        if (sub == 0) {
            const double a = S[p * LDS_ + p], b = S[q * LDS_ + q], g = S[p * LDS_ + q];
            double c = 1.0, s = 0.0;
            if (fabs(g) > sy.floor1) {
                const double zeta = (b - a) / (2.0 * g);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                c = 1.0 / sqrt(1.0 + t * t);
                s = c * t;
            }
            cs[2 * pr] = c;
            cs[2 * pr + 1] = s;
        }
        __syncthreads();
        const double c = cs[2 * pr], s = cs[2 * pr + 1];
        if (s != 0.0) {  // rows p, q:  S <- R^T S
            napplied = 1;
            applied_now = 1;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int col = sub + 16 * h;
                const double x = S[p * LDS_ + col], y = S[q * LDS_ + col];
                S[p * LDS_ + col] = c * x - s * y;
                S[q * LDS_ + col] = s * x + c * y;
            }
        }
        __syncthreads();
        if (s != 0.0) {  // columns p, q:  S <- S R,  J <- J R
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int row = sub + 16 * h;
                const double x = S[row * LDS_ + p], y = S[row * LDS_ + q];
                S[row * LDS_ + p] = c * x - s * y;
                S[row * LDS_ + q] = s * x + c * y;
                const double u = J[row * LDS_ + p], w = J[row * LDS_ + q];
                J[row * LDS_ + p] = c * u - s * w;
                J[row * LDS_ + q] = s * u + c * w;
            }
        }
        __syncthreads();
    }
    if (!__syncthreads_or(applied_now)) break;  // the sub-problem is diagonal to the rounding floor
    }
    if (!__syncthreads_or(napplied)) return;
    if (tid == 0) atomicAdd(sy.nrot, 1);
    for (int e = tid; e < B2 * B2; e += 256) {
        const int a = e >> 5, b = e & 31;
        JT[a * LDJ + b] = J[b * LDS_ + a];
    }
    __syncthreads();
    // ---- phase 4: rows <- J^T rows for G and for V^T, on the tensor pipe.  128-column chunks of the 32 rows are
    // double buffered with cp.async; warp w owns columns 16w .. 16w+15 of a chunk (4 x 2 m8n8 tiles, K = 32);
    // the A fragments (J^T) stay in registers for the whole pass.
    double areg[4][8];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int kk = 0; kk < 8; kk++) areg[mi][kk] = JT[(8 * mi + g) * LDJ + 4 * kk + q];
    const int nch4 = (ncol + CK4 - 1) / CK4;
    auto load4 = [&](int stage, int t) {  // t enumerates (array, chunk)
        const int arr = t / nch4, c0 = (t - arr * nch4) * CK4;
        const double* base = arr ? sy.Vt : sy.G;
        const int ld = arr ? sy.ldv : sy.ldg;
        double* X = buf + stage * B2 * LDX;
        for (int e = tid; e < B2 * (CK4 / 2); e += 256) {
            const int r = e / (CK4 / 2), k2 = (e - r * (CK4 / 2)) * 2;
            const int c = c0 + k2;
            if (c < ncol)
                cp_async16(X + r * LDX + k2, base + (size_t)grow(r) * ld + c);
            else
                *reinterpret_cast<double2*>(X + r * LDX + k2) = make_double2(0.0, 0.0);
        }
    };
    const int nt = 2 * nch4;
    load4(0, 0);
    cp_async_commit();
    for (int t = 0; t < nt; t++) {
        if (t + 1 < nt) load4((t + 1) & 1, t + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double* X = buf + (t & 1) * B2 * LDX + q * LDX + 16 * warp + g;
        double o[4][2][2];
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 2; ni++) o[mi][ni][0] = o[mi][ni][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
            const double b0 = X[kk * 4 * LDX], b1 = X[kk * 4 * LDX + 8];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) {
                dmma884(o[mi][0][0], o[mi][0][1], areg[mi][kk], b0);
                dmma884(o[mi][1][0], o[mi][1][1], areg[mi][kk], b1);
            }
        }
        const int arr = t / nch4, c0 = (t - arr * nch4) * CK4;
        double* base = arr ? sy.Vt : sy.G;
        const int ld = arr ? sy.ldv : sy.ldg;
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 2; ni++) {
                const int c = c0 + 16 * warp + 8 * ni + 2 * q;
                if (c < ncol)
                    *reinterpret_cast<double2*>(base + (size_t)grow(8 * mi + g) * ld + c) =
                        make_double2(o[mi][ni][0], o[mi][ni][1]);
            }
        __syncthreads();
    }
    cp_async_wait<0>();
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

int launch_jacobi_eigh_batch(const EighProblem* pr, int nsys, int max_sweeps, int* sweeps_done, cudaStream_t st) {
    if (sweeps_done) *sweeps_done = 0;
    if (nsys <= 0) return 0;
    B200_REQUIRE(nsys <= MAXB, "at most MAXB eigenproblems per batched call");
    static bool attr = false;
    if (!attr) {
        B200_CUDA(cudaFuncSetAttribute(k_jacobi_block_round, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BJ_SMEM));
        attr = true;
    }
    // device state per system: |A|_F^2, nmax per-row norms; then the nsys rotation counters
    int nmax = 0;
    for (int q = 0; q < nsys; q++) nmax = pr[q].n > nmax ? pr[q].n : nmax;
    if (nmax <= 0) return 0;
    const size_t per = (size_t)nmax + 2;
    void* wsv = nullptr;
    if (int rc = scratch(7, sizeof(double) * per * nsys + sizeof(int) * MAXB + 64, &wsv)) return rc;
    double* state = (double*)wsv;
    int* nrot = (int*)(state + per * nsys);
    EighBatch bt;
    int nblk_max = 0;
    for (int q = 0; q < nsys; q++) {
        const EighProblem& p = pr[q];
        int nblk = (p.n + BJ - 1) / BJ;
        nblk += nblk & 1;
        if (nblk < 2) nblk = 2;
        const int ntot = nblk * BJ;
        B200_REQUIRE(p.n > 0 && p.lda >= ntot && p.ldv >= ntot && p.lda % 2 == 0 && p.ldv % 2 == 0,
                     "eigh: A and Vt must be padded (identity) to a multiple of 32 rows/columns, even leading dimensions");
        double* st_q = state + per * q;
        k_jacobi_init<<<ntot, 256, 0, st>>>(p.Vt, p.ldv, p.n, ntot, p.A, p.lda, st_q + 1);
        k_jacobi_fro<<<1, 256, 0, st>>>(st_q + 1, p.n, st_q);
        B200_LAUNCHED(2);
        bt.s[q].G = p.A;
        bt.s[q].Vt = p.Vt;
        bt.s[q].ldg = p.lda;
        bt.s[q].ldv = p.ldv;
        bt.s[q].nblk = p.n > 1 ? nblk : 0;
        bt.s[q].nrot = nrot + q;
        nblk_max = nblk > nblk_max ? nblk : nblk_max;
    }
    for (int q = nsys; q < MAXB; q++) {
        bt.s[q] = bt.s[0];
        bt.s[q].nblk = 0;
    }
    B200_CUDA(cudaGetLastError());
    double fro2_h[MAXB];
    for (int q = 0; q < nsys; q++)
        B200_CUDA(cudaMemcpyAsync(&fro2_h[q], state + per * q, sizeof(double), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    double floor_mult = 1.0;
    if (const char* e = getenv("B200_EIGH_FLOOR_MULT")) floor_mult = atof(e);  // experiment knob
    int inner = EIGH_INNER_DEFAULT;
    if (const char* e = getenv("B200_EIGH_INNER")) inner = atoi(e) > 0 ? atoi(e) : 1;  // experiment knob
    for (int q = 0; q < MAXB; q++) bt.s[q].inner = inner;
    for (int q = 0; q < nsys; q++)
        bt.s[q].floor1 = floor_mult * 2.0 * 2.220446049250313e-16 * sqrt((double)pr[q].n * fro2_h[q]);
    int sweep = 0;
    prof_begin(PROF_EIGH, st);
    for (; sweep < max_sweeps; sweep++) {
        bool live = false;
        for (int q = 0; q < nsys; q++) live = live || bt.s[q].nblk > 0;
        if (!live) break;
        B200_CUDA(cudaMemsetAsync(nrot, 0, sizeof(int) * MAXB, st));
        for (int r = 0; r < nblk_max - 1; r++)
            k_jacobi_block_round<<<dim3(nblk_max / 2, nsys), 256, BJ_SMEM, st>>>(bt, r);
        B200_LAUNCHED(nblk_max - 1);
        B200_CUDA(cudaGetLastError());
        int h[MAXB];
        B200_CUDA(cudaMemcpyAsync(h, nrot, sizeof(int) * nsys, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        for (int q = 0; q < nsys; q++)
            if (h[q] == 0) bt.s[q].nblk = 0;  // converged: its CTAs exit at once from now on
    }
    prof_end(0.0, st);
    if (sweeps_done) *sweeps_done = sweep;
    for (int q = 0; q < nsys; q++) {
        k_jacobi_finish<<<pr[q].n, JT, 0, st>>>(pr[q].A, pr[q].lda, pr[q].Vt, pr[q].ldv, pr[q].n, pr[q].lam);
        B200_LAUNCHED(1);
    }
    B200_CUDA(cudaGetLastError());
    return 0;
}

int launch_jacobi_eigh(double* A, int lda, int n, double* Vt, int ldv, double* lam, int max_sweeps, int* sweeps_done,
                       cudaStream_t st) {
    if (sweeps_done) *sweeps_done = 0;
    if (n <= 0) return 0;
    EighProblem p;
    p.A = A;
    p.Vt = Vt;
    p.lam = lam;
    p.lda = lda;
    p.ldv = ldv;
    p.n = n;
    return launch_jacobi_eigh_batch(&p, 1, max_sweeps, sweeps_done, st);
}

}  // namespace b200
