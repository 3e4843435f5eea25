// FP64 dense linear algebra for the CholKernel path: a DMMA (mma.sync m8n8k4 f64)
// tile GEMM, a shared-memory Cholesky + triangular inverse of one 128x128 diagonal block, and the
// batched right-looking driver that factors  W = A + kappa*I = L L^T  and solves
// Ti (L L^T) = mBhalf  for all m right-hand-side rows.
//
// Replaces scipy.linalg.cholesky / cho_solve at lakernel.py:263, 276, 304, 358.
//
// Data layout (all row-major, leading dimensions and padded sizes multiples of NB = 128):
//   W (npad, ldw)  in: A + kappa*I (lower triangle read, identity in the padding)
//                  out: L in the lower triangle (diagonal blocks included), L^T in the strictly upper
//                       block triangle (written by the panel step so that every product below is "NT").
//   X (mpad, ldx)  in: mBhalf rows; out: Ti rows.
//   Dinv (2*nb, 128, 128): inv(L_kk) for k < nb, then inv(L_kk)^T.
//
// With rows as right-hand sides the solve is   Z = X L^-T  (forward, fused into the factorisation:
// the X rows are simply extra panel rows below the matrix) followed by  Ti = Z L^-1  (backward).
// Every product is  C (+)= A[.,K] * B[.,K]^T  with both operands K-contiguous, so one tile routine serves all
// phases: gemm_tile64 (64x64 output, 4 warps of 32x32, K stages of 16 in a 4-deep cp.async ring, three CTAs per SM,
// DMMA.8x8x4).  gemm_tile_nt is the 128x128 / 256-thread tile it replaced, kept selectable (B200_TILE64=0).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "ozaki.cuh"

namespace b200 {

namespace {

constexpr int KT = 32;           // K extent of one pipeline stage
constexpr int LDSM = KT + 4;     // smem row stride in doubles: 20 = 4 mod 16 -> conflict-free fragment loads
constexpr int STAGES = 3;
// Tried and dropped (round 1): splitting the tile between the tensor pipe and the FP64 FMA pipe.  The two pipes do
// issue concurrently on sm_100a (tools/probe/pipe_probe.cu: 8 DMMA warps per SM sustain 37 TFLOP/s and DFMA warps
// added next to them cost nothing; DFMA alone peaks at 24.6 TFLOP/s), but two extra FMA warps computing the last 16
// (32) columns of the tile from the same staged operands brought DGEMM 8192^3 from 32.2 down to 24.6 (19.4) TFLOP/s:
// at one DFMA per 3 clk per scheduler they cannot retire their share inside the tensor warps' k-tile time and become
// the critical path at every barrier.
#ifndef B200_GEMM_WARPS_M
#define B200_GEMM_WARPS_M 2
#endif
constexpr int WARPS_M = B200_GEMM_WARPS_M, WARPS_N = 4;  // warp grid over the 128x128 tile
constexpr int GT = 32 * WARPS_M * WARPS_N;               // threads per GEMM CTA
constexpr int NI = NB / WARPS_N / 8;  // n8 fragments per warp tile (m8 fragments: 8 for a full tile, 4 for a half tile)
constexpr int SP_DEFAULT = 4;    // block columns per super-panel (512 matrix columns); B200_SP overrides (tuning)
constexpr int STAGE_DOUBLES = 2 * NB * LDSM;
constexpr size_t GEMM_SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);  // 163840 B

enum TileMode { TILE_ASSIGN = 0, TILE_SUB = 1, TILE_ADD = 2 };

// One 128x128 output tile:  C = op(C, A[128,K] * B[128,K]^T).
//   TILE_ASSIGN: C = A B^T       (C may alias A: every A chunk is in shared memory before C is written)
//   TILE_SUB   : C = C - A B^T   (accumulators start at -C, result is the negated accumulator)
//   TILE_ADD   : C = C + A B^T
// Ct != nullptr additionally stores the tile transposed: Ct[c * ldct + r] = C[r][c].
// HALF: only rows 0..63 of the tile exist as work (the last row-tile of the right-hand sides when m mod 128 <= 64: rows
// 64..127 are zero padding and stay zero); the eight warps then share a 64 x 128 tile (warp tile 32 x 32) and the A
// operand is staged for 64 rows only, so the tile costs half the tensor instructions instead of idling half the warps.
template <int MODE, bool HALF = false>
__device__ __forceinline__ void gemm_tile_nt(const double* A, int lda, const double* B,
                                             int ldb, double* C, int ldc, int K, double* Ct, int ldct,
                                             double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;  // warp tile (8 MI) x (8 NI)
    const int g = lane >> 2, q = lane & 3;
    constexpr int TM = HALF ? NB / 2 : NB;
    constexpr int MI = TM / WARPS_M / 8;
    constexpr int WTM = 8 * MI, WTN = 8 * NI;

    double acc[MI][NI][2];
    if (MODE == TILE_ASSIGN) {
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }

    const int nk = (K + KT - 1) / KT;
    // loader mapping: 1024 16-byte chunks per operand per stage; thread handles chunks tid + GT*r
    auto load_stage = [&](int slot, int kt) {
        double* As = smem + (size_t)slot * STAGE_DOUBLES;
        double* Bs = As + NB * LDSM;
        const int k0 = kt * KT;
#pragma unroll
        for (int r = 0; r < NB * (KT / 2) / GT; r++) {
            const int c = tid + GT * r;
            const int row = c / (KT / 2), kc = (c % (KT / 2)) * 2;
            if (k0 + kc < K) {
                if (!HALF || row < TM) cp_async16(As + row * LDSM + kc, A + (size_t)row * lda + k0 + kc);
                cp_async16(Bs + row * LDSM + kc, B + (size_t)row * ldb + k0 + kc);
            } else {  // K tail of the last stage (K is even): zero operands contribute nothing
                *reinterpret_cast<double2*>(As + row * LDSM + kc) = make_double2(0.0, 0.0);
                *reinterpret_cast<double2*>(Bs + row * LDSM + kc) = make_double2(0.0, 0.0);
            }
        }
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    if (MODE != TILE_ASSIGN) {  // overlaps with the pipeline fill
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                const double2 v = *reinterpret_cast<const double2*>(
                    C + (size_t)(wm * WTM + mi * 8 + g) * ldc + wn * WTN + ni * 8 + q * 2);
                acc[mi][ni][0] = (MODE == TILE_SUB) ? -v.x : v.x;
                acc[mi][ni][1] = (MODE == TILE_SUB) ? -v.y : v.y;
            }
    }
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < nk) load_stage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
        cp_async_commit();
        const double* As = smem + (size_t)(kt % STAGES) * STAGE_DOUBLES + (wm * WTM + g) * LDSM + q;
        const double* Bs = smem + (size_t)(kt % STAGES) * STAGE_DOUBLES + NB * LDSM + (wn * WTN + g) * LDSM + q;
#pragma unroll
        for (int kk = 0; kk < KT / 4; kk++) {
            double a[MI], b[NI];
#pragma unroll
            for (int mi = 0; mi < MI; mi++) a[mi] = As[mi * 8 * LDSM + kk * 4];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) b[ni] = Bs[ni * 8 * LDSM + kk * 4];
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            const int r = wm * WTM + mi * 8 + g, c = wn * WTN + ni * 8 + q * 2;
            double2 v;
            v.x = (MODE == TILE_SUB) ? -acc[mi][ni][0] : acc[mi][ni][0];
            v.y = (MODE == TILE_SUB) ? -acc[mi][ni][1] : acc[mi][ni][1];
            *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = v;
            if (Ct) {
                Ct[(size_t)c * ldct + r] = v.x;
                Ct[(size_t)(c + 1) * ldct + r] = v.y;
            }
        }
}

// ---- the 64x64 tile: several small independent CTAs per SM -------------------------------------
// One 128x128 CTA per SM leaves the tensor pipe idle whenever its eight warps meet at the stage barrier (ncu: DMMA pipe
// 87 % busy, 30.5 TFLOP/s).  The library DGEMM that sets the roofline (ncu on torch.matmul, tools/probe/dgemm_probe.py:
// cutlass d884gemm 64x64_16x4, 128 threads, 120 registers, 64 KB, three CTAs per SM, DMMA pipe 97.5 % busy) hides that
// barrier behind the other resident CTAs.  Same recipe here: 64x64 output tile, 4 warps of 32x32, K stages of 16 in a
// 4-deep cp.async ring (64 KB, so three CTAs share an SM), rows of one stage are exactly 128 B and the four 32-byte
// granules of a row are XOR-swizzled with (row & 3): cp.async keeps its 16-byte granularity, and a fragment load (8 rows x
// 4 consecutive doubles per instruction = one 32-byte granule per row) is served as two half-warps of 4 rows whose
// granules fall on four different bank groups - conflict-free.  (Round 1 swizzled the 16-byte chunks with (row & 7),
// TMA's 128-byte pattern: rows g and g ^ 1 of a half-warp then share their two chunks, a two-way conflict on every
// fragment load - ncu: 43 % of the shared wavefronts.)
// Used by the four update kernels (operands and output never overlap there); the in-place products with inv(L_kk)
// keep the 128x128 tile, which stages the whole K = 128 operand before it writes.
#ifndef B200_T64_STAGES
#define B200_T64_STAGES 4
#endif
#ifndef B200_T64_CTAS
#define B200_T64_CTAS 3
#endif
constexpr int KT2 = 16, ST2 = B200_T64_STAGES, TB = 64, GT2 = 128;
constexpr int CTAS2 = B200_T64_CTAS;  // resident CTAs per SM the 64x64-tile kernels are compiled for
constexpr int STAGE2_DOUBLES = 2 * TB * KT2;
constexpr size_t GEMM2_SMEM = (size_t)ST2 * STAGE2_DOUBLES * sizeof(double);  // 65536 B

template <int MODE>
__device__ __forceinline__ void gemm_tile64(const double* A, int lda, const double* B, int ldb, double* C, int ldc,
                                            int K, double* smem, double* Ct = nullptr, int ldct = 0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;  // 2 x 2 warps of 32 x 32
    const int g = lane >> 2, q = lane & 3;
    constexpr int MI = 4, NJ = 4;

    double acc[MI][NJ][2];
    if (MODE == TILE_ASSIGN) {
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int ni = 0; ni < NJ; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }
    const int nk = (K + KT2 - 1) / KT2;
    // loader: 512 16-byte chunks per operand per stage, 4 per thread; 8 consecutive threads fetch one 128-byte row
    auto load_stage = [&](int slot, int kt) {
        double* As = smem + (size_t)slot * STAGE2_DOUBLES;
        double* Bs = As + TB * KT2;
        const int k0 = kt * KT2;
#pragma unroll
        for (int r = 0; r < TB * (KT2 / 2) / GT2; r++) {
            const int c = tid + GT2 * r;
            const int row = c >> 3, ch = c & 7;
            const int dst = row * KT2 + (((((ch >> 1) ^ (row & 3)) << 1) | (ch & 1)) << 1);
            if (k0 + 2 * ch < K) {
                cp_async16(As + dst, A + (size_t)row * lda + k0 + 2 * ch);
                cp_async16(Bs + dst, B + (size_t)row * ldb + k0 + 2 * ch);
            } else {  // K tail (K is even): zero operands contribute nothing
                *reinterpret_cast<double2*>(As + dst) = make_double2(0.0, 0.0);
                *reinterpret_cast<double2*>(Bs + dst) = make_double2(0.0, 0.0);
            }
        }
    };
#pragma unroll
    for (int s = 0; s < ST2 - 1; s++) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    if (MODE != TILE_ASSIGN) {  // overlaps with the pipeline fill
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int ni = 0; ni < NJ; ni++) {
                const double2 v = *reinterpret_cast<const double2*>(
                    C + (size_t)(wm * 32 + mi * 8 + g) * ldc + wn * 32 + ni * 8 + q * 2);
                acc[mi][ni][0] = (MODE == TILE_SUB) ? -v.x : v.x;
                acc[mi][ni][1] = (MODE == TILE_SUB) ? -v.y : v.y;
            }
    }
    // element (row, 4 kk + q) of a stage sits at row * 16 + (off0 ^ 4 kk): all fragment rows have row & 3 == g & 3
    const int off0 = ((g & 3) << 2) | q;
    const int arow = (wm * 32 + g) * KT2, brow = TB * KT2 + (wn * 32 + g) * KT2;
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<ST2 - 2>();
        __syncthreads();
        if (kt + ST2 - 1 < nk) load_stage((kt + ST2 - 1) % ST2, kt + ST2 - 1);
        cp_async_commit();
        const double* St = smem + (size_t)(kt % ST2) * STAGE2_DOUBLES;
#pragma unroll
        for (int kk = 0; kk < KT2 / 4; kk++) {
            const double* As = St + arow + (off0 ^ (4 * kk));
            const double* Bs = St + brow + (off0 ^ (4 * kk));
            double a[MI], b[NJ];
#pragma unroll
            for (int mi = 0; mi < MI; mi++) a[mi] = As[mi * 8 * KT2];
#pragma unroll
            for (int ni = 0; ni < NJ; ni++) b[ni] = Bs[ni * 8 * KT2];
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NJ; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NJ; ni++) {
            const int r = wm * 32 + mi * 8 + g, c = wn * 32 + ni * 8 + q * 2;
            double2 v;
            v.x = (MODE == TILE_SUB) ? -acc[mi][ni][0] : acc[mi][ni][0];
            v.y = (MODE == TILE_SUB) ? -acc[mi][ni][1] : acc[mi][ni][1];
            *reinterpret_cast<double2*>(C + (size_t)r * ldc + c) = v;
            if (Ct) {
                Ct[(size_t)c * ldct + r] = v.x;
                Ct[(size_t)(c + 1) * ldct + r] = v.y;
            }
        }
}

// In-place product of a 64-row strip with one of the triangular 128x128 inverses:  C = C D^T, C = Cp[0:64, 0:128].
//   LOWER  (panel step, D = inv(L_kk), D[c][k] = 0 for k > c): columns 64..127 need all 128 k, columns 0..63 only k < 64;
//   !LOWER (backward step, D = inv(L_kk)^T, D[c][k] = 0 for k < c): columns 0..63 need all k, columns 64..127 only k >= 64.
// Two passes of the 64x64 tile in that order: the first pass has every operand chunk in shared memory before it writes
// its own columns, and the second pass never reads the columns the first one wrote.  The structural zeros are skipped
// (a quarter of the products), which does not change any sum.
template <bool LOWER>
__device__ __forceinline__ void trsm_strip64(double* Cp, int ldc, const double* D, double* Ct, int ldct, double* smem) {
    if (LOWER) {
        gemm_tile64<TILE_ASSIGN>(Cp, ldc, D + (size_t)TB * NB, NB, Cp + TB, ldc, NB, smem,
                                 Ct ? Ct + (size_t)TB * ldct : nullptr, ldct);
        __syncthreads();
        gemm_tile64<TILE_ASSIGN>(Cp, ldc, D, NB, Cp, ldc, TB, smem, Ct, ldct);
    } else {
        gemm_tile64<TILE_ASSIGN>(Cp, ldc, D, NB, Cp, ldc, NB, smem);
        __syncthreads();
        gemm_tile64<TILE_ASSIGN>(Cp + TB, ldc, D + (size_t)TB * NB + TB, NB, Cp + TB, ldc, TB, smem);
    }
}

// One 128x128 output tile of an update kernel, either by one 256-thread CTA (T64 = false) or as four 64x64 quadrants
// (qm, qn) by four 128-thread CTAs.  half: only rows 0..63 of the tile are real (x_half_tile).
template <int MODE, bool T64>
__device__ __forceinline__ void update_tile(bool half, int qm, int qn, const double* A, int lda, const double* B,
                                            int ldb, double* C, int ldc, int K, double* smem) {
    if (T64) {
        if (half && qm) return;
        gemm_tile64<MODE>(A + (size_t)qm * TB * lda, lda, B + (size_t)qn * TB * ldb, ldb,
                          C + (size_t)qm * TB * ldc + qn * TB, ldc, K, smem);
    } else if (half) {
        gemm_tile_nt<MODE, true>(A, lda, B, ldb, C, ldc, K, nullptr, 0, smem);
    } else {
        gemm_tile_nt<MODE, false>(A, lda, B, ldb, C, ldc, K, nullptr, 0, smem);
    }
}

// ---- generic C (+)= A B^T over a grid of tiles (also the public DGEMM of the library) ----------
template <int MODE>
__global__ void __launch_bounds__(GT, 1) k_gemm_nt(const double* __restrict__ A, int lda,
                                                   const double* __restrict__ B, int ldb, double* C, int ldc, int K) {
    extern __shared__ __align__(16) double smem[];
    const int tn = blockIdx.x, tm = blockIdx.y;
    gemm_tile_nt<MODE>(A + (size_t)tm * NB * lda, lda, B + (size_t)tn * NB * ldb, ldb,
                       C + (size_t)tm * NB * ldc + (size_t)tn * NB, ldc, K, nullptr, 0, smem);
}

template <int MODE>
__global__ void __launch_bounds__(GT2, CTAS2) k_gemm64_nt(const double* __restrict__ A, int lda,
                                                      const double* __restrict__ B, int ldb, double* C, int ldc, int K) {
    extern __shared__ __align__(16) double smem[];
    const int tn = blockIdx.x, tm = blockIdx.y;
    gemm_tile64<MODE>(A + (size_t)tm * TB * lda, lda, B + (size_t)tn * TB * ldb, ldb,
                      C + (size_t)tm * TB * ldc + (size_t)tn * TB, ldc, K, smem);
}

// The same tile over several independent products at once (blockIdx.z = problem): the eigensolver's Gram matrices and
// compact-WY panel products of all systems of a batch share each launch.
template <int MODE>
__global__ void __launch_bounds__(GT2, CTAS2) k_gemm64_nt_batch(GemmBatch gb) {
    extern __shared__ __align__(16) double smem[];
    const GemmProb& g = gb.p[blockIdx.z];
    const int tn = blockIdx.x, tm = blockIdx.y;
    if (tm * TB >= g.M || tn * TB >= g.N) return;
    gemm_tile64<MODE>(g.A + (size_t)tm * TB * g.lda, g.lda, g.B + (size_t)tn * TB * g.ldb, g.ldb,
                      g.C + (size_t)tm * TB * g.ldc + (size_t)tn * TB, g.ldc, g.K, smem);
}

// Row-tile r of the right-hand sides holds at most 64 real rows (mrows = number of real rows; 0 = unknown: all tiles full).
__device__ __forceinline__ bool x_half_tile(const SolveSys& s, int r) {
    return s.mrows > 0 && s.mrows - r * NB <= NB / 2;
}

// ---- factorisation phases ----------------------------------------------------------------------
// The driver works on SUPER-PANELS of SP block columns [c0, c1): left-looking between super-panels (one
// long-K GEMM launch brings the whole super-panel up to date: every output tile is read and written once
// per super-panel, so the tile kernel runs at its large-K rate), right-looking with K = 128 inside.
//
// Super-panel update:  W[i][j] -= W[i][0:c0] W[j][0:c0]^T  (c0 <= j < c1, j <= i)  and the same for the X rows.
template <bool T64>
__global__ void __launch_bounds__(T64 ? GT2 : GT, T64 ? CTAS2 : 1) k_chol_super_update(SolveBatch bt, int c0, int c1) {
    extern __shared__ __align__(16) double smem[];
    const SolveSys& s = bt.s[blockIdx.z];
    const int nb = s.npad / NB, mb = s.mpad / NB;
    const int qn = T64 ? blockIdx.x & 1 : 0, qm = T64 ? blockIdx.y & 1 : 0;
    const int j = c0 + (T64 ? blockIdx.x >> 1 : blockIdx.x), t = T64 ? blockIdx.y >> 1 : blockIdx.y;
    const int nrow = nb - c0;
    if (j >= nb || j >= c1 || t >= nrow + mb) return;
    const double* Bp = s.W + (size_t)j * NB * s.ldw;
    const int K = c0 * NB;
    if (t < nrow) {
        const int i = c0 + t;
        if (j > i || (T64 && i == j && qn > qm)) return;  // (the upper quadrant of a diagonal tile is never read)
        update_tile<TILE_SUB, T64>(false, qm, qn, s.W + (size_t)i * NB * s.ldw, s.ldw, Bp, s.ldw,
                                   s.W + (size_t)i * NB * s.ldw + (size_t)j * NB, s.ldw, K, smem);
    } else {
        const int r = t - nrow;
        update_tile<TILE_SUB, T64>(x_half_tile(s, r), qm, qn, s.X + (size_t)r * NB * s.ldx, s.ldx, Bp, s.ldw,
                                   s.X + (size_t)r * NB * s.ldx + (size_t)j * NB, s.ldx, K, smem);
    }
}

// Panel step k: rows below the diagonal block (and all X rows) times inv(L_kk)^T; the W part is also
// stored transposed into the upper block triangle.
template <bool T64>
__global__ void __launch_bounds__(T64 ? GT2 : GT, T64 ? CTAS2 : 1) k_chol_panel(SolveBatch bt, int k) {
    extern __shared__ __align__(16) double smem[];
    const SolveSys& s = bt.s[blockIdx.y];
    const int nb = s.npad / NB, mb = s.mpad / NB;
    const int nrow = nb - 1 - k;
    const int t = T64 ? blockIdx.x >> 1 : blockIdx.x, qm = T64 ? blockIdx.x & 1 : 0;
    if (k >= nb || t >= nrow + mb) return;
    const double* Dk = s.Dinv + (size_t)k * NB * NB;
    if (t < nrow) {
        const int i = k + 1 + t;
        double* Cp = s.W + (size_t)i * NB * s.ldw + (size_t)k * NB;
        double* Ctp = s.W + (size_t)k * NB * s.ldw + (size_t)i * NB;
        if (T64)
            trsm_strip64<true>(Cp + (size_t)qm * TB * s.ldw, s.ldw, Dk, Ctp + qm * TB, s.ldw, smem);
        else
            gemm_tile_nt<TILE_ASSIGN>(Cp, s.ldw, Dk, NB, Cp, s.ldw, NB, Ctp, s.ldw, smem);
    } else {
        double* Cp = s.X + (size_t)(t - nrow) * NB * s.ldx + (size_t)k * NB;
        const bool half = x_half_tile(s, t - nrow);
        if (T64) {
            if (half && qm) return;
            trsm_strip64<true>(Cp + (size_t)qm * TB * s.ldx, s.ldx, Dk, nullptr, 0, smem);
        } else if (half) {
            gemm_tile_nt<TILE_ASSIGN, true>(Cp, s.ldx, Dk, NB, Cp, s.ldx, NB, nullptr, 0, smem);
        } else {
            gemm_tile_nt<TILE_ASSIGN>(Cp, s.ldx, Dk, NB, Cp, s.ldx, NB, nullptr, 0, smem);
        }
    }
}

// Update inside the super-panel after panel k:  W[i][j] -= W[i][k] W[j][k]^T (k < j < c1, j <= i)  and
// X[t][j] -= X[t][k] W[j][k]^T.
template <bool T64>
__global__ void __launch_bounds__(T64 ? GT2 : GT, T64 ? CTAS2 : 1) k_chol_update(SolveBatch bt, int k, int c1) {
    extern __shared__ __align__(16) double smem[];
    const SolveSys& s = bt.s[blockIdx.z];
    const int nb = s.npad / NB, mb = s.mpad / NB;
    const int nrow = nb - 1 - k;
    const int qn = T64 ? blockIdx.x & 1 : 0, qm = T64 ? blockIdx.y & 1 : 0;
    const int jj = T64 ? blockIdx.x >> 1 : blockIdx.x, t = T64 ? blockIdx.y >> 1 : blockIdx.y;
    const int j = k + 1 + jj;
    if (k >= nb || j >= nb || j >= c1 || t >= nrow + mb) return;
    const double* Bp = s.W + (size_t)j * NB * s.ldw + (size_t)k * NB;
    if (t < nrow) {
        const int i = k + 1 + t;
        if (j > i || (T64 && i == j && qn > qm)) return;
        update_tile<TILE_SUB, T64>(false, qm, qn, s.W + (size_t)i * NB * s.ldw + (size_t)k * NB, s.ldw, Bp, s.ldw,
                                   s.W + (size_t)i * NB * s.ldw + (size_t)j * NB, s.ldw, NB, smem);
    } else {
        const int r = t - nrow;
        update_tile<TILE_SUB, T64>(x_half_tile(s, r), qm, qn, s.X + (size_t)r * NB * s.ldx + (size_t)k * NB, s.ldx, Bp,
                                   s.ldw, s.X + (size_t)r * NB * s.ldx + (size_t)j * NB, s.ldx, NB, smem);
    }
}

// Backward substitution Ti = Z L^-1, super-panels taken from the right.  In "from the end" coordinates
// [e0, e1) the super-panel of a system with nb blocks is lo = max(0, nb - e1) <= k < hi = nb - e0.
// Super-panel update:  X[t][j] -= X[t][hi:nb] L[hi:nb][j]  (lo <= j < hi); L[k][j]^T lives at W[j][k] (upper triangle).
template <bool T64>
__global__ void __launch_bounds__(T64 ? GT2 : GT, T64 ? CTAS2 : 1) k_back_super_update(SolveBatch bt, int e0, int e1) {
    extern __shared__ __align__(16) double smem[];
    const SolveSys& s = bt.s[blockIdx.z];
    const int nb = s.npad / NB, mb = s.mpad / NB;
    const int hi = nb - e0;
    const int qn = T64 ? blockIdx.x & 1 : 0, qm = T64 ? blockIdx.y & 1 : 0;
    const int j = hi - 1 - (int)(T64 ? blockIdx.x >> 1 : blockIdx.x), t = T64 ? blockIdx.y >> 1 : blockIdx.y;
    if (hi <= 0 || j < 0 || j < nb - e1 || t >= mb) return;
    update_tile<TILE_SUB, T64>(x_half_tile(s, t), qm, qn, s.X + (size_t)t * NB * s.ldx + (size_t)hi * NB, s.ldx,
                               s.W + (size_t)j * NB * s.ldw + (size_t)hi * NB, s.ldw,
                               s.X + (size_t)t * NB * s.ldx + (size_t)j * NB, s.ldx, e0 * NB, smem);
}

// Backward step k, part 1:  X[t][k] = X[t][k] inv(L_kk)   (B operand = inv(L_kk)^T)
template <bool T64>
__global__ void __launch_bounds__(T64 ? GT2 : GT, T64 ? CTAS2 : 1) k_back_diag(SolveBatch bt, int kfromtop) {
    extern __shared__ __align__(16) double smem[];
    const SolveSys& s = bt.s[blockIdx.y];
    const int nb = s.npad / NB, mb = s.mpad / NB;
    const int k = nb - 1 - kfromtop;
    const int t = T64 ? blockIdx.x >> 1 : blockIdx.x, qm = T64 ? blockIdx.x & 1 : 0;
    if (k < 0 || t >= mb) return;
    double* Cp = s.X + (size_t)t * NB * s.ldx + (size_t)k * NB;
    const double* Dk = s.Dinv + (size_t)(nb + k) * NB * NB;
    const bool half = x_half_tile(s, t);
    if (T64) {
        if (half && qm) return;
        trsm_strip64<false>(Cp + (size_t)qm * TB * s.ldx, s.ldx, Dk, nullptr, 0, smem);
    } else if (half) {
        gemm_tile_nt<TILE_ASSIGN, true>(Cp, s.ldx, Dk, NB, Cp, s.ldx, NB, nullptr, 0, smem);
    } else {
        gemm_tile_nt<TILE_ASSIGN>(Cp, s.ldx, Dk, NB, Cp, s.ldx, NB, nullptr, 0, smem);
    }
}

// Backward step k, part 2, inside the super-panel:  X[t][j] -= X[t][k] L[k][j]  (lo <= j < k)
template <bool T64>
__global__ void __launch_bounds__(T64 ? GT2 : GT, T64 ? CTAS2 : 1) k_back_update(SolveBatch bt, int kfromtop, int e1) {
    extern __shared__ __align__(16) double smem[];
    const SolveSys& s = bt.s[blockIdx.z];
    const int nb = s.npad / NB, mb = s.mpad / NB;
    const int k = nb - 1 - kfromtop;
    const int qn = T64 ? blockIdx.x & 1 : 0, qm = T64 ? blockIdx.y & 1 : 0;
    const int j = k - 1 - (int)(T64 ? blockIdx.x >> 1 : blockIdx.x), t = T64 ? blockIdx.y >> 1 : blockIdx.y;
    if (k < 0 || j < 0 || j < nb - e1 || t >= mb) return;
    update_tile<TILE_SUB, T64>(x_half_tile(s, t), qm, qn, s.X + (size_t)t * NB * s.ldx + (size_t)k * NB, s.ldx,
                               s.W + (size_t)j * NB * s.ldw + (size_t)k * NB, s.ldw,
                               s.X + (size_t)t * NB * s.ldx + (size_t)j * NB, s.ldx, NB, smem);
}

// ---- 128x128 diagonal block: Cholesky in shared memory + explicit triangular inverse -------------
// One CTA (256 threads) per system.  info follows LAPACK dpotrf: 0, or 1 + index of the first
// non-positive (or NaN) pivot; the block is still completed with that pivot replaced by 1 so that
// everything stays finite -- the host checks info and takes the repair branch (lakernel.py:262-279).
//
// Right-looking over 8-column panels on the AUGMENTED block [A_kk ; I]: the 128 rows of the identity are
// carried along as extra right-hand-side rows, so that Z = I L^-T (upper triangular) -- the inverse the
// panel GEMMs need -- is complete when the factorisation is, with no separate serial substitution phase.
// Z[r][c] (c >= r) is parked at S[r][c+1], the unused strict upper triangle of the 128x129 array holding A/L.
// Per panel: every thread factors the 8x8 diagonal block redundantly in registers (rsqrt, no barrier);
// threads 0..127 forward-substitute the 8 panel entries of the A rows below the panel, threads 128..255
// those of the Z rows that are non-zero there (rows <= j0+7); then all threads apply the rank-8 update to
// the trailing A block (lower triangle) and to the trailing columns of those Z rows in 4x4 micro-tiles.
// Two barriers per panel.
//
// Shared-memory layout: row r starts at r * 129 + r / 4 doubles.  The trailing update works on 4 x 4 micro-tiles and the
// lanes of a warp take micro-tiles that are 4 ROWS apart (column-major enumeration), so the lane-dependent part of
// every address is a multiple of 4 rows = 517 doubles: odd, i.e. the 16 lanes of a 64-bit wavefront fall into 16
// different bank pairs.  (With a plain stride of 129 and lanes 4 columns or 4 rows apart, ncu counted 340 k bank
// conflicts in 601 k shared wavefronts per launch: profiles/ncu_r02_k_potrf_diag.txt.)
constexpr int PD = NB + 1;
__device__ __forceinline__ int prow(int r) { return r * PD + (r >> 2); }
constexpr size_t POTRF_SMEM = ((size_t)NB * PD + NB / 4) * sizeof(double);

__global__ void __launch_bounds__(256, 1) k_potrf_diag(SolveBatch bt, int k) {
    extern __shared__ __align__(16) double S[];  // [128][129]
    const SolveSys& s = bt.s[blockIdx.x];
    const int nb = s.npad / NB;
    if (k >= nb) return;
    const int tid = threadIdx.x;
    double* Wkk = s.W + (size_t)k * NB * s.ldw + (size_t)k * NB;
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e >> 7, c = e & 127;
        if (c <= r) S[prow(r) + c] = Wkk[(size_t)r * s.ldw + c];
        if (c >= r) S[prow(r) + c + 1] = (c == r) ? 1.0 : 0.0;  // Z = I
    }
    __syncthreads();
    int bad = 0;
    for (int j0 = 0; j0 < NB; j0 += 8) {
        double D[8][8], rinv[8];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b <= a; b++) D[a][b] = S[prow(j0 + a) + j0 + b];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            double d = D[c][c];
            if (!(d > 0.0)) {
                if (!bad) bad = k * NB + j0 + c + 1;
                d = 1.0;
            }
            rinv[c] = rsqrt(d);
            D[c][c] = d * rinv[c];
#pragma unroll
            for (int a = c + 1; a < 8; a++) D[a][c] *= rinv[c];
#pragma unroll
            for (int a = c + 1; a < 8; a++)
#pragma unroll
                for (int b = c + 1; b <= a; b++) D[a][b] -= D[a][c] * D[b][c];
        }
        {
            // row (r of A for tid < 128, r of Z otherwise): x[c] = (x[c] - sum_{q<c} x[q] L[c][q]) / L[c][c]
            const bool isz = tid >= NB;
            const int r = isz ? tid - NB : tid;
            double* rowp = S + prow(r) + j0 + (isz ? 1 : 0);
            if (isz ? (r < j0 + 8) : (r >= j0 + 8)) {
                double row[8];
#pragma unroll
                for (int c = 0; c < 8; c++) row[c] = (!isz || j0 + c >= r) ? rowp[c] : 0.0;  // Z is upper triangular
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    double v = row[c];
#pragma unroll
                    for (int q = 0; q < c; q++) v -= row[q] * D[c][q];
                    row[c] = v * rinv[c];
                }
                if (isz) {
                    // entries left of the diagonal of Z stay zero by construction; only c >= r is storage of Z
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        if (j0 + c >= r) rowp[c] = row[c];
                } else {
#pragma unroll
                    for (int c = 0; c < 8; c++) rowp[c] = row[c];
                }
            }
        }
        __syncthreads();
        if (tid == 255) {
            // the factored 8x8 block goes back only now: before the barrier slower threads may still be loading
            // the unfactored block.  One thread, static register indices (a per-thread row select would push D
            // into local memory); nothing reads these entries again before the final write-out.
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b <= a; b++) S[prow(j0 + a) + j0 + b] = D[a][b];
        }
        const int T0 = j0 + 8;
        const int nt4 = (NB - T0) >> 2;
        const int cntA = nt4 * (nt4 + 1) / 2;
        const int cnt = cntA + (T0 >> 2) * nt4;
        for (int t = tid; t < cnt; t += 256) {
            const double *pa, *pb;
            double* pc;
            int zrow = -1;  // first Z row of the tile (entries left of Z's diagonal are storage of A/L: read as 0)
            bool diag = false;  // diagonal A tile: only its lower triangle is A's storage (the rest belongs to Z)
            if (t < cntA) {
                // column-major walk of the lower triangle of micro-tiles (consecutive threads: consecutive tile ROWS of
                // one tile column), written as the row-major walk of the mirrored triangle from the end
                const int tr = cntA - 1 - t;
                int mi = (int)((sqrtf(8.0f * (float)tr + 1.0f) - 1.0f) * 0.5f);
                while (mi * (mi + 1) / 2 > tr) mi--;
                while ((mi + 1) * (mi + 2) / 2 <= tr) mi++;
                const int mj = tr - mi * (mi + 1) / 2;
                const int bj = nt4 - 1 - mi, bi = nt4 - 1 - mj;
                diag = bi == bj;
                pa = S + prow(T0 + 4 * bi) + j0;
                pb = S + prow(T0 + 4 * bj) + j0;
                pc = S + prow(T0 + 4 * bi) + T0 + 4 * bj;
            } else {
                const int u = t - cntA;
                const int nz4 = T0 >> 2;
                const int bj = u / nz4, zi = u - bj * nz4;  // consecutive threads: consecutive Z row groups
                zrow = 4 * zi;
                pa = S + prow(4 * zi) + j0 + 1;  // Z rows 4 zi .. 4 zi + 3, panel columns
                pb = S + prow(T0 + 4 * bj) + j0;
                pc = S + prow(4 * zi) + T0 + 4 * bj + 1;
            }
            double c4[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) c4[i][j] = pc[i * PD + j];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                double av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    av[i] = (zrow < 0 || j0 + q >= zrow + i) ? pa[i * PD + q] : 0.0;
                    bv[i] = pb[i * PD + q];
                }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) c4[i][j] -= av[i] * bv[j];
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (!diag || j <= i) pc[i * PD + j] = c4[i][j];
        }
        __syncthreads();
    }
    if (tid == 0 && bad && *s.info == 0) *s.info = bad;
    double* Di = s.Dinv + (size_t)k * NB * NB;
    double* Dt = s.Dinv + (size_t)(nb + k) * NB * NB;
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e >> 7, c = e & 127;
        if (c <= r) Wkk[(size_t)r * s.ldw + c] = S[prow(r) + c];  // L, lower triangle of the diagonal block
        Dt[e] = (r <= c) ? S[prow(r) + c + 1] : 0.0;              // inv(L)^T[r][c] = Z[r][c]
        Di[e] = (c <= r) ? S[prow(c) + r + 1] : 0.0;              // inv(L)[r][c] = Z[c][r]
    }
}

// W <- A (n x n, lda) + running sum of diagonal increments, identity in the padding (lakernel.py:295-299, 356)
struct DiagIncs {
    double v[16];
    int n;
};
__global__ void k_pad_system(double* __restrict__ W, int ldw, int n, int npad, const double* __restrict__ A, int lda,
                             DiagIncs incs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= npad) return;
    double v;
    if (r < n && c < n) {
        v = A[(size_t)r * lda + c];
        if (r == c)
            for (int q = 0; q < incs.n; q++) v += incs.v[q];
    } else {
        v = (r == c) ? 1.0 : 0.0;
    }
    W[(size_t)r * ldw + c] = v;
}

__global__ void k_transpose(const double* __restrict__ A, int lda, double* __restrict__ At, int ldat, int rows,
                            int cols) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? A[(size_t)r * lda + c] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) At[(size_t)c * ldat + r] = tile[threadIdx.x][i];
    }
}

bool g_attr_done = false;
bool g_tile64 = true;
bool g_ozaki = true;  // B200_OZAKI=0: never use the sliced INT8 path, even when the systems carry a workspace
int g_sp = SP_DEFAULT;
int gemm_attrs() {
    if (g_attr_done) return 0;
    B200_CUDA(cudaFuncSetAttribute(k_gemm_nt<TILE_ASSIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    B200_CUDA(cudaFuncSetAttribute(k_gemm_nt<TILE_SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    B200_CUDA(cudaFuncSetAttribute(k_gemm_nt<TILE_ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
#define B200_ATTR2(kern)                                                                                          \
    B200_CUDA(cudaFuncSetAttribute(kern<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));  \
    B200_CUDA(cudaFuncSetAttribute(kern<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM2_SMEM));
    B200_ATTR2(k_chol_panel)
    B200_ATTR2(k_back_diag)
    B200_ATTR2(k_chol_update)
    B200_ATTR2(k_chol_super_update)
    B200_ATTR2(k_back_super_update)
    B200_ATTR2(k_back_update)
#undef B200_ATTR2
    B200_CUDA(cudaFuncSetAttribute(k_gemm64_nt<TILE_ASSIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM2_SMEM));
    B200_CUDA(cudaFuncSetAttribute(k_gemm64_nt<TILE_SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM2_SMEM));
    B200_CUDA(cudaFuncSetAttribute(k_gemm64_nt<TILE_ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM2_SMEM));
    B200_CUDA(cudaFuncSetAttribute(k_gemm64_nt_batch<TILE_ASSIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM2_SMEM));
    B200_CUDA(cudaFuncSetAttribute(k_gemm64_nt_batch<TILE_SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM2_SMEM));
    B200_CUDA(cudaFuncSetAttribute(k_gemm64_nt_batch<TILE_ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM2_SMEM));
    {
        const char* e = getenv("B200_TILE64");  // 0: the 128x128 one-CTA-per-SM tile everywhere (A/B comparisons)
        g_tile64 = !(e && e[0] == '0');
        const char* oz = getenv("B200_OZAKI");
        g_ozaki = !(oz && oz[0] == '0');
        const char* sp = getenv("B200_SP");
        if (sp && atoi(sp) >= 1 && atoi(sp) <= 64) g_sp = atoi(sp);
    }
    B200_CUDA(cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM));
    g_attr_done = true;
    return 0;
}

}  // namespace

// ---- sliced INT8 path: workspace of one system ---------------------------------------------------
// [scaleL npad][scaleU npad][scaleX nsp x mpad] doubles, then (1024-aligned) the digit planes of W and of X.
struct OzWork {
    double *scaleL, *scaleU, *scaleX;
    int8_t *SLW, *SLX;
};
static size_t oz_scale_bytes(int npad, int mpad) {
    const int nsp = npad / NB;  // (one chunk per block column covers every super-panel width)
    return (((size_t)(2 * npad + (size_t)nsp * mpad) * sizeof(double)) + 1023) & ~(size_t)1023;
}
size_t chol_work_bytes(int npad, int mpad) {
    return oz_scale_bytes(npad, mpad) + (size_t)OZ_NS * npad * ((size_t)npad + mpad) + 1024;
}
static OzWork oz_carve(const SolveSys& s) {
    OzWork w;
    uint8_t* b = reinterpret_cast<uint8_t*>(((uintptr_t)s.work + 1023) & ~(uintptr_t)1023);
    w.scaleL = reinterpret_cast<double*>(b);
    w.scaleU = w.scaleL + s.npad;
    w.scaleX = w.scaleU + s.npad;
    w.SLW = reinterpret_cast<int8_t*>(b + oz_scale_bytes(s.npad, s.mpad));
    w.SLX = w.SLW + (size_t)OZ_NS * s.npad * s.npad;
    return w;
}

// ---- launchers ---------------------------------------------------------------------------------
int launch_chol_solve(const SolveSys* h_sys, int nsys, int do_factor, int do_solve, cudaStream_t st) {
    if (nsys <= 0) return 0;
    B200_REQUIRE(nsys <= MAXB, "at most MAXB systems per batched call");
    if (int rc = gemm_attrs()) return rc;
    SolveBatch bt;
    int nbmax = 0, mbmax = 0;
    for (int i = 0; i < nsys; i++) {
        bt.s[i] = h_sys[i];
        if (!do_solve) bt.s[i].mpad = 0;
        const SolveSys& s = bt.s[i];
        B200_REQUIRE(s.npad > 0 && s.npad % NB == 0 && s.mpad % NB == 0 && s.ldw % 2 == 0 && s.ldx % 2 == 0 &&
                         s.ldw >= s.npad && (s.mpad == 0 || s.ldx >= s.npad),
                     "padded sizes must be multiples of 128");
        nbmax = s.npad / NB > nbmax ? s.npad / NB : nbmax;
        mbmax = s.mpad / NB > mbmax ? s.mpad / NB : mbmax;
    }
    for (int i = nsys; i < MAXB; i++) bt.s[i] = bt.s[0];
    // Sliced INT8 path (ozaki.cu): every system brings a workspace and the batch has more than one super-panel.
    // Long-K products then run on tcgen05 from digit planes of the finished panels:
    //   W part  left-looking as before (rows of L share one a-priori scale sqrt(W_ii), so the INT32 accumulators run
    //           over the whole K range);
    //   X parts RIGHT-looking per super-panel chunk (the rows of Z and of Ti have no a-priori bound: each finished
    //           chunk is scaled by its exact row maxima and applied to all remaining columns at once).
    bool oz = g_ozaki && g_tile64 && do_factor && nbmax > g_sp;
    for (int i = 0; i < nsys && oz; i++)
        oz = bt.s[i].work != nullptr && bt.s[i].work_bytes >= chol_work_bytes(bt.s[i].npad, bt.s[i].mpad) &&
             bt.s[i].npad <= 60000;
    OzWork ow[MAXB];
    CUtensorMap mapWA[MAXB], mapWB[MAXB], mapXA[MAXB];
    if (oz) {
        for (int i = 0; i < nsys; i++) {
            const SolveSys& s = bt.s[i];
            ow[i] = oz_carve(s);
            if (int rc = oz_make_map(&mapWA[i], ow[i].SLW, s.npad, s.npad / OZ_BK, OZ_BM)) return rc;
            if (int rc = oz_make_map(&mapWB[i], ow[i].SLW, s.npad, s.npad / OZ_BK, OZ_BN)) return rc;
            if (s.mpad > 0)
                if (int rc = oz_make_map(&mapXA[i], ow[i].SLX, s.mpad, s.npad / OZ_BK, OZ_BM)) return rc;
        }
    }
    const int KB = NB / OZ_BK;  // K chunks of the digit planes per block column
    // slice rows [r0, r1) x block columns [b0, b1) of W (which: 0 = L part with the a-priori scales, 1 = L^T part with
    // exact row maxima) and / or all rows of X (chunk index xc >= 0) for every system; one launch each for scales / digits
    auto oz_slice = [&](int which, int wr0_blk, int wr1_blk, int wb0, int wb1, bool from_end, int xc, int xb0, int xb1) -> int {
        OzSliceBatch sb;
        int nj = 0, rows_max = 0, nkb_max = 0;
        OzSliceBatch mx;  // jobs that need an exact-maximum scale pass first
        int nmx = 0, mx_rows = 0;
        for (int q = 0; q < nsys; q++) {
            const SolveSys& s = bt.s[q];
            const int nb = s.npad / NB;
            if (which >= 0) {
                int r0 = wr0_blk, r1 = wr1_blk, b0 = wb0, b1 = wb1;
                r1 = r1 < nb ? r1 : nb;
                b1 = b1 < nb ? b1 : nb;
                if (r1 > r0 && b1 > b0) {
                    OzSliceSys j{s.W, s.ldw, r0 * NB, (r1 - r0) * NB, b0 * KB, (b1 - b0) * KB,
                                 which == 0 ? ow[q].scaleL : ow[q].scaleU, ow[q].SLW, s.npad, which == 1,
                                 which == 1 ? ow[q].scaleL : nullptr, which == 1 ? 2 : 0};
                    sb.s[nj++] = j;
                    rows_max = j.nrows > rows_max ? j.nrows : rows_max;
                    nkb_max = j.nkb > nkb_max ? j.nkb : nkb_max;
                    if (which == 1) {
                        mx.s[nmx++] = j;
                        mx_rows = j.nrows > mx_rows ? j.nrows : mx_rows;
                    }
                }
            }
            if (xc >= 0 && s.mpad > 0) {
                int b0 = xb0, b1 = xb1;
                if (from_end) {  // block columns counted from the end of this system
                    const int hi = nb - xb0, lo = nb - xb1 > 0 ? nb - xb1 : 0;
                    b0 = lo;
                    b1 = hi;
                }
                b1 = b1 < nb ? b1 : nb;
                if (b1 > b0) {
                    // (backward solve: the columns of Ti are balanced with the row scales of L, see OzSliceSys)
                    OzSliceSys j{s.X, s.ldx, 0, s.mpad, b0 * KB, (b1 - b0) * KB, ow[q].scaleX + (size_t)xc * s.mpad,
                                 ow[q].SLX, s.mpad, 0, from_end ? ow[q].scaleL : nullptr, from_end ? 1 : 0};
                    // (the scale array is indexed by absolute row; chunk xc has its own array)
                    sb.s[nj++] = j;
                    mx.s[nmx++] = j;
                    rows_max = j.nrows > rows_max ? j.nrows : rows_max;
                    mx_rows = j.nrows > mx_rows ? j.nrows : mx_rows;
                    nkb_max = j.nkb > nkb_max ? j.nkb : nkb_max;
                }
            }
        }
        if (nj == 0) return 0;
        prof_begin(PROF_OZ_SLICE, st);
        if (nmx > 0)
            if (int rc = oz_launch_scale_max(mx, nmx, mx_rows, st)) return rc;
        const int rc = oz_launch_slice(sb, nj, rows_max, nkb_max, st);
        double bytes = 0;
        for (int j = 0; j < nj; j++) bytes += 16.0 * sb.s[j].nrows * (double)sb.s[j].nkb * OZ_BK;
        prof_end(bytes, st);
        return rc;
    };
    // algorithmic flop counts of the launches (profiling only): tiles actually computed x 2*128^2*K
    // right-hand-side row tiles of system q, a half tile (x_half_tile) counting as half the work
    auto mb_eff = [&](int q) {
        const int mb = bt.s[q].mpad / NB;
        const bool half = mb > 0 && bt.s[q].mrows > 0 && bt.s[q].mrows - (mb - 1) * NB <= NB / 2;
        return mb - (half ? 0.5 : 0.0);
    };
    auto tiles_fwd = [&](int c0, int c1, int rows_from) {  // W tiles (j <= i) + X tiles over all systems
        double t = 0;
        for (int q = 0; q < nsys; q++) {
            const int nb = bt.s[q].npad / NB;
            const double mb = mb_eff(q);
            for (int j = c0; j < c1 && j < nb; j++) {
                const int ifrom = rows_from > j ? rows_from : j;
                t += (nb - ifrom > 0 ? nb - ifrom : 0) + mb;
                if (g_tile64 && ifrom == j && nb > j) t -= 0.25;  // diagonal tile: three of its four quadrants
            }
        }
        return t;
    };
    const double tile_flops = 2.0 * NB * NB;
    const double tri = g_tile64 ? 0.75 : 1.0;  // the strip kernels skip the zero quarter of the triangular inverse
    // Right-looking update of the right-hand sides with one finished K range of Z (forward) or Ti (backward):
    //   forward : X[:, j] -= Z[:, ka:kb] L[j, ka:kb]^T      for block columns ja <= j < jb   (B operand: rows of L)
    //   backward: X[:, j] -= Ti[:, ka:kb] L[ka:kb, j]        for block columns ja <= j < jb   (B operand: rows of L^T)
    // Ranges are absolute block columns (forward) or counted from the END of each system (backward: [nb - kb, nb - ka),
    // [nb - jb, nb - ja)); xc selects the scale array the K range was sliced with.
    auto oz_x_update = [&](bool backward, int ka, int kb, int ja, int jb, int xc) -> int {
        OzBatch gb;
        int mt = 0, nt = 0;
        double work = 0;
        for (int q = 0; q < nsys; q++) {
            const SolveSys& s = bt.s[q];
            const int nb = s.npad / NB;
            int k0 = ka, k1 = kb < nb ? kb : nb, j0 = ja, j1 = jb < nb ? jb : nb;
            if (backward) {
                k0 = nb - kb > 0 ? nb - kb : 0;
                k1 = nb - ka;
                j0 = nb - jb > 0 ? nb - jb : 0;
                j1 = nb - ja;
            }
            const bool any = s.mpad > 0 && k1 > k0 && j1 > j0;
            OzSys& g = gb.s[q];
            g.mapA = mapXA[q];
            g.mapB = mapWB[q];
            g.scaleA = ow[q].scaleX + (size_t)xc * s.mpad;
            g.scaleB = backward ? ow[q].scaleU : ow[q].scaleL;
            g.C = s.X + (size_t)(any ? j0 : 0) * NB;
            g.ldc = s.ldx;
            g.m_tiles = any ? s.mpad / NB : 0;
            g.n_tiles = any ? (j1 - j0) * (NB / OZ_BN) : 0;
            g.rowA0 = 0;
            g.rowB0 = (any ? j0 : 0) * NB;
            g.rowC0 = g.colC0 = 0;
            g.kb0 = (any ? k0 : 0) * KB;
            g.kb1 = (any ? k1 : 0) * KB;
            g.tri = 0;
            mt = g.m_tiles > mt ? g.m_tiles : mt;
            nt = g.n_tiles > nt ? g.n_tiles : nt;
            if (any) work += (double)(s.mpad / NB) * (j1 - j0) * (k1 - k0);
        }
        for (int q = nsys; q < MAXB; q++) gb.s[q] = gb.s[0];
        if (mt == 0 || nt == 0) return 0;
        prof_begin(PROF_OZ_GEMM, st);
        const int rc = oz_launch_gemm(gb, nsys, mt, nt, st);
        prof_end(work * tile_flops * NB, st);
        return rc;
    };
    if (do_factor) {
        const int SP = g_sp;
        if (oz) {  // a-priori row scales of L from the diagonal of W, before it is overwritten
            OzSliceBatch sb;
            int rows_max = 0;
            for (int q = 0; q < nsys; q++) {
                sb.s[q] = OzSliceSys{bt.s[q].W, bt.s[q].ldw, 0, bt.s[q].npad, 0, 0, ow[q].scaleL, ow[q].SLW, bt.s[q].npad, 0};
                rows_max = bt.s[q].npad > rows_max ? bt.s[q].npad : rows_max;
            }
            if (int rc = oz_launch_scale_diag(sb, nsys, rows_max, st)) return rc;
        }
        for (int c0 = 0; c0 < nbmax; c0 += SP) {
            const int c1 = c0 + SP < nbmax ? c0 + SP : nbmax;
            if (c0 > 0 && oz) {
                // W[i][j] -= L[i][0:c0] L[j][0:c0]^T  (c0 <= j < c1, j <= i) on the INT8 tensor cores
                OzBatch gb;
                int mt = 0, nt = 0;
                double tiles = 0;
                for (int q = 0; q < nsys; q++) {
                    const SolveSys& s = bt.s[q];
                    const int nb = s.npad / NB;
                    OzSys& g = gb.s[q];
                    g.mapA = mapWA[q];
                    g.mapB = mapWB[q];
                    g.scaleA = g.scaleB = ow[q].scaleL;
                    g.C = s.W + (size_t)c0 * NB * s.ldw + (size_t)c0 * NB;
                    g.ldc = s.ldw;
                    const int cc1 = c1 < nb ? c1 : nb;
                    g.m_tiles = nb > c0 ? nb - c0 : 0;
                    g.n_tiles = cc1 > c0 ? (cc1 - c0) * (NB / OZ_BN) : 0;
                    g.rowA0 = g.rowB0 = g.rowC0 = g.colC0 = c0 * NB;
                    g.kb0 = 0;
                    g.kb1 = c0 * KB;
                    g.tri = 1;
                    mt = g.m_tiles > mt ? g.m_tiles : mt;
                    nt = g.n_tiles > nt ? g.n_tiles : nt;
                    for (int j = c0; j < cc1; j++) tiles += (nb - j) - 0.25;
                }
                for (int q = nsys; q < MAXB; q++) gb.s[q] = gb.s[0];
                prof_begin(PROF_OZ_GEMM, st);
                if (int rc = oz_launch_gemm(gb, nsys, mt, nt, st)) return rc;
                prof_end(tiles * tile_flops * c0 * NB, st);
            } else if (c0 > 0) {
                prof_begin(PROF_CHOL_SUPER, st);
                if (g_tile64)
                    k_chol_super_update<true><<<dim3(2 * (c1 - c0), 2 * (nbmax - c0 + mbmax), nsys), GT2, GEMM2_SMEM, st>>>(bt, c0, c1);
                else
                    k_chol_super_update<false><<<dim3(c1 - c0, nbmax - c0 + mbmax, nsys), GT, GEMM_SMEM, st>>>(bt, c0, c1);
                prof_end(tiles_fwd(c0, c1, c0) * tile_flops * c0 * NB, st);
                B200_LAUNCHED(1);
            }
            for (int k = c0; k < c1; k++) {
                prof_begin(PROF_POTRF_DIAG, st);
                k_potrf_diag<<<nsys, 256, POTRF_SMEM, st>>>(bt, k);
                prof_end(nsys * (NB * (double)NB * NB), st);
                B200_LAUNCHED(1);
                const int nrow = nbmax - 1 - k;
                if (nrow + mbmax > 0) {
                    prof_begin(PROF_CHOL_PANEL, st);
                    if (g_tile64)
                        k_chol_panel<true><<<dim3(2 * (nrow + mbmax), nsys), GT2, GEMM2_SMEM, st>>>(bt, k);
                    else
                        k_chol_panel<false><<<dim3(nrow + mbmax, nsys), GT, GEMM_SMEM, st>>>(bt, k);
                    prof_end(tiles_fwd(k, k + 1, k + 1) * tile_flops * NB * tri, st);
                    B200_LAUNCHED(1);
                }
                if (c1 - 1 - k > 0) {
                    prof_begin(PROF_CHOL_INNER, st);
                    if (g_tile64)
                        k_chol_update<true><<<dim3(2 * (c1 - 1 - k), 2 * (nrow + mbmax), nsys), GT2, GEMM2_SMEM, st>>>(bt, k, c1);
                    else
                        k_chol_update<false><<<dim3(c1 - 1 - k, nrow + mbmax, nsys), GT, GEMM_SMEM, st>>>(bt, k, c1);
                    prof_end(tiles_fwd(k + 1, c1, k + 1) * tile_flops * NB, st);
                    B200_LAUNCHED(1);
                }
            }
            if (oz && c1 < nbmax) {
                // the finished super-panel becomes digit planes: L rows below it (a-priori scales), its own rows of
                // L^T to the right of it (exact row maxima; read by the backward solve), and the chunk of Z
                const int xc = c0 / SP;
                if (int rc = oz_slice(0, c1, nbmax, c0, c1, false, -1, 0, 0)) return rc;
                if (mbmax > 0) {
                    if (int rc = oz_slice(1, c0, c1, c0 + 1, nbmax, false, -1, 0, 0)) return rc;
                    // Chunks of Z are applied in PAIRS where possible (K = 2 SP blocks per launch amortises the drain of
                    // the accumulators): the first chunk of a pair only brings the NEXT super-panel up to date; when the
                    // second is finished both are sliced again with their joint row maxima and applied to everything
                    // to the right of it in one launch.
                    if (xc % 2 == 0) {
                        if (int rc = oz_slice(-1, 0, 0, 0, 0, false, xc, c0, c1)) return rc;
                        if (int rc = oz_x_update(false, c0, c1, c1, c1 + SP, xc)) return rc;
                    } else {
                        if (int rc = oz_slice(-1, 0, 0, 0, 0, false, xc, c0 - SP, c1)) return rc;
                        if (int rc = oz_x_update(false, c0 - SP, c1, c1, nbmax, xc)) return rc;
                    }
                }
            }
        }
        B200_CUDA(cudaGetLastError());
    }
    if (do_solve == 1 && mbmax > 0) {  // (do_solve == 2: forward substitution only, X <- X L^-T)
        double mbsum = 0;
        for (int q = 0; q < nsys; q++) mbsum += mb_eff(q);
        const int SP = g_sp;
        for (int e0 = 0; e0 < nbmax; e0 += SP) {
            const int e1 = e0 + SP < nbmax ? e0 + SP : nbmax;
            if (e0 > 0 && !oz) {
                prof_begin(PROF_BACK_SUPER, st);
                if (g_tile64)
                    k_back_super_update<true><<<dim3(2 * (e1 - e0), 2 * mbmax, nsys), GT2, GEMM2_SMEM, st>>>(bt, e0, e1);
                else
                    k_back_super_update<false><<<dim3(e1 - e0, mbmax, nsys), GT, GEMM_SMEM, st>>>(bt, e0, e1);
                double t = 0;
                for (int q = 0; q < nsys; q++) {
                    const int nb = bt.s[q].npad / NB;
                    const int hi = nb - e0, lo = nb - e1 > 0 ? nb - e1 : 0;
                    if (hi > lo) t += (double)(hi - lo) * mb_eff(q);
                }
                prof_end(t * tile_flops * e0 * NB, st);
                B200_LAUNCHED(1);
            }
            for (int kk = e0; kk < e1; kk++) {
                prof_begin(PROF_BACK_DIAG, st);
                if (g_tile64)
                    k_back_diag<true><<<dim3(2 * mbmax, nsys), GT2, GEMM2_SMEM, st>>>(bt, kk);
                else
                    k_back_diag<false><<<dim3(mbmax, nsys), GT, GEMM_SMEM, st>>>(bt, kk);
                prof_end(mbsum * tile_flops * NB * tri, st);
                B200_LAUNCHED(1);
                if (e1 - 1 - kk > 0) {
                    prof_begin(PROF_BACK_INNER, st);
                    if (g_tile64)
                        k_back_update<true><<<dim3(2 * (e1 - 1 - kk), 2 * mbmax, nsys), GT2, GEMM2_SMEM, st>>>(bt, kk, e1);
                    else
                        k_back_update<false><<<dim3(e1 - 1 - kk, mbmax, nsys), GT, GEMM_SMEM, st>>>(bt, kk, e1);
                    prof_end(mbsum * (e1 - 1 - kk) * tile_flops * NB, st);
                    B200_LAUNCHED(1);
                }
            }
            if (oz && e1 < nbmax) {
                // the finished chunk of Ti (block columns [nb - e1, nb - e0) of each system) is sliced and applied to all
                // columns to its left:  X[t][j] -= Ti[t][lo:hi] L[lo:hi][j]  (j < lo; L^T rows live in W's upper triangle)
                const int xc = e0 / SP;
                if (xc % 2 == 0) {  // (pairs of chunks, as in the forward solve)
                    if (int rc = oz_slice(-1, 0, 0, 0, 0, true, xc, e0, e1)) return rc;
                    if (int rc = oz_x_update(true, e0, e1, e1, e1 + SP, xc)) return rc;
                } else {
                    if (int rc = oz_slice(-1, 0, 0, 0, 0, true, xc, e0 - SP, e1)) return rc;
                    if (int rc = oz_x_update(true, e0 - SP, e1, e1, nbmax, xc)) return rc;
                }
            }
        }
        B200_CUDA(cudaGetLastError());
    }
    return 0;
}

// C (M x N) = [C +/-] A (M x K) * B (N x K)^T ; M, N multiples of 128, K even.
// accumulate: 0 assign, 1 add, -1 subtract.
int launch_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                   int accumulate, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    B200_REQUIRE(M % NB == 0 && N % NB == 0 && K % 2 == 0 && K > 0 && lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0,
                 "gemm_nt wants M,N multiples of 128, even K and leading dimensions");
    if (int rc = gemm_attrs()) return rc;
    dim3 grid(N / NB, M / NB);
    prof_begin(PROF_GEMM, st);
    if (g_tile64) {
        dim3 grid2(N / TB, M / TB);
        if (accumulate == 0)
            k_gemm64_nt<TILE_ASSIGN><<<grid2, GT2, GEMM2_SMEM, st>>>(A, lda, B, ldb, C, ldc, K);
        else if (accumulate > 0)
            k_gemm64_nt<TILE_ADD><<<grid2, GT2, GEMM2_SMEM, st>>>(A, lda, B, ldb, C, ldc, K);
        else
            k_gemm64_nt<TILE_SUB><<<grid2, GT2, GEMM2_SMEM, st>>>(A, lda, B, ldb, C, ldc, K);
    } else if (accumulate == 0)
        k_gemm_nt<TILE_ASSIGN><<<grid, GT, GEMM_SMEM, st>>>(A, lda, B, ldb, C, ldc, K);
    else if (accumulate > 0)
        k_gemm_nt<TILE_ADD><<<grid, GT, GEMM_SMEM, st>>>(A, lda, B, ldb, C, ldc, K);
    else
        k_gemm_nt<TILE_SUB><<<grid, GT, GEMM_SMEM, st>>>(A, lda, B, ldb, C, ldc, K);
    prof_end(2.0 * M * (double)N * K, st);
    B200_LAUNCH_CHECK();
    return 0;
}

// Several products C_q (M_q x N_q) = [C_q +/-] A_q B_q^T in one launch (M, N multiples of 64, K even, at most MAXB).
int launch_gemm_nt_batch(const GemmProb* probs, int nprob, int accumulate, cudaStream_t st) {
    if (nprob <= 0) return 0;
    B200_REQUIRE(nprob <= MAXB, "at most MAXB products per batched GEMM");
    if (int rc = gemm_attrs()) return rc;
    GemmBatch gb;
    int mmax = 0, nmax = 0;
    double flops = 0;
    for (int q = 0; q < nprob; q++) {
        const GemmProb& g = probs[q];
        B200_REQUIRE(g.M % TB == 0 && g.N % TB == 0 && g.K % 2 == 0 && g.K > 0 && g.lda % 2 == 0 && g.ldb % 2 == 0 &&
                         g.ldc % 2 == 0,
                     "gemm_nt_batch wants M, N multiples of 64, even K and leading dimensions");
        gb.p[q] = g;
        mmax = g.M > mmax ? g.M : mmax;
        nmax = g.N > nmax ? g.N : nmax;
        flops += 2.0 * g.M * (double)g.N * g.K;
    }
    for (int q = nprob; q < MAXB; q++) gb.p[q] = gb.p[0];
    if (mmax == 0 || nmax == 0) return 0;
    const dim3 grid(nmax / TB, mmax / TB, nprob);
    prof_begin(PROF_GEMM, st);
    if (accumulate == 0)
        k_gemm64_nt_batch<TILE_ASSIGN><<<grid, GT2, GEMM2_SMEM, st>>>(gb);
    else if (accumulate > 0)
        k_gemm64_nt_batch<TILE_ADD><<<grid, GT2, GEMM2_SMEM, st>>>(gb);
    else
        k_gemm64_nt_batch<TILE_SUB><<<grid, GT2, GEMM2_SMEM, st>>>(gb);
    prof_end(flops, st);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_pad_system(double* W, int ldw, int n, int npad, const double* A, int lda, const double* incs, int ninc,
                      cudaStream_t st) {
    B200_REQUIRE(ninc <= 16, "at most 16 diagonal increments");
    DiagIncs di;
    di.n = ninc;
    for (int i = 0; i < 16; i++) di.v[i] = i < ninc ? incs[i] : 0.0;
    k_pad_system<<<dim3((npad + 255) / 256, npad), 256, 0, st>>>(W, ldw, n, npad, A, lda, di);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_transpose(const double* A, int lda, double* At, int ldat, int rows, int cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    k_transpose<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, st>>>(A, lda, At, ldat, rows, cols);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace b200
