// Host-visible descriptors of the sliced INT8 (Ozaki) GEMM path: see ozaki.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace b200 {

#ifndef B200_OZ_NS
#define B200_OZ_NS 8
#endif
constexpr int OZ_NS = B200_OZ_NS;  // radix-128 digit planes per operand (7 bits each: 8 planes = 56 bits below the row scale)
constexpr int OZ_BM = 128, OZ_BN = 64, OZ_BK = 64, OZ_STAGES = 2;

// One C -= A B^T problem of a batched launch.  Operand rows are addressed in the digit-plane arrays the tensor maps
// describe: tile (tm, tn) reads rows rowA0 + 128 tm .. of map A and rows rowB0 + 64 tn .. of map B, K chunks
// kb0 .. kb1-1 (64 columns each), and updates C[128 tm .. , 64 tn ..] (C points at the first output element).
struct alignas(64) OzSys {
    CUtensorMap mapA, mapB;
    const double* scaleA;  // per row of the A array: 2^(e - 7)
    const double* scaleB;
    double* C;
    int ldc;
    int m_tiles, n_tiles;
    int rowA0, rowB0;
    int rowC0, colC0;  // position of C[0][0] in its matrix (only for the triangular skip)
    int kb0, kb1;
    int tri;  // 1: skip tiles that lie strictly above the diagonal of the matrix
};

struct OzBatch {
    OzSys s[MAXB];
    unsigned long long* dbgbuf = nullptr;  // B200_OZ_TIMERS: per-phase %globaltimer sums of the CTAs
    int dbg = 0;  // timing experiments (B200_OZ_DBG): 1 = skip the MMAs, 2 = skip the TMA loads after the first fill
};

// One slicing job: rows row0 .. row0 + nrows - 1 of `src` (row-major, ld), K chunks kb0 .. kb0 + nkb - 1.
struct OzSliceSys {
    const double* src;
    int ld;
    int row0, nrows;
    int kb0, nkb;
    double* scale;   // [rows_total] indexed by absolute row
    int8_t* S;       // digit planes [.][NS][rows_total][64]
    int rows_total;
    int tri;         // 1: row r only owns the K chunks to the right of its own 128 x 128 diagonal block (rows of L^T)
    // Column (K index) scaling by powers of two, exact: colmode 1 slices src[r][k] * colscale[k], colmode 2 slices
    // src[r][k] / colscale[k] with the row scale fixed at 1 (the quotient is bounded by 64 by construction).  The backward
    // solve pairs X[i][k] with L[k][j]: X's columns scale like 1 / sqrt(W_kk) and L's row k like sqrt(W_kk), so the
    // balanced operands are X[i][k] s_k and L[k][j] / s_k with s_k the a-priori row scale of L.
    const double* colscale;
    int colmode;
};

struct OzSliceBatch {
    OzSliceSys s[2 * MAXB];
};

int oz_make_map(CUtensorMap* map, const int8_t* S, int rows, int nkb, int box_rows);
int oz_launch_gemm(const OzBatch& b, int nsys, int m_tiles_max, int n_tiles_max, cudaStream_t st);
int oz_launch_scale_max(const OzSliceBatch& b, int nsys, int nrows_max, cudaStream_t st);
int oz_launch_scale_diag(const OzSliceBatch& b, int nsys, int nrows_max, cudaStream_t st);
int oz_launch_slice(const OzSliceBatch& b, int nsys, int nrows_max, int nkb_max, cudaStream_t st);
size_t oz_gemm_work_bytes(int M, int N, int K);
int launch_ozaki_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                         void* work, size_t work_bytes, cudaStream_t st);

}  // namespace b200
