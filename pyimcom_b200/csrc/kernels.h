// Internal launcher declarations shared by the .cu files and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/pyimcom_b200.h"

namespace b200 {

// the structs that cross the C ABI are defined once, in include/pyimcom_b200.h
using TableRef = ::b200_table_ref;
using SolveSys = ::b200_solve_sys;
using FinalizeArgs = ::b200_finalize_args;
using PairDesc = ::b200_pair_desc;
using AsmDesc = ::b200_asm_desc;
using PartCell = ::b200_part_cell;
using EighProblem = ::b200_eigh_problem;

constexpr int NB = B200_NB;      // Cholesky/TRSM block size == DMMA GEMM tile edge
constexpr int MAXB = B200_MAXB;  // systems per batched launch (descriptors travel as kernel parameters)

struct SolveBatch {
    SolveSys s[MAXB];
};

// one product of a batched GEMM launch (linalg.cu: launch_gemm_nt_batch)
struct GemmProb {
    const double* A;
    const double* B;
    double* C;
    int lda, ldb, ldc, M, N, K;
};
struct GemmBatch {
    GemmProb p[MAXB];
};

struct DiagNodes {
    double v[16];
};

// interp.cu
int launch_getw(double* w, double fh, cudaStream_t s);
int launch_iD5512C(const double* f, int nlayer, int ngy, int ngx, const double* x, const double* y, long nout,
                   double* out, cudaStream_t s);
int launch_iD5512C_sym(const double* f, int nlayer, int ngy, int ngx, const double* x, const double* y, long nout,
                       double* out, cudaStream_t s);
int launch_gridD5512C(const double* f, int ngy, int ngx, const double* x, const double* y, long npi, int nxo, int nyo,
                      double* out, cudaStream_t s);
int launch_layout_tables(const double* src, int ntab, int ns, int pad, int ngrid, int poly, double* dst,
                         cudaStream_t s);
int launch_cmul_conj(const double* ar, const double* ai, const double* br, const double* bi, long long per,
                     long long stride2, long long n1, double im_sign, double* gr, double* gi, cudaStream_t s);
int launch_gather_stamp(const int* idx, int n, int npad, const double* src_x, const double* src_y, const int* src_code,
                        const float* src_data, long src_ld, int n_inframe, double* px, double* py, int* pcode,
                        float* indata, int ldi, cudaStream_t s);
int launch_build_A(const double* px, const double* py, const int* pcode, int n, int npad, const double* tables,
                   const TableRef* lut, int nimg, int ncode, int ngrid, double dscale, double nc, double flat_penalty,
                   double* A, int lda, double diag_add, int poly, cudaStream_t s);
int launch_pair_blocks(const double* gx, const double* gy, const int* gimg, const PairDesc* descs,
                       const int* tile_prefix, int npair, int ntiles, const double* tables, const TableRef* lut,
                       int nimg, int ngrid, double dscale, double nc, double flat_penalty, int poly, double* pool,
                       double points, cudaStream_t s);
int launch_assemble_A(const AsmDesc& d, const int* gidx, int n, int npad, const double* pool, double* A, int lda,
                      double diag_add, cudaStream_t s);
int launch_build_B(const double* px, const double* py, const int* pcode, int n, int npad, const double* tables,
                   const long long* lut_io, int n_out, int ngrid, double dscale, double nc, int n2f, int mpad,
                   double x0out, double y0out, double* B, int ldb, size_t strideB, cudaStream_t s);

// linalg.cu
int launch_chol_solve(const SolveSys* h_sys, int nsys, int do_factor, int do_solve, cudaStream_t s);
size_t chol_work_bytes(int npad, int mpad);
int launch_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                   int accumulate, cudaStream_t s);
int launch_gemm_nt_batch(const GemmProb* probs, int nprob, int accumulate, cudaStream_t s);
int launch_pad_system(double* W, int ldw, int n, int npad, const double* A, int lda, const double* incs, int ninc,
                      cudaStream_t s);
int launch_transpose(const double* A, int lda, double* At, int ldat, int rows, int cols, cudaStream_t s);

// kappa.cu
int launch_lakernel1(const double* lam, const double* mPhalf, int ldp, int m, int n, double C, double targetleak,
                     double kCmin, double kCmax, int nbis, double* kappa, double* Sigma, double* UC, double* T, int ldt,
                     double smax, cudaStream_t s);
int launch_eigen_single(const double* lam, const double* mPhalf, int ldp, int m, int n, double C, double kappa,
                        double* Sigma, double* UC, double* T, int ldt, cudaStream_t s);
int launch_lsolve_sps(int N, double* A, double* x, const double* b, double* work, cudaStream_t s);
int launch_build_reduced_T(const double* Nflat, const double* Dflat, const double* Eflat, const double* kappa, int nv,
                           int m, double ucmin, double smax, double* out_kappa, double* out_Sigma, double* out_UC,
                           double* out_w, int* out_iv, int* out_branch, cudaStream_t s);
int launch_node_stats(const double* mB, int ldb, const double* Tpi, int ldt, size_t strideT, int nv, int m, int n,
                      const double* kappa_nodes, double Cnorm, double* Dp, double* Npq, double* Epq, double* DpC,
                      double* EpqC, const double* Epq_in, cudaStream_t s);
int launch_rowdot(const double* X, int ldx, const double* Y, int ldy, int m, int n, double* out, int ostride,
                  cudaStream_t s);

int launch_single_kappa_maps(const double* D, const double* N, const double* E, int m, double kappa, double C,
                             double* kappa_out, double* Sigma_out, double* UC_out, cudaStream_t s);
int launch_empir_T(const double* inx, const double* iny, const double* outx, const double* outy, int m, int mpad, int n,
                   int npad, double rho, double* T, int ldt, cudaStream_t s);
int launch_scale(const double* in, double scale, int m, double* out, cudaStream_t s);

// iter.cu
int launch_iter_cg(const double* AA, int lda, double diag_add, const double* mB, int ldb, int m, int n,
                   const double* inx, const double* iny, const double* outx, const double* outy, double rho_acc,
                   double rtol, int maxiter, double* Ti, int ldt, int* niter, int* nsel, cudaStream_t s);

// eigen.cu
int launch_jacobi_eigh_batch(const EighProblem* pr, int nsys, int max_sweeps, int* sweeps_done, cudaStream_t s);
int launch_jacobi_eigh(double* A, int lda, int n, double* Vt, int ldv, double* lam, int max_sweeps, int* sweeps_done,
                       cudaStream_t s);

// trieig.cu
int launch_tri_eigh_batch(const EighProblem* pr, int nsys, cudaStream_t s);
long long eigh_fallback_count();
int launch_tridiag(double* A, int lda, int n, double* d, double* e, double* tau, cudaStream_t s);

// coadd.cu
int launch_finalize(const FinalizeArgs& a, cudaStream_t s);
int launch_stamp_maps(const double* kappa, const double* Sigma, const double* UC, int m, int n2f, int fade,
                      int clamp_iter, const double* fade_w, float* kappa32, float* Sigma32, float* UC32,
                      const double* Tsum_image, int n_img, int n2, double* Tsum_stamp, double* Tsum_inpix, double* Neff,
                      cudaStream_t s);
int launch_accumulate(const void* src, int src_is_f64, int nlayer, int n2f, float* dst, int side, int y0, int x0,
                      cudaStream_t s);
int launch_accumulate_stamp(const float* outimage, int nfr, const float* UC, const float* Sigma, const float* kappa,
                            const double* Tsum_inpix, const double* Neff, const double* Tsum_stamp, int n_img, int n2f,
                            float* out_map, float* UC_map, float* Sigma_map, float* kappa_map, float* Tsum_map,
                            float* Neff_map, int side, int y0, int x0, float* T_weight, int tw_stride, cudaStream_t s);
int launch_unfade_crop(const float* in, int nlayer, int side, int fk, int recover, int pb, int pt, int pl, int pr,
                       const double* fade_w, float* out, cudaStream_t s);
int launch_compress_map(const float* in, long n, int coef, int is_unsigned, void* out, cudaStream_t s);
int launch_partition(const PartCell* cells, int ncell, const double* ox, const double* oy, const unsigned char* mask,
                     int sca, const unsigned char* use, int ns, int n2, double lower, double upper, int npixmax,
                     int* sid_tmp, unsigned* rank_tmp, int* cellmeta, unsigned* cellcnt, unsigned* cellbase,
                     unsigned* run, unsigned* pix_count, unsigned short* y_idx, unsigned short* x_idx, double* y_val,
                     double* x_val, int* err, cudaStream_t s);
int launch_extract_layers(const float* indata, int n_inframe, int sca, const unsigned short* y_idx,
                          const unsigned short* x_idx, const unsigned* pix_count, int nstamp, int npixmax, int max_count,
                          float* data, cudaStream_t s);
int launch_assemble_instamps(const double* x_val, const double* y_val, const float* data, int n_inframe, int nstamp,
                             int npixmax, int max_count, const unsigned* pix_count, const long long* dst_off, int image,
                             long long npix, double* gx, double* gy, int* gimg, float* gdata, cudaStream_t s);

}  // namespace b200
