/*
 * pyimcom_b200.h -- C ABI of libpyimcom_b200.so: the B200 (sm_100a) implementation of PyIMCOM's
 * per-postage-stamp coaddition hot path.
 *
 * Conventions
 *   - Every entry point returns 0 on success or a non-zero code; b200_last_error() then returns a
 *     human-readable message (thread-local).  Nothing is thrown across the boundary.
 *   - Plain pointers and sizes only.  "Host seam" functions (section 1) take HOST pointers, do their own
 *     H2D/D2H copies on the default stream and return after the result is in the caller's array: they
 *     are what a ctypes/CPython binding of furry_parakeet.pyimcom_croutines would bind.
 *   - "Device" functions (b200_dev_*, sections 2-6) take DEVICE pointers plus a cudaStream_t passed as
 *     void* (NULL = legacy default stream), only enqueue work and never synchronise.
 *   - All matrices are row-major ("C order"), float64 unless stated.
 *
 * Each declaration cites the reference interface (file:line under the PyIMCOM source tree) it replaces.
 */
#ifndef PYIMCOM_B200_H
#define PYIMCOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- 0. library ------------------------------------------------------------------------------- */
const char* b200_last_error(void);
int b200_version(void);
/* number of kernels launched by this library in this process since load (bench.py's gpu_launches) */
long long b200_launch_count(void);
/* frees the grow-only device scratch used by the host-seam functions */
int b200_release_scratch(void);

/* Per-launch device timing for bench.py's roofline: b200_profile(1) starts recording CUDA events around the
 * launches of the kinds below (b200_profile(0) stops and clears); b200_profile_read synchronises the device and
 * returns the summed duration (ms), the summed algorithmic work (flops for the DMMA kinds, bytes for the HBM
 * kinds) and the number of launches of one kind. */
enum b200_prof_kind {
    B200_PROF_CHOL_SUPER = 0, B200_PROF_POTRF_DIAG, B200_PROF_CHOL_PANEL, B200_PROF_CHOL_INNER,
    B200_PROF_BACK_SUPER, B200_PROF_BACK_DIAG, B200_PROF_BACK_INNER, B200_PROF_BUILD_A, B200_PROF_BUILD_B,
    B200_PROF_FINALIZE, B200_PROF_GEMM, B200_PROF_ITER_CG, B200_PROF_LAKERNEL1, B200_PROF_EIGH, B200_PROF_ASSEMBLE_A,
    B200_PROF_NKINDS
};
int b200_profile(int on);
int b200_profile_read(int kind, double* ms, double* work, long long* count);

/* ---- 1. host seam: furry_parakeet.pyimcom_croutines / pyimcom.routine ------------------------- */
/* routine.py:125-181  iD5512C(infunc (nlayer,ngy,ngx), xpos (nout), ypos (nout), fhatout (nlayer,nout));
 * off-grid points leave fhatout untouched (routine.py:166-167). */
int b200_iD5512C(const double* infunc, int nlayer, int ngy, int ngx, const double* xpos, const double* ypos,
                 long nout, double* fhatout);
/* routine.py:184-253  same, nout a perfect square, upper triangle evaluated then mirrored. */
int b200_iD5512C_sym(const double* infunc, int nlayer, int ngy, int ngx, const double* xpos, const double* ypos,
                     long nout, double* fhatout);
/* routine.py:256-338  gridD5512C(infunc (ngy,ngx), xpos (npi,nxo), ypos (npi,nyo), fhatout (npi,nyo*nxo));
 * off-grid rows/columns produce 0 (routine.py:307-310). */
int b200_gridD5512C(const double* infunc, int ngy, int ngx, const double* xpos, const double* ypos, long npi,
                    int nxo, int nyo, double* fhatout);
/* routine.py:29-122  the ten D5512 weights for fh = frac - 1/2. */
int b200_iD5512C_getw(double* w, double fh);
/* routine.py:341-430  lakernel1(lam (n), Q (unused, as in the reference), mPhalf (m,n), C, targetleak, kCmin,
 * kCmax, nbis, kappa (m), Sigma (m), UC (m), T (m,n), smax). */
int b200_lakernel1(const double* lam, const double* mPhalf, long m, long n, double C, double targetleak,
                   double kCmin, double kCmax, int nbis, double* kappa, double* Sigma, double* UC, double* T,
                   double smax);
/* routine.py:433-484  lsolve_sps(N, A (N,N) destroyed, x (N) out, b (N)). */
int b200_lsolve_sps(int N, double* A, double* x, const double* b);
/* routine.py:487-588  build_reduced_T_wrap(Nflat (m*nv*nv), Dflat (m*nv), Eflat (m*nv*nv), kappa (nv), ucmin, smax,
 * out_kappa (m), out_Sigma (m), out_UC (m), out_w (m*nv)); out_iv/out_branch (int32, m) optional diagnostics:
 * bracket index and the 12-step branch word (bit k set = step k divided kappa). */
int b200_build_reduced_T_wrap(const double* Nflat, const double* Dflat, const double* Eflat, const double* kappa,
                              int nv, long m, double ucmin, double smax, double* out_kappa, double* out_Sigma,
                              double* out_UC, double* out_w, int32_t* out_iv, int32_t* out_branch);

/* ---- 2. device: interpolation and system-matrix assembly (stage a) ---------------------------- */
int b200_dev_iD5512C(const double* infunc, int nlayer, int ngy, int ngx, const double* xpos, const double* ypos,
                     long nout, double* fhatout, void* stream);
int b200_dev_iD5512C_sym(const double* infunc, int nlayer, int ngy, int ngx, const double* xpos, const double* ypos,
                         long nout, double* fhatout, void* stream);
int b200_dev_gridD5512C(const double* infunc, int ngy, int ngx, const double* xpos, const double* ypos, long npi,
                        int nxo, int nyo, double* fhatout, void* stream);

/* One PSF-overlap table as seen by the A-assembly kernel. */
typedef struct b200_table_ref {
    long long offset;   /* offset in doubles of the zero-padded (ngrid x ngrid) table inside the arena; <0: absent */
    int flip;           /* read the table mirrored in both axes (psfutil.py:1659-1665) */
    int pad_;
    double penalty_sub; /* flat_penalty / n_in of this PSF-group pair (psfutil.py:1484, 1706) */
} b200_table_ref;

/* PSF-overlap tables as the host holds them, src (ntab, ns, ns), into the layout the assembly kernels read: each
 * table zero-padded by pad on every side (np.pad(ovl, 6), psfutil.py:1471, 1580, 1696) to ngrid = ns + 2 pad,
 * row-major (poly == 0: ngrid*ngrid doubles per table) or polyphase of period poly (see b200_dev_build_A:
 * poly*poly*ncell*ncell doubles per table, ncell = ceil(ngrid / poly)). */
int b200_dev_layout_tables(const double* src, int ntab, int ns, int pad, int ngrid, int poly, double* dst,
                           void* stream);

/* PSF-overlap spectra (PSFOvl._build_psfovl, psfutil.py:1244-1294: rft1 * conj(rft2)): n1 spectra of `per` complex
 * entries each, stored as separate real and imaginary planes, times the conjugate of spectrum t of the second stack
 * (stride2 doubles apart; stride2 == 0 broadcasts one spectrum); im_sign = -1 returns the conjugate product
 * conj(F1) * F2 instead. */
int b200_dev_cmul_conj(const double* ar, const double* ai, const double* br, const double* bi, long long per,
                       long long stride2, long long n1, double im_sign, double* gr, double* gi, void* stream);

/* Gather the selected input pixels of one output stamp (coadd.py:969-977): out[k] = src[idx[k]] for positions,
 * codes (src_code/pcode may both be NULL) and the n_inframe float32 layers (src_data (n_inframe, src_ld) ->
 * indata (n_inframe, ldi), zero padded). */
int b200_dev_gather_stamp(const int* idx, int n, int npad, const double* src_x, const double* src_y,
                          const int* src_code, const float* src_data, long src_ld, int n_inframe, double* px,
                          double* py, int* pcode, float* indata, int ldi, void* stream);

/* A (npad x lda) for one output stamp: replaces PSFOvl._call_ii_self/_call_ii_cross (psfutil.py:1401-1495,
 * 1597-1732) and the 9+36 block scatter of OutStamp._build_system_matrices (coadd.py:1028-1069).
 * px,py (n): positions in output pixels; pcode (n): local_group*nimg + image; lut (ncode*ncode).
 * Entries with index >= n form an identity block; diag_add is added to the first n diagonal entries.
 * poly == 0: the tables referenced by lut are row-major (ngrid x ngrid).  poly == P > 0 (P = oversamp, the native
 * pixel pitch in table samples): they are stored polyphase, T'[y % P][x % P][y / P][x / P] with planes of
 * ncell x ncell doubles, ncell = ceil(ngrid / P), i.e. P*P*ncell*ncell doubles per table -- neighbouring input
 * pixels then read neighbouring doubles.  Same values, same arithmetic order, bit-identical results. */
int b200_dev_build_A(const double* px, const double* py, const int* pcode, int n, int npad, const double* tables,
                     const b200_table_ref* lut, int nimg, int ncode, int ngrid, double dscale, double nc,
                     double flat_penalty, double* A, int lda, double diag_add, int poly, void* stream);
/* InStamp-pair blocks of A (SysMatA.get_iisubmat / _compute_iisubmats, psfutil.py:1764-2092): the device form of
 * the reference's block cache.  One descriptor per FULL block: all nA pixels of InStamp a (global pixel range
 * offA..) against all nB pixels of InStamp b, a <= b in raster order, written row-major with leading dimension ld
 * at pool + out (doubles).  same != 0 (a == b): the upper triangle is interpolated and mirrored (psfutil.py:1692-1714).
 * lut selects the (nimg x nimg) table-reference block of the two PSF groups inside the lut array. */
typedef struct b200_pair_desc {
    int offA, nA, offB, nB;
    long long out;
    int ld, lut, same, pad_;
} b200_pair_desc;
/* gx, gy (npix): positions of every input pixel of the mosaic block in output pixels; gimg (npix): image index;
 * descs, tile_prefix (npair, device): descriptors and the exclusive prefix sum of their 32x32 tile counts
 * (ceil(nA/32)*ceil(nB/32)); ntiles: the total; points: number of entries evaluated (profiling only). */
int b200_dev_pair_blocks(const double* gx, const double* gy, const int* gimg, const b200_pair_desc* descs,
                         const int* tile_prefix, int npair, int ntiles, const double* tables,
                         const b200_table_ref* lut, int nimg, int ngrid, double dscale, double nc, double flat_penalty,
                         int poly, double* pool, double points, void* stream);
/* One OutStamp's A cut out of cached pair blocks (the np.ix_ scatter of coadd.py:1028-1069): the stamp's n pixels are
 * 9 consecutive segments (one per InStamp of the 3x3 neighbourhood, seg_start[0..9]); pixel k is global pixel gidx[k],
 * i.e. pixel gidx[k] - inst_off[s] of its InStamp; blk[9*s+t], ld[9*s+t] (s <= t) locate block (s, t) inside pool.
 * Identity in rows/columns n..npad-1, diag_add on the first n diagonal entries. */
typedef struct b200_asm_desc {
    long long blk[81];
    int ld[81];
    int seg_start[10];
    int inst_off[9];
    int pad_;
} b200_asm_desc;
int b200_dev_assemble_A(const b200_asm_desc* desc, const int* gidx, int n, int npad, const double* pool, double* A,
                        int lda, double diag_add, void* stream);
/* mBhalf (n_out, mpad, ldb) for one output stamp: replaces PSFOvl._call_io_cross (psfutil.py:1497-1595) and
 * coadd.py:1075-1082.  lut_io (ncode*n_out) table offsets (<0 absent); output pixel (iy,ix) sits at
 * (x0out + ix, y0out + iy) (coadd.py:879-882).  Padding rows/columns are zero-filled.  tables must be 16-byte aligned
 * and readable one double past its last table (aligned 16-byte loads). */
int b200_dev_build_B(const double* px, const double* py, const int* pcode, int n, int npad, const double* tables,
                     const long long* lut_io, int n_out, int ngrid, double dscale, double nc, int n2f, int mpad,
                     double x0out, double y0out, double* B, int ldb, size_t strideB, void* stream);

/* ---- 3. device: dense FP64 linear algebra on the DMMA tensor pipe (stage b, CholKernel) -------- */
#define B200_NB 128   /* block size of the factorisation == GEMM tile edge */
#define B200_MAXB 16  /* systems per batched call */

/* One SPD system (A + kappa I) Ti^T = mBhalf^T (lakernel.py:295-304).  Padded sizes are multiples of B200_NB. */
typedef struct b200_solve_sys {
    double* W;    /* (npad, ldw) in: A + kappa I (lower triangle read, identity padding); out: L (lower) and L^T
                     in the strictly upper block triangle */
    double* X;    /* (mpad, ldx) in: mBhalf rows; out: Ti rows */
    double* Dinv; /* (2*npad/128, 128, 128) workspace: inverses of the diagonal blocks of L and their transposes */
    int* info;    /* device int, LAPACK dpotrf convention: 0 or 1 + index of the first non-positive pivot */
    int npad, mpad, ldw, ldx;
    int mrows;    /* number of real rows of X (rows mrows..mpad-1 are zero padding and stay zero); 0 = treat all as real.
                     A last row-tile with at most 64 real rows is solved at half cost. */
    int pad_;
    void* work;        /* optional device scratch of b200_chol_work_bytes(npad, mpad) bytes, 1024-byte aligned: digit planes
                          and row scales of the finished panels.  With it (for every system of the call) the long-K panel
                          updates run on the INT8 tcgen05 tensor cores (see b200_dev_ozaki_gemm_nt); NULL: all-DMMA path. */
    size_t work_bytes;
} b200_solve_sys;

/* scipy.linalg.cholesky + cho_solve (lakernel.py:263, 276, 304, 358) for up to B200_MAXB systems at once.
 * do_solve: 0 factor only, 1 X <- X (L L^T)^-1, 2 forward substitution only (X <- X L^-T). */
int b200_dev_chol_solve(const b200_solve_sys* sys, int nsys, int do_factor, int do_solve, void* stream);
size_t b200_chol_work_bytes(int npad, int mpad);
/* W <- A (n x n) + sum(incs) on the diagonal, identity in rows/cols n..npad-1 (lakernel.py:295-299, 356). */
int b200_dev_pad_system(double* W, int ldw, int n, int npad, const double* A, int lda, const double* incs, int ninc,
                        void* stream);
/* C (M x N) = [C +/-] A (M x K) * B (N x K)^T; M,N multiples of 128, K even; accumulate 0/+1/-1. */
int b200_dev_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                     int accumulate, void* stream);
/* C (M x N) -= A (M x K) * B (N x K)^T evaluated on the INT8 tcgen05 tensor cores from error-free integer slices of
 * the float64 operands (Ozaki scheme: 8 radix-128 digit planes per row-scaled operand, exact INT32 accumulation in TMEM,
 * float64 recombination; agrees with the float64 product to ~2^-50 of |A_i||B_j| per entry).  M % 128 == 0, N % 64 == 0,
 * K % 64 == 0; work: device scratch of at least b200_ozaki_gemm_work_bytes(M, N, K) bytes.  The same kernels carry the
 * long-K panel updates inside b200_dev_chol_solve (there the operands are sliced once per finished panel). */
int b200_dev_ozaki_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                           void* work, size_t work_bytes, void* stream);
size_t b200_ozaki_gemm_work_bytes(int M, int N, int K);
/* digit planes per operand, NS: a float64 product costs NS (NS + 1) / 2 INT8 products */
int b200_ozaki_slices(void);
int b200_dev_transpose(const double* A, int lda, double* At, int ldat, int rows, int cols, void* stream);
/* np.linalg.eigh (lakernel.py:162, 201, 266): block Jacobi (pairs of 16-row blocks, 32x32 sub-problems in shared
 * memory).  A (n x n, lda) is destroyed and must be padded with the identity up to ntot = 16 * (ceil(n/16) rounded up
 * to even) rows and columns (lda, ldv >= ntot, even); Vt (ntot, ldv) receives the eigenvectors as ROWS, lam (n) the
 * eigenvalues (unsorted).  Returns the sweep count in *sweeps (host int, may be NULL).  This call synchronises the
 * stream once per sweep (convergence test). */
int b200_dev_eigh(double* A, int lda, int n, double* Vt, int ldv, double* lam, int max_sweeps, int* sweeps,
                  void* stream);

/* np.linalg.eigh (lakernel.py:162, 201, 266) for up to B200_MAXB independent matrices at once (the OutStamps of a
 * batch), every launch over all of them: blocked Householder tridiagonalisation, bisection, inverse iteration,
 * Cholesky-QR orthonormalisation, compact-WY back-transformation (csrc/trieig.cu).  A and Vt must be padded to a multiple
 * of 128 rows and columns (A with the identity), lda and ldv even.  Synchronises the stream once at the end; *sweeps is
 * set to 0.  A system whose orthonormalisation fails (linearly dependent inverse-iteration vectors) is solved again with
 * the block-Jacobi method of b200_dev_eigh; b200_eigh_fallback_count() counts those.  Environment: B200_EIGH=jacobi selects
 * the Jacobi solver for everything (then max_sweeps / *sweeps have the meaning of b200_dev_eigh and the padding rule is
 * b200_dev_eigh's). */
typedef struct b200_eigh_problem {
    double* A;   /* (ntot, lda) in, destroyed; identity padding as for b200_dev_eigh */
    double* Vt;  /* (ntot, ldv) out: eigenvectors as rows */
    double* lam; /* (n) out */
    int lda, ldv, n, pad_;
} b200_eigh_problem;
int b200_dev_eigh_batch(const b200_eigh_problem* problems, int nsys, int max_sweeps, int* sweeps, void* stream);
long long b200_eigh_fallback_count(void);
/* Householder tridiagonalisation A = Q T Q^T of one symmetric matrix (the first stage of the eigensolver; exposed for the
 * tests): A (n x n, lda) is overwritten -- row k holds reflector k at columns k+1.. (leading 1 stored) -- and the device
 * arrays d, e, tau (n doubles each) receive T's diagonal, its sub-diagonal (e[n-1] = 0) and the reflector scales. */
int b200_dev_tridiag(double* A, int lda, int n, double* d, double* e, double* tau, void* stream);

/* ---- 4. device: per-output-pixel Lagrange multiplier (stage b) --------------------------------- */
int b200_dev_lakernel1(const double* lam, const double* mPhalf, int ldp, int m, int n, double C, double targetleak,
                       double kCmin, double kCmax, int nbis, double* kappa, double* Sigma, double* UC, double* T,
                       int ldt, double smax, void* stream);
/* EigenKernel._call_single_kappa closed forms (lakernel.py:165-170). */
int b200_dev_eigen_single(const double* lam, const double* mPhalf, int ldp, int m, int n, double C, double kappa,
                          double* Sigma, double* UC, double* T, int ldt, void* stream);
int b200_dev_lsolve_sps(int N, double* A, double* x, const double* b, double* work, void* stream);
int b200_dev_build_reduced_T(const double* Nflat, const double* Dflat, const double* Eflat, const double* kappa,
                             int nv, int m, double ucmin, double smax, double* out_kappa, double* out_Sigma,
                             double* out_UC, double* out_w, int* out_iv, int* out_branch, void* stream);
/* D_p, N_pq, E_pq of the nv node solutions (lakernel.py:361-368); DpC/EpqC are D/C and E/C as handed to
 * build_reduced_T_wrap (lakernel.py:375-380).  Epq_in != NULL supplies an exact E (lakernel.py:709-716). */
int b200_dev_node_stats(const double* mB, int ldb, const double* Tpi, int ldt, size_t strideT, int nv, int m, int n,
                        const double* kappa_nodes, double Cnorm, double* Dp, double* Npq, double* Epq, double* DpC,
                        double* EpqC, const double* Epq_in, void* stream);
int b200_dev_rowdot(const double* X, int ldx, const double* Y, int ldy, int m, int n, double* out, int ostride,
                    void* stream);
/* single-kappa outputs (lakernel.py:312-316, 643-648): Sigma = N, UC = 1 - (kappa N + D)/C or, with E given,
 * UC = 1 + (E - 2D)/C; kappa map filled with the scalar. */
int b200_dev_single_kappa_maps(const double* D, const double* N, const double* E, int m, double kappa, double C,
                               double* kappa_out, double* Sigma_out, double* UC_out, void* stream);
/* EmpirKernel._call_single_kappa (lakernel.py:761-768): Ti[a, i] = max(rho_acc - hypot(dy, dx), 0), each row divided by
 * its sum; T (mpad, ldt) with rows m.. and columns n..npad-1 zero-filled. */
int b200_dev_empir_T(const double* inx, const double* iny, const double* outx, const double* outy, int m, int mpad, int n,
                     int npad, double rho_acc, double* T, int ldt, void* stream);
/* out[a] = scale * in[a] (kappa = out_kappa * C, lakernel.py:390, and the Eigen quirk lakernel.py:222) */
int b200_dev_scale(const double* in, double scale, int m, double* out, void* stream);

/* ---- 5. device: IterKernel (lakernel.py:397-443, 548-590, 615-619) ------------------------------ */
/* Per output pixel a at (outx[a], outy[a]): acceptance mask hypot(dy,dx) < rho_acc over the input pixels
 * (inx, iny), CG on the gathered sub-system of AA + diag_add I, result rounded to float32 and scattered into
 * Ti (m, ldt) (stored as f64).  niter/nsel (m) optional. */
int b200_dev_iter_cg(const double* AA, int lda, double diag_add, const double* mB, int ldb, int m, int n,
                     const double* inx, const double* iny, const double* outx, const double* outy, double rho_acc,
                     double rtol, int maxiter, double* Ti, int ldt, int* niter, int* nsel, void* stream);

/* ---- 6. device: apply T (stage c; coadd.py:1221-1363, 1976-1994) -------------------------------- */
typedef struct b200_finalize_args {
    const double* Tpi;     /* (nv, m, ldt) f64 node solutions (nv == 1: the solution itself) */
    size_t strideT;
    int ldt;
    const double* w;       /* (m, nv) node weights, or NULL when nv == 1 */
    int nv;
    const double* mB;      /* (m, ldb) -B/2 rows (for D); may be NULL */
    int ldb;
    int m, n, n2f, fade;
    const double* fade_w;  /* (2*fade) trapezoid weights s_k - sin(2 pi s_k)/(2 pi) (coadd.py:1269-1271) */
    const float* indata;   /* (n_inframe, ldi) f32 input layers in the stamp's pixel order */
    int ldi, n_inframe;
    const int* seg_end;    /* (nseg) exclusive end column of each (instamp,image) segment, ascending */
    const int* seg_img;    /* (nseg) image index of each segment */
    int nseg, n_img;
    float* T32;            /* (m, ldt32) f32 faded T out (nullable) */
    int ldt32;
    double* Ti64;          /* (m, ldt64) f64 pre-cast, pre-fade combined T out (nullable) */
    int ldt64;
    double* D;             /* (m) sum_i mB*Ti (nullable) */
    double* N;             /* (m) sum_i Ti^2  (nullable) */
    float* outimage;       /* (n_inframe, m) f32 */
    double* Tsum_image;    /* (m, n_img) f64 sums of the faded f32 T per image */
} b200_finalize_args;

int b200_dev_finalize(const b200_finalize_args* args, void* stream);
int b200_dev_stamp_maps(const double* kappa, const double* Sigma, const double* UC, int m, int n2f, int fade,
                        int clamp_iter, const double* fade_w, float* kappa32, float* Sigma32, float* UC32,
                        const double* Tsum_image, int n_img, int n2, double* Tsum_stamp, double* Tsum_inpix,
                        double* Neff, void* stream);
/* dst[l, y0+iy, x0+ix] += src[l, iy, ix] for an (nlayer, n2f, n2f) stamp into an (nlayer, side, side) f32 canvas. */
int b200_dev_accumulate(const void* src, int src_is_f64, int nlayer, int n2f, float* dst, int side, int y0, int x0,
                        void* stream);
/* Block._output_stamp_wrapper (coadd.py:1976-1994) for one output PSF of one stamp in ONE launch: the n_inframe coadded
 * layers into out_map (n_inframe, side, side), the float32 U/C, Sigma, kappa maps and the float64 Tsum_inpix, Neff maps
 * (n2f, n2f) into their (side, side) canvases, and T_weight[k * tw_stride] = Tsum_stamp[k] for the n_img input images. */
int b200_dev_accumulate_stamp(const float* outimage, int n_inframe, const float* UC, const float* Sigma,
                              const float* kappa, const double* Tsum_inpix, const double* Neff, const double* Tsum_stamp,
                              int n_img, int n2f, float* out_map, float* UC_map, float* Sigma_map, float* kappa_map,
                              float* Tsum_map, float* Neff_map, int side, int y0, int x0, float* T_weight,
                              int tw_stride, void* stream);

/* ---- 6. device: block output assembly (Block.build_output_file, coadd.py:2139-2176) ---------- */
/* out (nlayer, side-2fk, side-2fk) = in (nlayer, side, side) without its fade margin.  recover != 0 first divides the
 * trapezoid weights fade_w[0 .. 2fk-1] back out of the block boundary exactly as OutStamp.trapezoid(arr, fk,
 * recover_mode=True, pad_widths=(pb, pt, pl, pr)) does (coadd.py:1262-1292: sides B, T, L, R in turn, each division in
 * float64 rounded to float32).  in is not modified. */
int b200_dev_unfade_crop(const float* in, int nlayer, int side, int fk, int recover, int pb, int pt, int pl, int pr,
                         const double* fade_w, float* out, void* stream);
/* Block.compress_map (coadd.py:2086-2137): out[i] = clip(floor(coef * log10(max(in[i], 1e-32)) + 0.5), lo, hi) as
 * uint16 (is_unsigned, range 0..65535) or int16 (-32768..32767); float32 arithmetic as NumPy's, log10 correctly rounded. */
int b200_dev_compress_map(const float* in, long n, int coef, int is_unsigned, void* out, void* stream);

/* ---- 7. device: input-pixel partitioning (InImage.partition_pixels / extract_layers,
 *         coadd.py:333-360, 382-408) ------------------------------------------------------------------------ */
#define B200_PART_MAXSLOT 256 /* postage stamps one sparse-grid cell may touch */
/* One relevant cell of the sparse grid: detector rows bottom .. bottom+h-1, columns left .. left+w-1; its h*w output
 * positions start at ox[off], oy[off] in raster order (row by row), as _inpix2world2outpix returns them. */
typedef struct b200_part_cell {
    int bottom, left, h, w;
    long long off;
} b200_part_cell;
/* The double loop of coadd.py:333-360 for the ncell relevant cells, in the order given (the reference's traversal).
 * A pixel is kept when lower < x < upper and lower < y < upper (lower = -n2 - 0.5, upper = NsideP + n2 - 0.5), its
 * mask byte (sca x sca) is non-zero and its stamp (j_st, i_st) = ((y - lower) // n2, (x - lower) // n2) has use[] set
 * (ns x ns, ns = n1P + 2).  Outputs with the reference's layout: pix_count (ns*ns) u32; y_idx, x_idx (ns*ns, npixmax)
 * u16; y_val, x_val (ns*ns, npixmax) f64, each stamp's list in traversal order (entries beyond pix_count are left
 * untouched: pass zero-filled arrays).  Scratch: sid_tmp (int) and rank_tmp (u32) per position, cellmeta (4 ints),
 * cellcnt and cellbase (B200_PART_MAXSLOT u32) per cell, run (ns*ns u32).  err (device int) is 0 on success, 1 if a
 * cell touches more than B200_PART_MAXSLOT stamps, 2 if a stamp receives more than npixmax pixels (the reference would
 * raise IndexError). */
int b200_dev_partition(const b200_part_cell* cells, int ncell, const double* ox, const double* oy,
                       const unsigned char* mask, int sca, const unsigned char* use, int ns, int n2, double lower,
                       double upper, int npixmax, int* sid_tmp, unsigned* rank_tmp, int* cellmeta, unsigned* cellcnt,
                       unsigned* cellbase, unsigned* run, unsigned* pix_count, unsigned short* y_idx,
                       unsigned short* x_idx, double* y_val, double* x_val, int* err, void* stream);
/* InStamp.__init__ (coadd.py:682-707) for input image number `image`: its per-stamp lists (x_val, y_val (nstamp, npixmax),
 * data (n_inframe, nstamp, max_count), pix_count (nstamp)) are copied to dst_off[sid] .. dst_off[sid] + pix_count[sid] - 1
 * of the block's concatenated pixel arrays gx, gy (npix) f64, gimg (npix) i32, gdata (n_inframe, npix) f32. */
int b200_dev_assemble_instamps(const double* x_val, const double* y_val, const float* data, int n_inframe, int nstamp,
                               int npixmax, int max_count, const unsigned* pix_count, const long long* dst_off, int image,
                               long long npix, double* gx, double* gy, int* gimg, float* gdata, void* stream);
/* InImage.extract_layers (coadd.py:396-404): data (n_inframe, nstamp, max_count) f32 = indata (n_inframe, sca, sca) at
 * the listed pixels, zero beyond pix_count. */
int b200_dev_extract_layers(const float* indata, int n_inframe, int sca, const unsigned short* y_idx,
                            const unsigned short* x_idx, const unsigned* pix_count, int nstamp, int npixmax,
                            int max_count, float* data, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYIMCOM_B200_H */
