// C ABI of libpyimcom_b200.so (declared in include/pyimcom_b200.h): error plumbing, the grow-only device
// scratch of the host-seam functions, and thin extern "C" wrappers over the launchers in kernels.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "ozaki.cuh"

namespace b200 {

long long g_launches = 0;

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return 0;
    set_error("%s:%d: %s failed: %s (%s)", file, line, what, cudaGetErrorName(e), cudaGetErrorString(e));
    return 1000 + (int)e;
}

// ---- grow-only scratch ---------------------------------------------------------------------------
constexpr int NSLOT = 16;
static void* g_slot[NSLOT] = {nullptr};
static size_t g_slot_bytes[NSLOT] = {0};

int scratch(int slot, size_t bytes, void** out) {
    B200_REQUIRE(slot >= 0 && slot < NSLOT, "bad scratch slot");
    if (bytes < 256) bytes = 256;
    if (g_slot_bytes[slot] < bytes) {
        if (g_slot[slot]) {
            B200_CUDA(cudaDeviceSynchronize());
            B200_CUDA(cudaFree(g_slot[slot]));
            g_slot[slot] = nullptr;
            g_slot_bytes[slot] = 0;
        }
        const size_t want = bytes + bytes / 4;
        B200_CUDA(cudaMalloc(&g_slot[slot], want));
        g_slot_bytes[slot] = want;
    }
    *out = g_slot[slot];
    return 0;
}

void scratch_release() {
    for (int i = 0; i < NSLOT; i++) {
        if (g_slot[i]) cudaFree(g_slot[i]);
        g_slot[i] = nullptr;
        g_slot_bytes[i] = 0;
    }
}

// ---- per-launch profiling -------------------------------------------------------------------------
struct ProfRec {
    cudaEvent_t a, b;
    double work;
    int kind;
};
static int g_prof_on = 0;
static std::vector<ProfRec> g_prof;
static std::vector<size_t> g_prof_open;  // begin/end pairs nest (the eigensolver's scope holds its GEMMs' scopes)

// events are pooled: creating them inside the timed region would put driver calls between the launches
static std::vector<cudaEvent_t> g_event_pool;
static size_t g_event_next = 0;

static bool prof_event(cudaEvent_t* e) {
    if (g_event_next == g_event_pool.size()) {
        cudaEvent_t ne;
        if (cudaEventCreate(&ne) != cudaSuccess) return false;
        g_event_pool.push_back(ne);
    }
    *e = g_event_pool[g_event_next++];
    return true;
}

void prof_begin(int kind, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfRec r;
    r.kind = kind;
    r.work = 0.0;
    if (!prof_event(&r.a) || !prof_event(&r.b)) return;
    cudaEventRecord(r.a, st);
    g_prof_open.push_back(g_prof.size());
    g_prof.push_back(r);
}

void prof_end(double work, cudaStream_t st) {
    if (!g_prof_on || g_prof_open.empty()) return;
    ProfRec& r = g_prof[g_prof_open.back()];
    g_prof_open.pop_back();
    r.work = work;
    cudaEventRecord(r.b, st);
}

static void prof_reset() {
    g_prof.clear();
    g_prof_open.clear();
    g_event_next = 0;  // the pooled events are reused by the next recording
}

}  // namespace b200

using namespace b200;

#define ST(s) ((cudaStream_t)(s))

template <typename T>
static int up(int slot, const T* h, size_t count, T** d) {
    void* p = nullptr;
    if (int rc = scratch(slot, sizeof(T) * (count ? count : 1), &p)) return rc;
    if (count) B200_CUDA(cudaMemcpyAsync(p, h, sizeof(T) * count, cudaMemcpyHostToDevice, 0));
    *d = (T*)p;
    return 0;
}
template <typename T>
static int dev(int slot, size_t count, T** d) {
    void* p = nullptr;
    if (int rc = scratch(slot, sizeof(T) * (count ? count : 1), &p)) return rc;
    *d = (T*)p;
    return 0;
}
template <typename T>
static int down(T* h, const T* d, size_t count) {
    if (count) B200_CUDA(cudaMemcpyAsync(h, d, sizeof(T) * count, cudaMemcpyDeviceToHost, 0));
    return 0;
}
#define TRY(x)               \
    do {                     \
        if (int _r = (x)) return _r; \
    } while (0)

extern "C" {

const char* b200_last_error(void) { return g_err; }
int b200_version(void) { return 100; }
long long b200_launch_count(void) { return g_launches; }
int b200_release_scratch(void) {
    scratch_release();
    return 0;
}
int b200_profile(int on) {
    prof_reset();
    if (on) {  // pre-create events for ~8k launches so that the first recorded step does not pay for them
        while (g_event_pool.size() < 16384) {
            cudaEvent_t e;
            B200_CUDA(cudaEventCreate(&e));
            g_event_pool.push_back(e);
        }
    }
    g_prof_on = on;
    return 0;
}
int b200_profile_read(int kind, double* ms, double* work, long long* count) {
    B200_CUDA(cudaDeviceSynchronize());
    double t = 0.0, w = 0.0;
    long long c = 0;
    for (auto& r : g_prof) {
        if (r.kind != kind) continue;
        float e = 0.f;
        if (cudaEventElapsedTime(&e, r.a, r.b) != cudaSuccess) {
            (void)cudaGetLastError();  // a scope that never closed: drop it, and do not leave the error for the next call
            continue;
        }
        t += e;
        w += r.work;
        c++;
    }
    if (ms) *ms = t;
    if (work) *work = w;
    if (count) *count = c;
    return 0;
}

// ---- 1. host seam ----------------------------------------------------------------------------------
static int interp_host(int sym, const double* infunc, int nlayer, int ngy, int ngx, const double* xpos,
                       const double* ypos, long nout, double* fhatout) {
    B200_REQUIRE(nlayer >= 0 && ngy >= 0 && ngx >= 0 && nout >= 0, "negative size");
    if (nout == 0 || nlayer == 0) return 0;
    double *df, *dx, *dy, *dout;
    TRY(up(0, infunc, (size_t)nlayer * ngy * ngx, &df));
    TRY(up(1, xpos, (size_t)nout, &dx));
    TRY(up(2, ypos, (size_t)nout, &dy));
    TRY(up(3, fhatout, (size_t)nlayer * nout, &dout));  // off-grid points keep the caller's values
    TRY((sym ? launch_iD5512C_sym : launch_iD5512C)(df, nlayer, ngy, ngx, dx, dy, nout, dout, 0));
    TRY(down(fhatout, dout, (size_t)nlayer * nout));
    B200_CUDA(cudaStreamSynchronize(0));
    return 0;
}

int b200_iD5512C(const double* infunc, int nlayer, int ngy, int ngx, const double* xpos, const double* ypos, long nout,
                 double* fhatout) {
    return interp_host(0, infunc, nlayer, ngy, ngx, xpos, ypos, nout, fhatout);
}

int b200_iD5512C_sym(const double* infunc, int nlayer, int ngy, int ngx, const double* xpos, const double* ypos,
                     long nout, double* fhatout) {
    return interp_host(1, infunc, nlayer, ngy, ngx, xpos, ypos, nout, fhatout);
}

int b200_gridD5512C(const double* infunc, int ngy, int ngx, const double* xpos, const double* ypos, long npi, int nxo,
                    int nyo, double* fhatout) {
    B200_REQUIRE(ngy >= 0 && ngx >= 0 && npi >= 0 && nxo >= 0 && nyo >= 0, "negative size");
    if (npi == 0 || nxo == 0 || nyo == 0) return 0;
    double *df, *dx, *dy, *dout;
    TRY(up(0, infunc, (size_t)ngy * ngx, &df));
    TRY(up(1, xpos, (size_t)npi * nxo, &dx));
    TRY(up(2, ypos, (size_t)npi * nyo, &dy));
    TRY(dev(3, (size_t)npi * nxo * nyo, &dout));
    TRY(launch_gridD5512C(df, ngy, ngx, dx, dy, npi, nxo, nyo, dout, 0));
    TRY(down(fhatout, dout, (size_t)npi * nxo * nyo));
    B200_CUDA(cudaStreamSynchronize(0));
    return 0;
}

int b200_iD5512C_getw(double* w, double fh) {
    double* dw;
    TRY(dev(0, 10, &dw));
    TRY(launch_getw(dw, fh, 0));
    TRY(down(w, dw, 10));
    B200_CUDA(cudaStreamSynchronize(0));
    return 0;
}

int b200_lakernel1(const double* lam, const double* mPhalf, long m, long n, double C, double targetleak, double kCmin,
                   double kCmax, int nbis, double* kappa, double* Sigma, double* UC, double* T, double smax) {
    B200_REQUIRE(m >= 0 && n >= 0, "negative size");
    if (m == 0) return 0;
    double *dl, *dP, *dk, *dS, *dU, *dT;
    TRY(up(0, lam, (size_t)n, &dl));
    TRY(up(1, mPhalf, (size_t)m * n, &dP));
    TRY(dev(2, (size_t)3 * m, &dk));
    dS = dk + m;
    dU = dS + m;
    TRY(dev(3, (size_t)m * n, &dT));
    TRY(launch_lakernel1(dl, dP, (int)n, (int)m, (int)n, C, targetleak, kCmin, kCmax, nbis, dk, dS, dU, dT, (int)n, smax,
                         0));
    TRY(down(kappa, dk, (size_t)m));
    TRY(down(Sigma, dS, (size_t)m));
    TRY(down(UC, dU, (size_t)m));
    TRY(down(T, dT, (size_t)m * n));
    B200_CUDA(cudaStreamSynchronize(0));
    return 0;
}

int b200_lsolve_sps(int N, double* A, double* x, const double* b) {
    B200_REQUIRE(N >= 0, "negative size");
    if (N == 0) return 0;
    double *dA, *dx, *db, *dw;
    TRY(up(0, A, (size_t)N * N, &dA));
    TRY(up(1, b, (size_t)N, &db));
    TRY(dev(2, (size_t)N, &dx));
    TRY(dev(3, (size_t)N, &dw));
    TRY(launch_lsolve_sps(N, dA, dx, db, dw, 0));
    TRY(down(x, dx, (size_t)N));
    TRY(down(A, dA, (size_t)N * N));  // the reference destroys A (holds the Cholesky factor afterwards)
    B200_CUDA(cudaStreamSynchronize(0));
    return 0;
}

int b200_build_reduced_T_wrap(const double* Nflat, const double* Dflat, const double* Eflat, const double* kappa, int nv,
                              long m, double ucmin, double smax, double* out_kappa, double* out_Sigma, double* out_UC,
                              double* out_w, int32_t* out_iv, int32_t* out_branch) {
    B200_REQUIRE(m >= 0 && nv >= 1, "bad size");
    if (m == 0) return 0;
    double *dN, *dD, *dE, *dk, *dout, *dw;
    int* di;
    TRY(up(0, Nflat, (size_t)m * nv * nv, &dN));
    TRY(up(1, Dflat, (size_t)m * nv, &dD));
    TRY(up(2, Eflat, (size_t)m * nv * nv, &dE));
    TRY(up(3, kappa, (size_t)nv, &dk));
    TRY(dev(4, (size_t)3 * m, &dout));
    TRY(dev(5, (size_t)m * nv, &dw));
    TRY(dev(6, (size_t)2 * m, &di));
    TRY(launch_build_reduced_T(dN, dD, dE, dk, nv, (int)m, ucmin, smax, dout, dout + m, dout + 2 * m, dw, di, di + m, 0));
    TRY(down(out_kappa, dout, (size_t)m));
    TRY(down(out_Sigma, dout + m, (size_t)m));
    TRY(down(out_UC, dout + 2 * m, (size_t)m));
    TRY(down(out_w, dw, (size_t)m * nv));
    if (out_iv) TRY(down(out_iv, di, (size_t)m));
    if (out_branch) TRY(down(out_branch, di + m, (size_t)m));
    B200_CUDA(cudaStreamSynchronize(0));
    return 0;
}

// ---- 2. device: interpolation / assembly -----------------------------------------------------------
int b200_dev_iD5512C(const double* f, int nlayer, int ngy, int ngx, const double* x, const double* y, long nout,
                     double* out, void* s) {
    return launch_iD5512C(f, nlayer, ngy, ngx, x, y, nout, out, ST(s));
}
int b200_dev_iD5512C_sym(const double* f, int nlayer, int ngy, int ngx, const double* x, const double* y, long nout,
                         double* out, void* s) {
    return launch_iD5512C_sym(f, nlayer, ngy, ngx, x, y, nout, out, ST(s));
}
int b200_dev_gridD5512C(const double* f, int ngy, int ngx, const double* x, const double* y, long npi, int nxo, int nyo,
                        double* out, void* s) {
    return launch_gridD5512C(f, ngy, ngx, x, y, npi, nxo, nyo, out, ST(s));
}
int b200_dev_layout_tables(const double* src, int ntab, int ns, int pad, int ngrid, int poly, double* dst, void* s) {
    return launch_layout_tables(src, ntab, ns, pad, ngrid, poly, dst, ST(s));
}
int b200_dev_cmul_conj(const double* ar, const double* ai, const double* br, const double* bi, long long per,
                       long long stride2, long long n1, double im_sign, double* gr, double* gi, void* s) {
    return launch_cmul_conj(ar, ai, br, bi, per, stride2, n1, im_sign, gr, gi, ST(s));
}
int b200_dev_gather_stamp(const int* idx, int n, int npad, const double* sx, const double* sy, const int* scode,
                          const float* sdata, long src_ld, int n_inframe, double* px, double* py, int* pcode,
                          float* indata, int ldi, void* s) {
    return launch_gather_stamp(idx, n, npad, sx, sy, scode, sdata, src_ld, n_inframe, px, py, pcode, indata, ldi, ST(s));
}
int b200_dev_build_A(const double* px, const double* py, const int* pcode, int n, int npad, const double* tables,
                     const b200_table_ref* lut, int nimg, int ncode, int ngrid, double dscale, double nc,
                     double flat_penalty, double* A, int lda, double diag_add, int poly, void* s) {
    return launch_build_A(px, py, pcode, n, npad, tables, lut, nimg, ncode, ngrid, dscale, nc, flat_penalty, A, lda,
                          diag_add, poly, ST(s));
}
int b200_dev_pair_blocks(const double* gx, const double* gy, const int* gimg, const b200_pair_desc* descs,
                         const int* tile_prefix, int npair, int ntiles, const double* tables,
                         const b200_table_ref* lut, int nimg, int ngrid, double dscale, double nc, double flat_penalty,
                         int poly, double* pool, double points, void* s) {
    return launch_pair_blocks(gx, gy, gimg, descs, tile_prefix, npair, ntiles, tables, lut, nimg, ngrid, dscale, nc,
                              flat_penalty, poly, pool, points, ST(s));
}
int b200_dev_assemble_A(const b200_asm_desc* desc, const int* gidx, int n, int npad, const double* pool, double* A,
                        int lda, double diag_add, void* s) {
    B200_REQUIRE(desc != nullptr, "assemble_A: null descriptor");
    return launch_assemble_A(*desc, gidx, n, npad, pool, A, lda, diag_add, ST(s));
}
int b200_dev_build_B(const double* px, const double* py, const int* pcode, int n, int npad, const double* tables,
                     const long long* lut_io, int n_out, int ngrid, double dscale, double nc, int n2f, int mpad,
                     double x0out, double y0out, double* B, int ldb, size_t strideB, void* s) {
    return launch_build_B(px, py, pcode, n, npad, tables, lut_io, n_out, ngrid, dscale, nc, n2f, mpad, x0out, y0out, B,
                          ldb, strideB, ST(s));
}

// ---- 3. device: dense linear algebra ----------------------------------------------------------------
int b200_dev_chol_solve(const b200_solve_sys* sys, int nsys, int do_factor, int do_solve, void* s) {
    return launch_chol_solve(sys, nsys, do_factor, do_solve, ST(s));
}
int b200_dev_pad_system(double* W, int ldw, int n, int npad, const double* A, int lda, const double* incs, int ninc,
                        void* s) {
    return launch_pad_system(W, ldw, n, npad, A, lda, incs, ninc, ST(s));
}
int b200_dev_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                     int accumulate, void* s) {
    return launch_gemm_nt(A, lda, B, ldb, C, ldc, M, N, K, accumulate, ST(s));
}
int b200_dev_ozaki_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                           void* work, size_t work_bytes, void* s) {
    return launch_ozaki_gemm_nt(A, lda, B, ldb, C, ldc, M, N, K, work, work_bytes, ST(s));
}
size_t b200_ozaki_gemm_work_bytes(int M, int N, int K) { return oz_gemm_work_bytes(M, N, K); }
size_t b200_chol_work_bytes(int npad, int mpad) { return chol_work_bytes(npad, mpad); }
int b200_ozaki_slices(void) { return OZ_NS; }
int b200_dev_transpose(const double* A, int lda, double* At, int ldat, int rows, int cols, void* s) {
    return launch_transpose(A, lda, At, ldat, rows, cols, ST(s));
}
int b200_dev_eigh_batch(const b200_eigh_problem* problems, int nsys, int max_sweeps, int* sweeps, void* s) {
    B200_REQUIRE(problems != nullptr || nsys <= 0, "eigh_batch: null problem list");
    const char* e = getenv("B200_EIGH");  // "jacobi": the block-Jacobi solver of round 1 (eigen.cu)
    if (e && e[0] == 'j') return launch_jacobi_eigh_batch(problems, nsys, max_sweeps, sweeps, ST(s));
    if (sweeps) *sweeps = 0;
    return launch_tri_eigh_batch(problems, nsys, ST(s));
}
long long b200_eigh_fallback_count(void) { return eigh_fallback_count(); }
int b200_dev_tridiag(double* A, int lda, int n, double* d, double* e, double* tau, void* s) {
    return launch_tridiag(A, lda, n, d, e, tau, ST(s));
}
int b200_dev_eigh(double* A, int lda, int n, double* Vt, int ldv, double* lam, int max_sweeps, int* sweeps, void* s) {
    return launch_jacobi_eigh(A, lda, n, Vt, ldv, lam, max_sweeps, sweeps, ST(s));
}

// ---- 4. device: kappa kernels ------------------------------------------------------------------------
int b200_dev_lakernel1(const double* lam, const double* mPhalf, int ldp, int m, int n, double C, double targetleak,
                       double kCmin, double kCmax, int nbis, double* kappa, double* Sigma, double* UC, double* T, int ldt,
                       double smax, void* s) {
    return launch_lakernel1(lam, mPhalf, ldp, m, n, C, targetleak, kCmin, kCmax, nbis, kappa, Sigma, UC, T, ldt, smax,
                            ST(s));
}
int b200_dev_eigen_single(const double* lam, const double* mPhalf, int ldp, int m, int n, double C, double kappa,
                          double* Sigma, double* UC, double* T, int ldt, void* s) {
    return launch_eigen_single(lam, mPhalf, ldp, m, n, C, kappa, Sigma, UC, T, ldt, ST(s));
}
int b200_dev_lsolve_sps(int N, double* A, double* x, const double* b, double* work, void* s) {
    return launch_lsolve_sps(N, A, x, b, work, ST(s));
}
int b200_dev_build_reduced_T(const double* Nflat, const double* Dflat, const double* Eflat, const double* kappa, int nv,
                             int m, double ucmin, double smax, double* out_kappa, double* out_Sigma, double* out_UC,
                             double* out_w, int* out_iv, int* out_branch, void* s) {
    return launch_build_reduced_T(Nflat, Dflat, Eflat, kappa, nv, m, ucmin, smax, out_kappa, out_Sigma, out_UC, out_w,
                                  out_iv, out_branch, ST(s));
}
int b200_dev_node_stats(const double* mB, int ldb, const double* Tpi, int ldt, size_t strideT, int nv, int m, int n,
                        const double* kappa_nodes, double Cnorm, double* Dp, double* Npq, double* Epq, double* DpC,
                        double* EpqC, const double* Epq_in, void* s) {
    return launch_node_stats(mB, ldb, Tpi, ldt, strideT, nv, m, n, kappa_nodes, Cnorm, Dp, Npq, Epq, DpC, EpqC, Epq_in,
                             ST(s));
}
int b200_dev_rowdot(const double* X, int ldx, const double* Y, int ldy, int m, int n, double* out, int ostride, void* s) {
    return launch_rowdot(X, ldx, Y, ldy, m, n, out, ostride, ST(s));
}
int b200_dev_single_kappa_maps(const double* D, const double* N, const double* E, int m, double kappa, double C,
                               double* kappa_out, double* Sigma_out, double* UC_out, void* s) {
    return launch_single_kappa_maps(D, N, E, m, kappa, C, kappa_out, Sigma_out, UC_out, ST(s));
}
int b200_dev_empir_T(const double* inx, const double* iny, const double* outx, const double* outy, int m, int mpad, int n,
                     int npad, double rho_acc, double* T, int ldt, void* s) {
    return launch_empir_T(inx, iny, outx, outy, m, mpad, n, npad, rho_acc, T, ldt, ST(s));
}
int b200_dev_scale(const double* in, double scale, int m, double* out, void* s) {
    return launch_scale(in, scale, m, out, ST(s));
}

// ---- 5. device: IterKernel --------------------------------------------------------------------------
int b200_dev_iter_cg(const double* AA, int lda, double diag_add, const double* mB, int ldb, int m, int n,
                     const double* inx, const double* iny, const double* outx, const double* outy, double rho_acc,
                     double rtol, int maxiter, double* Ti, int ldt, int* niter, int* nsel, void* s) {
    return launch_iter_cg(AA, lda, diag_add, mB, ldb, m, n, inx, iny, outx, outy, rho_acc, rtol, maxiter, Ti, ldt,
                          niter, nsel, ST(s));
}

// ---- 6. device: apply T -----------------------------------------------------------------------------
int b200_dev_finalize(const b200_finalize_args* args, void* s) {
    B200_REQUIRE(args != nullptr, "null args");
    return launch_finalize(*args, ST(s));
}
int b200_dev_stamp_maps(const double* kappa, const double* Sigma, const double* UC, int m, int n2f, int fade,
                        int clamp_iter, const double* fade_w, float* kappa32, float* Sigma32, float* UC32,
                        const double* Tsum_image, int n_img, int n2, double* Tsum_stamp, double* Tsum_inpix, double* Neff,
                        void* s) {
    return launch_stamp_maps(kappa, Sigma, UC, m, n2f, fade, clamp_iter, fade_w, kappa32, Sigma32, UC32, Tsum_image,
                             n_img, n2, Tsum_stamp, Tsum_inpix, Neff, ST(s));
}
int b200_dev_accumulate(const void* src, int src_is_f64, int nlayer, int n2f, float* dst, int side, int y0, int x0,
                        void* s) {
    return launch_accumulate(src, src_is_f64, nlayer, n2f, dst, side, y0, x0, ST(s));
}
int b200_dev_accumulate_stamp(const float* outimage, int n_inframe, const float* UC, const float* Sigma,
                              const float* kappa, const double* Tsum_inpix, const double* Neff, const double* Tsum_stamp,
                              int n_img, int n2f, float* out_map, float* UC_map, float* Sigma_map, float* kappa_map,
                              float* Tsum_map, float* Neff_map, int side, int y0, int x0, float* T_weight,
                              int tw_stride, void* s) {
    return launch_accumulate_stamp(outimage, n_inframe, UC, Sigma, kappa, Tsum_inpix, Neff, Tsum_stamp, n_img, n2f, out_map,
                                   UC_map, Sigma_map, kappa_map, Tsum_map, Neff_map, side, y0, x0, T_weight, tw_stride,
                                   ST(s));
}
int b200_dev_unfade_crop(const float* in, int nlayer, int side, int fk, int recover, int pb, int pt, int pl, int pr,
                         const double* fade_w, float* out, void* s) {
    return launch_unfade_crop(in, nlayer, side, fk, recover, pb, pt, pl, pr, fade_w, out, ST(s));
}
int b200_dev_compress_map(const float* in, long n, int coef, int is_unsigned, void* out, void* s) {
    return launch_compress_map(in, n, coef, is_unsigned, out, ST(s));
}
int b200_dev_partition(const b200_part_cell* cells, int ncell, const double* ox, const double* oy,
                       const unsigned char* mask, int sca, const unsigned char* use, int ns, int n2, double lower,
                       double upper, int npixmax, int* sid_tmp, unsigned* rank_tmp, int* cellmeta, unsigned* cellcnt,
                       unsigned* cellbase, unsigned* run, unsigned* pix_count, unsigned short* y_idx,
                       unsigned short* x_idx, double* y_val, double* x_val, int* err, void* s) {
    return launch_partition(cells, ncell, ox, oy, mask, sca, use, ns, n2, lower, upper, npixmax, sid_tmp, rank_tmp,
                            cellmeta, cellcnt, cellbase, run, pix_count, y_idx, x_idx, y_val, x_val, err, ST(s));
}
int b200_dev_assemble_instamps(const double* x_val, const double* y_val, const float* data, int n_inframe, int nstamp,
                               int npixmax, int max_count, const unsigned* pix_count, const long long* dst_off, int image,
                               long long npix, double* gx, double* gy, int* gimg, float* gdata, void* s) {
    return launch_assemble_instamps(x_val, y_val, data, n_inframe, nstamp, npixmax, max_count, pix_count, dst_off, image,
                                    npix, gx, gy, gimg, gdata, ST(s));
}
int b200_dev_extract_layers(const float* indata, int n_inframe, int sca, const unsigned short* y_idx,
                            const unsigned short* x_idx, const unsigned* pix_count, int nstamp, int npixmax,
                            int max_count, float* data, void* s) {
    return launch_extract_layers(indata, n_inframe, sca, y_idx, x_idx, pix_count, nstamp, npixmax, max_count, data, ST(s));
}

}  // extern "C"
