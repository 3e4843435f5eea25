/*
 * oracle/croutines.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the numerical primitives on PyIMCOM's per-postage-stamp
 * coaddition hot path.  The real third-party implementation (furry_parakeet
 * pyimcom_croutines.c, requirements.txt:15, unpinned) is absent from /root/reference; its
 * published algorithm is the Numba twin src/pyimcom/routine.py, which the reference's own
 * tests/pyimcom/test_routine.py pins to the C code at 1e-9.  Every function below cites the
 * routine.py / lakernel.py lines it follows.  Pinned against the reference itself (run in the
 * build container) by tests/golden/make_golden.py -> tests/golden/*.npz.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path never does.
 *
 * Build: see oracle/build.py (gcc -O2 -fopenmp -shared -fPIC, -ffp-contract=off so that the
 * accumulation matches the Numba/C reference statement by statement).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* routine.py:29-122 -- the ten D5512 weights: 5 even + 5 odd polynomials in fh = frac - 1/2 */
static void orc_getw(double *w, double fh) {
    static const double ce[5][5] = {
        {+1.651881673372979740e-05, -3.145538007199505447e-04, +1.793518183780194427e-03,
         -2.904014557029917318e-03, +6.187591260980151433e-04},
        {-1.146756217210629335e-04, +2.883845374976550142e-03, -1.857047531896089884e-02,
         +3.147734488597204311e-02, -6.753293626461192439e-03},
        {+3.256838096371517067e-04, -9.702063770653997568e-03, +8.678848026470635524e-02,
         -1.659182651092198924e-01, +3.620560878249733799e-02},
        {-4.541830837949564726e-04, +1.494862093737218955e-02, -1.668775957435094937e-01,
         +5.879306056792649171e-01, -1.367845996704077915e-01},
        {+2.266560930061513573e-04, -7.815848920941316502e-03, +9.686607348538181506e-02,
         -4.505856722239036105e-01, +6.067135256905490381e-01}};
    static const double co[5][5] = {
        {-3.486978652054735998e-06, +6.753750285320532433e-05, -3.871378836550175566e-04,
         +6.279918076641771273e-04, -1.338434614116611838e-04},
        {+3.121412120355294799e-05, -8.040343683015897672e-04, +5.209574765466357636e-03,
         -8.847326408846412429e-03, +1.898674086370833597e-03},
        {-1.243658986204533102e-04, +3.804930695189636097e-03, -3.434861846914529643e-02,
         +6.581033749134083954e-02, -1.436476114189205733e-02},
        {+2.894406669584551734e-04, -9.794291009695265532e-03, +1.104231510875857830e-01,
         -3.906954914039130755e-01, +9.092432925988773451e-02},
        {-4.336085507644610966e-04, +1.537862263741893339e-02, -1.925091434770601628e-01,
         +8.993141455798455697e-01, -1.213035309579723942e+00}};
    const double fh2 = fh * fh;
    for (int k = 0; k < 5; k++) {
        double e = (((ce[k][0] * fh2 + ce[k][1]) * fh2 + ce[k][2]) * fh2 + ce[k][3]) * fh2 + ce[k][4];
        double o = ((((co[k][0] * fh2 + co[k][1]) * fh2 + co[k][2]) * fh2 + co[k][3]) * fh2 + co[k][4]) * fh;
        w[k] = e + o;
        w[9 - k] = e - o;
    }
}

void orc_iD5512C_getw(double *w, double fh) { orc_getw(w, fh); }

/* one scattered point; returns 0 if the point is off the grid (routine.py:160-181) */
static int orc_point(const double *f, long nlayer, long ngy, long ngx, double x, double y,
                     double *out, long ostride) {
    int32_t xi = (int32_t)x, yi = (int32_t)y;
    if (xi < 4 || xi >= ngx - 5 || yi < 4 || yi >= ngy - 5) return 0;
    double wx[10], wy[10];
    orc_getw(wx, x - xi - 0.5);
    orc_getw(wy, y - yi - 0.5);
    for (long l = 0; l < nlayer; l++) {
        const double *g = f + l * ngy * ngx;
        double acc = 0.0;
        for (int i = 0; i < 10; i++) {
            const double *row = g + (long)(yi - 4 + i) * ngx + (xi - 4);
            double strip = 0.0;
            for (int j = 0; j < 10; j++) strip += wx[j] * row[j];
            acc += strip * wy[i];
        }
        out[l * ostride] = acc;
    }
    return 1;
}

/* routine.py:125-181 */
void orc_iD5512C(const double *infunc, long nlayer, long ngy, long ngx, const double *xpos,
                 const double *ypos, long nout, double *fhatout, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
    for (long ip = 0; ip < nout; ip++)
        orc_point(infunc, nlayer, ngy, ngx, xpos[ip], ypos[ip], fhatout + ip, nout);
}

/* routine.py:184-253 : upper triangle of a sqrt(nout) x sqrt(nout) point matrix, then mirror */
void orc_iD5512C_sym(const double *infunc, long nlayer, long ngy, long ngx, const double *xpos,
                     const double *ypos, long nout, double *fhatout, int nthreads) {
    long sq = (long)(int32_t)sqrt((double)(nout + 1));
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads) if (nthreads > 1)
    for (long i1 = 0; i1 < sq; i1++)
        for (long i2 = i1; i2 < sq; i2++) {
            long ip = i1 * sq + i2;
            orc_point(infunc, nlayer, ngy, ngx, xpos[ip], ypos[ip], fhatout + ip, nout);
        }
    for (long i1 = 1; i1 < sq; i1++)
        for (long i2 = 0; i2 < i1; i2++)
            for (long l = 0; l < nlayer; l++)
                fhatout[l * nout + i1 * sq + i2] = fhatout[l * nout + i2 * sq + i1];
}

/* routine.py:256-338 : separable output grid; off-grid columns/rows get zero weights */
void orc_gridD5512C(const double *infunc, long ngy, long ngx, const double *xpos, const double *ypos,
                    long npi, long nxo, long nyo, double *fhatout, int nthreads) {
#pragma omp parallel num_threads(nthreads) if (nthreads > 1)
    {
        double *wx = (double *)malloc(sizeof(double) * 10 * nxo);
        double *wy = (double *)malloc(sizeof(double) * 10 * nyo);
        int32_t *xi = (int32_t *)malloc(sizeof(int32_t) * nxo);
        int32_t *yi = (int32_t *)malloc(sizeof(int32_t) * nyo);
#pragma omp for schedule(static)
        for (long p = 0; p < npi; p++) {
            for (long ix = 0; ix < nxo; ix++) {
                double x = xpos[p * nxo + ix];
                xi[ix] = (int32_t)x;
                if (xi[ix] < 4 || xi[ix] >= ngx - 5) {
                    xi[ix] = 4;
                    for (int k = 0; k < 10; k++) wx[ix * 10 + k] = 0.0;
                    continue;
                }
                orc_getw(wx + ix * 10, x - xi[ix] - 0.5);
            }
            for (long iy = 0; iy < nyo; iy++) {
                double y = ypos[p * nyo + iy];
                yi[iy] = (int32_t)y;
                if (yi[iy] < 4 || yi[iy] >= ngy - 5) {
                    yi[iy] = 4;
                    for (int k = 0; k < 10; k++) wy[iy * 10 + k] = 0.0;
                    continue;
                }
                orc_getw(wy + iy * 10, y - yi[iy] - 0.5);
            }
            double *o = fhatout + p * nyo * nxo;
            for (long iy = 0; iy < nyo; iy++)
                for (long ix = 0; ix < nxo; ix++) {
                    double acc = 0.0;
                    for (int i = 0; i < 10; i++) {
                        const double *row = infunc + (long)(yi[iy] - 4 + i) * ngx + (xi[ix] - 4);
                        double strip = 0.0;
                        for (int j = 0; j < 10; j++) strip += wx[ix * 10 + j] * row[j];
                        acc += strip * wy[iy * 10 + i];
                    }
                    *o++ = acc;
                }
        }
        free(wx); free(wy); free(xi); free(yi);
    }
}

/* routine.py:341-430 : per-output-pixel kappa bisection over the eigen-spectrum */
void orc_lakernel1(const double *lam, const double *mPhalf, long m, long n, double C,
                   double targetleak, double kCmin, double kCmax, long nbis, double *kappa,
                   double *Sigma, double *UC, double *T, double smax, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
    for (long a = 0; a < m; a++) {
        const double *P = mPhalf + a * n;
        double factor = sqrt(kCmax / kCmin);
        double kap = sqrt(kCmax * kCmin);
        double s1, s2, var;
        for (long b = 0; b < nbis; b++) {
            s1 = s2 = 0.0;
            for (long i = 0; i < n; i++) {
                var = P[i] / (lam[i] + kap);
                s2 += var * var;
                s1 += (lam[i] + 2.0 * kap) * var * var;
            }
            double udc = 1.0 - s1 / C;
            factor = sqrt(factor);
            kap *= (udc > targetleak && s2 < smax) ? 1.0 / factor : factor;
        }
        s1 = s2 = 0.0;
        for (long i = 0; i < n; i++) {
            T[a * n + i] = var = P[i] / (lam[i] + kap);
            s2 += var * var;
            s1 += (lam[i] + 2.0 * kap) * var * var;
        }
        Sigma[a] = s2;
        kappa[a] = kap;
        UC[a] = 1.0 - s1 / C;
    }
}

/* routine.py:433-484 : unblocked in-place lower Cholesky + two substitutions; destroys A */
void orc_lsolve_sps(long N, double *A, double *x, const double *b) {
    for (long i = 0; i < N; i++) {
        for (long j = 0; j < i; j++) {
            double s = 0.0;
            for (long k = 0; k < j; k++) s += A[i * N + k] * A[j * N + k];
            A[i * N + j] = (A[i * N + j] - s) / A[j * N + j];
        }
        double s = 0.0;
        for (long k = 0; k < i; k++) s += A[i * N + k] * A[i * N + k];
        A[i * N + i] = sqrt(A[i * N + i] - s);
    }
    double *p1 = (double *)malloc(sizeof(double) * (N > 0 ? N : 1));
    for (long i = 0; i < N; i++) {
        double s = 0.0;
        for (long j = 0; j < i; j++) s += A[i * N + j] * p1[j];
        p1[i] = (b[i] - s) / A[i * N + i];
    }
    for (long i = N - 1; i >= 0; i--) {
        double s = 0.0;
        for (long j = i + 1; j < N; j++) s += A[j * N + i] * x[j];
        x[i] = (p1[i] - s) / A[i * N + i];
    }
    free(p1);
}

/* routine.py:487-588 : reduced-space kappa search over nv Cholesky nodes.
 * Also returns the bracket index iv and the 12-step branch word per pixel (bit k = 1 when step k
 * divided kappa) so that tests can compare the discrete decisions (SURVEY 8c "P-discrete"). */
void orc_build_reduced_T(const double *Nflat, const double *Dflat, const double *Eflat,
                         const double *kappa, long nv, long m, double ucmin, double smax,
                         double *out_kappa, double *out_Sigma, double *out_UC, double *out_w,
                         int32_t *out_iv, int32_t *out_branch) {
    long nv2 = nv * nv;
    double *M2d = (double *)malloc(sizeof(double) * nv2);
    double *w = (double *)malloc(sizeof(double) * nv);
    for (long a = 0; a < m; a++) {
        const double *Na = Nflat + a * nv2, *Ea = Eflat + a * nv2, *Da = Dflat + a * nv;
        long iv = nv - 1;
        double UC = ucmin * 10, S = smax / 10;
        while (iv > 0 && ucmin < UC && smax > S) {
            iv -= 1;
            S = Na[iv * (nv + 1)];
            UC = 1.0 - 2.0 * Da[iv] + Ea[iv * (nv + 1)];
        }
        double kappamid = sqrt(kappa[iv] * kappa[iv + 1]);
        double factor = pow(kappa[iv + 1] / kappa[iv], 0.25);
        int32_t branch = 0;
        if (out_iv) out_iv[a] = (int32_t)iv;
        for (int it = 0; it < 12; it++) {
            for (long p = 0; p < nv; p++)
                for (long q = 0; q <= p; q++) M2d[p * nv + q] = Ea[p + nv * q] + kappamid * Na[p + nv * q];
            orc_lsolve_sps(nv, M2d, w, Da);
            for (long p = 0; p < nv; p++) out_w[a * nv + p] = w[p];
            S = 0.0;
            for (long p = 0; p < nv; p++) {
                double s = 0.0;
                for (long q = 0; q < nv; q++) s += Na[p + nv * q] * w[q];
                S += s * w[p];
            }
            UC = 1.0 - kappamid * S;
            for (long p = 0; p < nv; p++) UC -= Da[p] * w[p];
            int down = (ucmin < UC && smax > S);
            if (down) branch |= (1 << it);
            kappamid *= down ? 1.0 / factor : factor;
            factor = sqrt(factor);
        }
        out_kappa[a] = kappamid;
        out_Sigma[a] = S;
        out_UC[a] = UC;
        if (out_branch) out_branch[a] = branch;
    }
    free(M2d); free(w);
}

/* lakernel.py:397-443 : CG as coded (x0 = 0, atol = |b|*rtol, no final residual check).
 * A is the gathered (na, na) sub-matrix.  Returns the number of A@p products performed. */
static int orc_cg(const double *A, const double *b, long na, double rtol, long maxiter, double *x,
                  double *r, double *p, double *q) {
    double nb = 0.0;
    for (long i = 0; i < na; i++) nb += b[i] * b[i];
    double atol = sqrt(nb) * rtol;
    for (long i = 0; i < na; i++) { x[i] = 0.0; r[i] = b[i]; p[i] = b[i]; }
    double rho_prev = 0.0;
    int nprod = 0;
    for (long it = 0; it < maxiter; it++) {
        double rho = 0.0;
        for (long i = 0; i < na; i++) rho += r[i] * r[i];
        if (sqrt(rho) < atol) break;
        if (it > 0) {
            double beta = rho / rho_prev;
            for (long i = 0; i < na; i++) p[i] = p[i] * beta + r[i];
        }
        double pq = 0.0;
        for (long i = 0; i < na; i++) {
            const double *row = A + i * na;
            double s = 0.0;
            for (long j = 0; j < na; j++) s += row[j] * p[j];
            q[i] = s;
            pq += p[i] * s;
        }
        nprod++;
        double alpha = rho / pq;
        for (long i = 0; i < na; i++) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
        rho_prev = rho;
    }
    return nprod;
}

/* lakernel.py:548-590 : per output pixel gather the sub-system selected by relevant[a,:], CG it,
 * scatter into float32 Ti.  niter (m,) receives the per-pixel product count (P-discrete). */
void orc_iterative_wrapper(const double *AA, const double *mBhalf, const uint8_t *relevant, long m,
                           long n, double rtol, long maxiter, float *Ti, int32_t *niter,
                           int nthreads) {
#pragma omp parallel num_threads(nthreads) if (nthreads > 1)
    {
        long *sel = (long *)malloc(sizeof(long) * (n > 0 ? n : 1));
        double *As = (double *)malloc(sizeof(double) * (n > 0 ? n * n : 1));
        double *v = (double *)malloc(sizeof(double) * 5 * (n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 4)
        for (long a = 0; a < m; a++) {
            long na = 0;
            for (long i = 0; i < n; i++)
                if (relevant[a * n + i]) sel[na++] = i;
            for (long j = 0; j < na; j++)
                for (long i = 0; i <= j; i++) {
                    double val = AA[sel[j] * n + sel[i]];
                    As[j * na + i] = val;
                    As[i * na + j] = val;
                }
            double *b = v, *x = v + n, *r = v + 2 * n, *p = v + 3 * n, *q = v + 4 * n;
            for (long j = 0; j < na; j++) b[j] = mBhalf[a * n + sel[j]];
            int np_ = orc_cg(As, b, na, rtol, maxiter, x, r, p, q);
            if (niter) niter[a] = np_;
            for (long i = 0; i < n; i++) Ti[a * n + i] = 0.0f;
            for (long j = 0; j < na; j++) Ti[a * n + sel[j]] = (float)x[j];
        }
        free(sel); free(As); free(v);
    }
}
