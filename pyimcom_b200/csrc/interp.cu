// D5512 interpolation kernels: the furry_parakeet function seam (iD5512C, iD5512C_sym, gridD5512C)
// and the fused system-matrix assembly kernels that replace the reference's per-image-pair Python
// loops (psfutil.py:1401-1732, coadd.py:1028-1082).
//
// Arithmetic follows routine.py statement by statement (inner sum over x taps then outer over y
// taps) with fused multiply-adds; see the note at d5512_getw.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace b200 {

__device__ __constant__ double c_d5512_e[5][5] = {
    {+1.651881673372979740e-05, -3.145538007199505447e-04, +1.793518183780194427e-03, -2.904014557029917318e-03,
     +6.187591260980151433e-04},
    {-1.146756217210629335e-04, +2.883845374976550142e-03, -1.857047531896089884e-02, +3.147734488597204311e-02,
     -6.753293626461192439e-03},
    {+3.256838096371517067e-04, -9.702063770653997568e-03, +8.678848026470635524e-02, -1.659182651092198924e-01,
     +3.620560878249733799e-02},
    {-4.541830837949564726e-04, +1.494862093737218955e-02, -1.668775957435094937e-01, +5.879306056792649171e-01,
     -1.367845996704077915e-01},
    {+2.266560930061513573e-04, -7.815848920941316502e-03, +9.686607348538181506e-02, -4.505856722239036105e-01,
     +6.067135256905490381e-01}};
__device__ __constant__ double c_d5512_o[5][5] = {
    {-3.486978652054735998e-06, +6.753750285320532433e-05, -3.871378836550175566e-04, +6.279918076641771273e-04,
     -1.338434614116611838e-04},
    {+3.121412120355294799e-05, -8.040343683015897672e-04, +5.209574765466357636e-03, -8.847326408846412429e-03,
     +1.898674086370833597e-03},
    {-1.243658986204533102e-04, +3.804930695189636097e-03, -3.434861846914529643e-02, +6.581033749134083954e-02,
     -1.436476114189205733e-02},
    {+2.894406669584551734e-04, -9.794291009695265532e-03, +1.104231510875857830e-01, -3.906954914039130755e-01,
     +9.092432925988773451e-02},
    {-4.336085507644610966e-04, +1.537862263741893339e-02, -1.925091434770601628e-01, +8.993141455798455697e-01,
     -1.213035309579723942e+00}};

// ---- D5512 interpolation weights ---------------------------------------------------------------
// Reference: furry_parakeet iD5512C_getw == pyimcom/routine.py:29-122.  Ten weights from five even
// and five odd polynomials in fh = frac - 1/2, Horner order as in the reference.  The Horner steps
// and the tap sums below are fused multiply-adds: each differs from the reference's separately
// rounded multiply/add by at most one rounding (relative 1e-16 per step; parity tests state 1e-13).
// The sample COORDINATES are computed with the reference's exact operation sequence (no contraction)
// so that int(x), the on-grid test and the fractional offset are bit-identical to the CPU statement.

__device__ __forceinline__ void d5512_getw(double* __restrict__ w, double fh) {
    const double fh2 = fh * fh;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double e = c_d5512_e[k][0];
        double o = c_d5512_o[k][0];
#pragma unroll
        for (int q = 1; q < 5; q++) {
            e = fma(e, fh2, c_d5512_e[k][q]);
            o = fma(o, fh2, c_d5512_o[k][q]);
        }
        o *= fh;
        w[k] = e + o;
        w[9 - k] = e - o;
    }
}

// off-grid test of the reference (routine.py:166): the 10-tap window must fit the grid
__device__ __forceinline__ bool d5512_on_grid(int xi, int ngx) { return !(xi < 4 || xi >= ngx - 5); }

// 10x10 tap sum at integer corner (yi-4, xi-4) of grid g (row stride ngx).  flip=1 reads the grid
// mirrored in both axes (np.flip of the table, psfutil.py:1659-1665) without materialising it.
__device__ __forceinline__ double d5512_taps(const double* __restrict__ g, int ngy, int ngx, int yi, int xi,
                                             const double* wx, const double* wy, int flip) {
    double acc = 0.0;
    if (!flip) {
        const double* p = g + (size_t)(yi - 4) * ngx + (xi - 4);
#pragma unroll
        for (int i = 0; i < 10; i++) {
            double strip = 0.0;
#pragma unroll
            for (int j = 0; j < 10; j++) strip = fma(wx[j], __ldg(p + j), strip);
            acc = fma(strip, wy[i], acc);
            p += ngx;
        }
    } else {
        const double* p = g + (size_t)(ngy - 1 - (yi - 4)) * ngx + (ngx - 1 - (xi - 4));
#pragma unroll
        for (int i = 0; i < 10; i++) {
            double strip = 0.0;
#pragma unroll
            for (int j = 0; j < 10; j++) strip = fma(wx[j], __ldg(p - j), strip);
            acc = fma(strip, wy[i], acc);
            p -= ngx;
        }
    }
    return acc;
}

// ---- polyphase table layout (A assembly) -------------------------------------------------------------------------
// The native pixel pitch of the input images is exactly P = oversamp table samples (dscale = pitch / P,
// psfutil.py:610), so the 32 consecutive input pixels a warp handles read table positions that are ~P samples
// apart: in the row-major table every lane touches its own cache line.  The assembly kernel therefore reads the
// PSF-overlap tables in a polyphase layout  T'[y mod P][x mod P][y / P][x / P]  (planes of ncell x ncell doubles,
// ncell = ceil(ngrid / P)): lanes whose positions differ by P samples now read neighbouring doubles of one plane.
// The re-layout is done once per table when the arena is built; values and the order of the arithmetic are
// untouched, so the result is bit-identical to the row-major path.
// base + off (bytes, < 4 GiB) as ONE instruction (IMAD.WIDE.U32) instead of the four-instruction 64-bit index
// arithmetic the compiler emits for  tables[offset64 + oy + ox]
__device__ __forceinline__ const double* addr_u32(const void* base, unsigned off) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off), "l"(base));
    return reinterpret_cast<const double*>(r);
}

template <int P>
struct PolyOff {
    unsigned ox[10], oy[10];  // byte offsets inside one table (a polyphase table is < 4 GiB)
    // integer corner (yi-4, xi-4); flip mirrors both axes (np.flip of the table, psfutil.py:1659-1665)
    __device__ __forceinline__ PolyOff(int yi, int xi, int ngrid, int ncell, int flip) {
        const unsigned plane = (unsigned)(ncell * ncell);
#pragma unroll
        for (int j = 0; j < 10; j++) {
            const unsigned x = (unsigned)(flip ? ngrid - 1 - (xi - 4 + j) : xi - 4 + j);
            const unsigned y = (unsigned)(flip ? ngrid - 1 - (yi - 4 + j) : yi - 4 + j);
            ox[j] = ((x % P) * plane + x / P) * 8u;
            oy[j] = ((y % P) * (P * plane) + (y / P) * (unsigned)ncell) * 8u;
        }
    }
};

template <int P>
__device__ __forceinline__ double d5512_taps_poly(const double* __restrict__ g, const PolyOff<P>& o, const double* wx,
                                                  const double* wy) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const double* p = addr_u32(g, o.oy[i]);
        double strip = 0.0;
#pragma unroll
        for (int j = 0; j < 10; j++) strip = fma(wx[j], __ldg(addr_u32(p, o.ox[j])), strip);
        acc = fma(strip, wy[i], acc);
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// iD5512C: routine.py:125-181.  One thread per scattered point, all layers.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_iD5512C(const double* __restrict__ f, int nlayer, int ngy, int ngx,
                                                 const double* __restrict__ xpos, const double* __restrict__ ypos,
                                                 long nout, double* __restrict__ out) {
    long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= nout) return;
    const double x = xpos[ip], y = ypos[ip];
    const int xi = (int)x, yi = (int)y;
    if (!d5512_on_grid(xi, ngx) || !d5512_on_grid(yi, ngy)) return;  // output untouched
    double wx[10], wy[10];
    d5512_getw(wx, x - xi - 0.5);
    d5512_getw(wy, y - yi - 0.5);
    for (int l = 0; l < nlayer; l++)
        out[(size_t)l * nout + ip] = d5512_taps(f + (size_t)l * ngy * ngx, ngy, ngx, yi, xi, wx, wy, 0);
}

// iD5512C_sym: routine.py:184-253.  Upper triangle of a sq x sq point matrix, mirrored.
__global__ void __launch_bounds__(128) k_iD5512C_sym(const double* __restrict__ f, int nlayer, int ngy, int ngx,
                                                     const double* __restrict__ xpos, const double* __restrict__ ypos,
                                                     long nout, int sq, double* __restrict__ out) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)sq * sq) return;
    const int i1 = (int)(t / sq), i2 = (int)(t % sq);
    if (i2 < i1) return;
    const long ip = (long)i1 * sq + i2, ipm = (long)i2 * sq + i1;
    const double x = xpos[ip], y = ypos[ip];
    const int xi = (int)x, yi = (int)y;
    if (!d5512_on_grid(xi, ngx) || !d5512_on_grid(yi, ngy)) {
        // the reference leaves the upper entry untouched and still copies it to the lower triangle
        if (i1 != i2)
            for (int l = 0; l < nlayer; l++) out[(size_t)l * nout + ipm] = out[(size_t)l * nout + ip];
        return;
    }
    double wx[10], wy[10];
    d5512_getw(wx, x - xi - 0.5);
    d5512_getw(wy, y - yi - 0.5);
    for (int l = 0; l < nlayer; l++) {
        double v = d5512_taps(f + (size_t)l * ngy * ngx, ngy, ngx, yi, xi, wx, wy, 0);
        out[(size_t)l * nout + ip] = v;
        out[(size_t)l * nout + ipm] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// gridD5512C: routine.py:256-338.  One CTA per input pixel: weights for the nxo columns and nyo rows
// are computed once into shared memory (off-grid => zero weights, index 4), then the nyo*nxo outputs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gridD5512C(const double* __restrict__ f, int ngy, int ngx,
                                                    const double* __restrict__ xpos, const double* __restrict__ ypos,
                                                    int nxo, int nyo, double* __restrict__ out) {
    extern __shared__ double sm[];
    double* wx = sm;                    // [nxo][10]
    double* wy = sm + 10 * (size_t)nxo;  // [nyo][10]
    int* xi = (int*)(wy + 10 * (size_t)nyo);
    int* yi = xi + nxo;
    const long p = blockIdx.x;
    for (int t = threadIdx.x; t < nxo + nyo; t += blockDim.x) {
        const bool isx = t < nxo;
        const int k = isx ? t : t - nxo;
        const double v = isx ? xpos[p * nxo + k] : ypos[p * nyo + k];
        int vi = (int)v;
        double w[10];
        if (!d5512_on_grid(vi, isx ? ngx : ngy)) {
            vi = 4;
#pragma unroll
            for (int q = 0; q < 10; q++) w[q] = 0.0;
        } else {
            d5512_getw(w, v - vi - 0.5);
        }
        double* dst = (isx ? wx : wy) + 10 * (size_t)k;
#pragma unroll
        for (int q = 0; q < 10; q++) dst[q] = w[q];
        (isx ? xi : yi)[k] = vi;
    }
    __syncthreads();
    const int npt = nxo * nyo;
    for (int t = threadIdx.x; t < npt; t += blockDim.x) {
        const int iy = t / nxo, ix = t - iy * nxo;
        out[p * npt + t] = d5512_taps(f, ngy, ngx, yi[iy], xi[ix], wx + 10 * ix, wy + 10 * iy, 0);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused A assembly for one output stamp (stage a; replaces psfutil.py:1401-1495, 1597-1732 and the
// 9+36 block scatter of coadd.py:1028-1069).
//
// The n selected input pixels of the 3x3 InStamp neighbourhood are given in the reference's
// concatenation order with, per pixel, position (x,y) in output-pixel units and a dense code
// (local PSF-group id * nimg + image id).  Entry (i,j), i<=j, is the D5512 interpolation of the
// PSF-overlap table of (group_i,image_i ; group_j,image_j) at ((p_i - p_j)/dscale + nc + 6) -- the
// reference evaluates exactly these upper-triangle entries (same stamp & image: iD5512C_sym;
// same stamp, image j<i: block (j,i); stamp a<b: block (a,b)) and mirrors them; so do we.
// Tables are stored zero-padded by 6 (np.pad(ovl, 6), psfutil.py:1471, 1696).
// lut[(ci * ncode + cj)] = {table offset (doubles) into `tables`, flip, flat-penalty subtrahend}.
// Rows/columns n..npad-1 are filled with the identity so the padded matrix stays SPD.
// One CTA per upper-triangular 32x32 tile; the mirrored tile is written through shared memory so
// both stores are row-contiguous.
// ------------------------------------------------------------------------------------------------
template <int P>  // P == 0: row-major tables; P > 0: polyphase tables of period P
__global__ void __launch_bounds__(256) k_build_A(const double* __restrict__ px, const double* __restrict__ py,
                                                 const int* __restrict__ pcode, int n, int npad,
                                                 const double* __restrict__ tables, const TableRef* __restrict__ lut,
                                                 int nimg, int ncode, int ngrid, double dscale, double nc,
                                                 double flat_penalty, double* __restrict__ A, int lda,
                                                 double diag_add) {
    __shared__ double tile[32][33];
    const int nt = npad / 32;
    int bi, rem = blockIdx.x;
    {
        // invert the triangular tile numbering: row bi holds (nt - bi) tiles
        const double b = 2.0 * nt + 1.0;
        int guess = (int)((b - sqrt(b * b - 8.0 * rem)) * 0.5);
        if (guess < 0) guess = 0;
        if (guess > nt - 1) guess = nt - 1;
        while (guess > 0 && (long)guess * (2 * nt - guess + 1) / 2 > rem) guess--;
        while ((long)(guess + 1) * (2 * nt - guess) / 2 <= rem) guess++;
        bi = guess;
        rem -= (int)((long)bi * (2 * nt - bi + 1) / 2);
    }
    const int bj = bi + rem;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int j = bj * 32 + tx;
    double xj = 0, yj = 0;
    int cj = 0;
    if (j < n) {
        xj = px[j];
        yj = py[j];
        cj = pcode[j];
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int li = ty + 8 * r;
        const int i = bi * 32 + li;
        double v = 0.0;
        if (i < n && j < n) {
            if (i <= j) {
                const int ci = pcode[i];
                // dd = (p_i - p_j); dd /= dscale; dd += nc; (+6 at the call) -- psfutil.py:1423-1428, 1474
                const double x = __dadd_rn(__dadd_rn(__ddiv_rn(__dadd_rn(px[i], -xj), dscale), nc), 6.0);
                const double y = __dadd_rn(__dadd_rn(__ddiv_rn(__dadd_rn(py[i], -yj), dscale), nc), 6.0);
                const TableRef tr = lut[(size_t)ci * ncode + cj];
                const int xi = (int)x, yi = (int)y;
                if (tr.offset >= 0 && d5512_on_grid(xi, ngrid) && d5512_on_grid(yi, ngrid)) {
                    double wx[10], wy[10];
                    d5512_getw(wx, x - xi - 0.5);
                    d5512_getw(wy, y - yi - 0.5);
                    if (P > 0) {
                        const PolyOff<(P > 0 ? P : 1)> po(yi, xi, ngrid, (ngrid + P - 1) / (P > 0 ? P : 1), tr.flip);
                        v = d5512_taps_poly(tables + tr.offset, po, wx, wy);
                    } else {
                        v = d5512_taps(tables + tr.offset, ngrid, ngrid, yi, xi, wx, wy, tr.flip);
                    }
                }
                if (flat_penalty != 0.0) {  // psfutil.py:1483-1486, 1705-1708
                    v = __dadd_rn(v, -tr.penalty_sub);
                    if (ci % nimg == cj % nimg) v = __dadd_rn(v, flat_penalty);
                }
                if (i == j) v += diag_add;
                A[(size_t)i * lda + j] = v;
            }
        } else if (i < npad && j < npad) {
            v = (i == j) ? 1.0 : 0.0;
            if (i <= j) A[(size_t)i * lda + j] = v;
        }
        tile[li][tx] = v;
    }
    __syncthreads();
    // mirror: A[j][i] = A[i][j] for i<j, written row-contiguously through the shared tile
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int lj = ty + 8 * r;  // local column index of the tile -> output row
        const int jj = bj * 32 + lj, ii = bi * 32 + tx;
        if (jj < npad && ii < npad && ii < jj) A[(size_t)jj * lda + ii] = tile[tx][lj];
    }
}

// ------------------------------------------------------------------------------------------------
// InStamp-pair blocks of A and their cache (SysMatA, psfutil.py:1764-2092).
//
// Neighbouring output stamps share six of their nine InStamps, so most of the 9 + 36 InStamp-pair blocks of a
// stamp's A recur in its neighbours: the reference keeps them in a reference-counted cache keyed
// (ji_st1, ji_st2) and interpolates each only once (13 distinct blocks per InStamp instead of 45 per OutStamp).
// k_pair_blocks evaluates a whole list of FULL pair blocks (all pixels of InStamp a x all pixels of InStamp b,
// a <= b in raster order) in one launch; k_assemble_A then cuts one OutStamp's A out of the cached blocks
// through its pixel selections.  Entry values, table choice, flips and the upper-triangle-then-mirror rule inside
// a self block are those of k_build_A, so both routes give bit-identical matrices.
// ------------------------------------------------------------------------------------------------
// One 32 x 32 tile of the pair-block list, by NT threads (NT / 32 warps, 1024 / NT entries per thread).  Returns true
// when the shared tile was read back (self block: the caller has to synchronise before the tile is written again).
template <int P, int NT>
__device__ __forceinline__ bool pair_tile(int tile_idx, double (*tile)[33], const double* __restrict__ gx,
                                          const double* __restrict__ gy, const int* __restrict__ gimg,
                                          const PairDesc* __restrict__ descs, const int* __restrict__ tile_prefix,
                                          int npair, const double* __restrict__ tables,
                                          const TableRef* __restrict__ lut, int nimg, int ngrid, double dscale,
                                          double nc, double flat_penalty, double* __restrict__ pool) {
    constexpr int NW = NT / 32, NR = 1024 / NT;
    // which pair does this tile belong to: largest q with tile_prefix[q] <= tile_idx
    int lo = 0, hi = npair - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_prefix[mid] <= tile_idx)
            lo = mid;
        else
            hi = mid - 1;
    }
    const PairDesc d = descs[lo];
    const int t = tile_idx - tile_prefix[lo];
    const int ntj = (d.nB + 31) >> 5;
    const int bi = t / ntj, bj = t - bi * ntj;
    if (d.same && bj < bi) return false;  // self block: upper tiles only, mirrored below
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // A warp covers a 4 x 8 patch of the tile (4 consecutive pixels i, 8 consecutive pixels j), not one row of 32:
    // consecutive pixels of one image row are P table samples apart, so entry (i+1, j+1) reads (almost) the window of
    // entry (i, j) -- the Toeplitz structure of the block -- and the 32 lanes of a load touch ~11 neighbouring doubles
    // of one polyphase plane (1-2 cache lines) instead of 32 doubles spread over the 3-4 image rows a 32-pixel run spans.
    const int pli = tx >> 3, plj = tx & 7;
    double* blk = pool + d.out;
    const TableRef* plut = lut + (size_t)d.lut * nimg * nimg;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const int sub = ty + NW * r;            // 32 patches per tile: 8 patch rows x 4 patch columns
        const int li = 4 * (sub >> 2) + pli;    // local row 0..31
        const int lj = 8 * (sub & 3) + plj;     // local column 0..31
        const int i = bi * 32 + li, j = bj * 32 + lj;
        double v = 0.0;
        if (i < d.nA && j < d.nB && (!d.same || i <= j)) {
            const int ci = gimg[d.offA + i], cj = gimg[d.offB + j];
            const double x = __dadd_rn(__dadd_rn(__ddiv_rn(__dadd_rn(gx[d.offA + i], -gx[d.offB + j]), dscale), nc), 6.0);
            const double y = __dadd_rn(__dadd_rn(__ddiv_rn(__dadd_rn(gy[d.offA + i], -gy[d.offB + j]), dscale), nc), 6.0);
            const TableRef tr = plut[ci * nimg + cj];
            const int xi = (int)x, yi = (int)y;
            if (tr.offset >= 0 && d5512_on_grid(xi, ngrid) && d5512_on_grid(yi, ngrid)) {
                double wx[10], wy[10];
                d5512_getw(wx, x - xi - 0.5);
                d5512_getw(wy, y - yi - 0.5);
                if (P > 0) {
                    const PolyOff<(P > 0 ? P : 1)> po(yi, xi, ngrid, (ngrid + P - 1) / (P > 0 ? P : 1), tr.flip);
                    v = d5512_taps_poly(tables + tr.offset, po, wx, wy);
                } else {
                    v = d5512_taps(tables + tr.offset, ngrid, ngrid, yi, xi, wx, wy, tr.flip);
                }
            }
            if (flat_penalty != 0.0) {  // psfutil.py:1483-1486, 1705-1708
                v = __dadd_rn(v, -tr.penalty_sub);
                if (ci == cj) v = __dadd_rn(v, flat_penalty);
            }
            __stcs(blk + (size_t)i * d.ld + j, v);  // streaming store: the blocks must not evict the tables from L2
        }
        tile[li][lj] = v;
    }
    if (!d.same) return false;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const int lj = ty + NW * r;
        const int jj = bj * 32 + lj, ii = bi * 32 + tx;
        if (jj < d.nB && ii < d.nA && ii < jj) __stcs(blk + (size_t)jj * d.ld + ii, tile[tx][lj]);
    }
    return true;
}

// (B200_PAIR_MINB: resident CTAs per SM the register allocation aims at; 3 = 80 registers, 4 = 64 with ~150 B of spills)
#ifndef B200_PAIR_MINB
#define B200_PAIR_MINB 3
#endif
template <int P>
__global__ void __launch_bounds__(256, B200_PAIR_MINB) k_pair_blocks(const double* __restrict__ gx, const double* __restrict__ gy,
                                                     const int* __restrict__ gimg, const PairDesc* __restrict__ descs,
                                                     const int* __restrict__ tile_prefix, int npair,
                                                     const double* __restrict__ tables,
                                                     const TableRef* __restrict__ lut, int nimg, int ngrid,
                                                     double dscale, double nc, double flat_penalty,
                                                     double* __restrict__ pool) {
    __shared__ double tile[32][33];
    pair_tile<P, 256>((int)blockIdx.x, tile, gx, gy, gimg, descs, tile_prefix, npair, tables, lut, nimg, ngrid, dscale, nc,
                      flat_penalty, pool);
}

// The same tiles by a SMALL resident grid (B200_PAIR_RESIDENT = CTAs of 128 threads per SM, each walking the tile list
// with the grid's stride): an EXPERIMENT, off by default (tools/coresident_check.py).  The block
// scheduler hands out the CTAs of one grid before it turns to the next, so a grid of thousands of tiles never shares an
// SM with another stream's kernel; a grid that is resident at once does, and one such CTA (88 registers x 128 threads,
// 8 KB of shared memory) fits beside a k_oz_gemm CTA.  Measured: it does run there, but makes 7 % of its standalone
// progress while slowing the GEMM by 18 %, and with 4 warps per SM it is latency-bound (54 ms per launch against
// 10.5 ms for one CTA per tile).  Same entries, same arithmetic: bit-identical blocks.
template <int P>
__global__ void __maxnreg__(88) k_pair_blocks_resident(const double* __restrict__ gx, const double* __restrict__ gy,
                                                              const int* __restrict__ gimg,
                                                              const PairDesc* __restrict__ descs,
                                                              const int* __restrict__ tile_prefix, int npair, int ntiles,
                                                              const double* __restrict__ tables,
                                                              const TableRef* __restrict__ lut, int nimg, int ngrid,
                                                              double dscale, double nc, double flat_penalty,
                                                              double* __restrict__ pool) {
    __shared__ double tile[32][33];
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
        if (pair_tile<P, 128>(t, tile, gx, gy, gimg, descs, tile_prefix, npair, tables, lut, nimg, ngrid, dscale, nc,
                              flat_penalty, pool))
            __syncthreads();
}

// One OutStamp's A (npad x lda) from cached pair blocks.  Stamp pixel k (0 <= k < n) is pixel gidx[k] of the block's
// global pixel list and belongs to segment s (one of the 9 InStamps, seg_start[s] <= k < seg_start[s+1]); its index
// inside that InStamp is gidx[k] - inst_off[s].  Upper-triangle tiles are read row-contiguously from block
// (s_i, s_j), s_i <= s_j, and the mirrored tile is written through shared memory; rows/columns n..npad-1 carry the
// identity; diag_add goes onto the first n diagonal entries.
__global__ void __launch_bounds__(256) k_assemble_A(AsmDesc d, const int* __restrict__ gidx, int n, int npad,
                                                    const double* __restrict__ pool, double* __restrict__ A, int lda,
                                                    double diag_add) {
    __shared__ double tile[32][33];
    const int nt = npad / 32;
    int bi, rem = blockIdx.x;
    {
        const double b = 2.0 * nt + 1.0;
        int guess = (int)((b - sqrt(b * b - 8.0 * rem)) * 0.5);
        if (guess < 0) guess = 0;
        if (guess > nt - 1) guess = nt - 1;
        while (guess > 0 && (long)guess * (2 * nt - guess + 1) / 2 > rem) guess--;
        while ((long)(guess + 1) * (2 * nt - guess) / 2 <= rem) guess++;
        bi = guess;
        rem -= (int)((long)bi * (2 * nt - bi + 1) / 2);
    }
    const int bj = bi + rem;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = bj * 32 + tx;
    int sj = 0, lj = 0;
    if (j < n) {
        while (sj < 8 && j >= d.seg_start[sj + 1]) sj++;
        lj = gidx[j] - d.inst_off[sj];
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int li = ty + 8 * r;
        const int i = bi * 32 + li;
        double v = 0.0;
        if (i < n && j < n) {
            if (i <= j) {
                int si = 0;
                while (si < 8 && i >= d.seg_start[si + 1]) si++;
                const int ii = gidx[i] - d.inst_off[si];
                const int q = si * 9 + sj;  // si <= sj because segments are consecutive runs of the stamp order
                v = pool[d.blk[q] + (size_t)ii * d.ld[q] + lj];
                if (i == j) v += diag_add;
                A[(size_t)i * lda + j] = v;
            }
        } else if (i < npad && j < npad) {
            v = (i == j) ? 1.0 : 0.0;
            if (i <= j) A[(size_t)i * lda + j] = v;
        }
        tile[li][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int lj2 = ty + 8 * r;
        const int jj = bj * 32 + lj2, ii = bi * 32 + tx;
        if (jj < npad && ii < npad && ii < jj) A[(size_t)jj * lda + ii] = tile[tx][lj2];
    }
}

// ------------------------------------------------------------------------------------------------
// Fused mBhalf assembly for one output stamp (stage a; replaces psfutil.py:1497-1595 and
// coadd.py:1075-1082).  mBhalf[o][a=(iy,ix)][i] = gridD5512C of io-table(group_i,image_i,o) at
// ((x_i - xout[ix])/dscale + nc + 6, (y_i - yout[iy])/dscale + nc + 6).
// One CTA per tile of TI input pixels: the x/y weight sets of every pixel of the tile are computed
// once into shared memory, then threads sweep (a, i) with i fastest so that stores are contiguous.
// Rows m..mpad-1 and columns n..npad-1 are zero-filled (the solver's padding).
// ------------------------------------------------------------------------------------------------
template <int TI>
__global__ void __launch_bounds__(256) k_build_B(const double* __restrict__ px, const double* __restrict__ py,
                                                 const int* __restrict__ pcode, int n, int npad,
                                                 const double* __restrict__ tables,
                                                 const long long* __restrict__ lut_io /* [ncode][n_out] */, int n_out,
                                                 int ngrid, double dscale, double nc, int n2f, int mpad, double x0out,
                                                 double y0out, double* __restrict__ B, int ldb, size_t strideB) {
    extern __shared__ double sm[];
    double* wx = sm;                               // [TI][n2f][10]
    double* wy = wx + (size_t)TI * n2f * 10;       // [TI][n2f][10]
    int* xi = (int*)(wy + (size_t)TI * n2f * 10);  // [TI][n2f]
    int* yi = xi + TI * n2f;
    const int i0 = blockIdx.x * TI;
    for (int t = threadIdx.x; t < TI * n2f * 2; t += blockDim.x) {
        const int isy = t / (TI * n2f);
        const int u = t - isy * TI * n2f;
        const int li = u / n2f, k = u - li * n2f;
        const int i = i0 + li;
        double w[10];
        int vi = 4;
        bool ok = false;
        if (i < n) {
            const double pin = isy ? py[i] : px[i];
            const double pout = (isy ? y0out : x0out) + (double)k;  // integer output grid (coadd.py:879-882)
            const double v = __dadd_rn(__dadd_rn(__ddiv_rn(__dadd_rn(pin, -pout), dscale), nc), 6.0);
            vi = (int)v;
            if (d5512_on_grid(vi, ngrid)) {
                d5512_getw(w, v - vi - 0.5);
                ok = true;
            } else {
                vi = 4;
            }
        }
        double* dst = (isy ? wy : wx) + (size_t)u * 10;
#pragma unroll
        for (int q = 0; q < 10; q++) dst[q] = ok ? w[q] : 0.0;
        (isy ? yi : xi)[u] = vi;
    }
    __syncthreads();
    const int m = n2f * n2f;
    for (int o = 0; o < n_out; o++) {
        for (int t = threadIdx.x; t < mpad * TI; t += blockDim.x) {
            const int a = t / TI, li = t - a * TI;
            const int i = i0 + li;
            if (i >= npad) continue;
            double v = 0.0;
            if (i < n && a < m) {
                const int iy = a / n2f, ix = a - iy * n2f;
                const long long off = lut_io[(size_t)pcode[i] * n_out + o];
                if (off >= 0)
                    v = d5512_taps(tables + off, ngrid, ngrid, yi[li * n2f + iy], xi[li * n2f + ix],
                                   wx + (size_t)(li * n2f + ix) * 10, wy + (size_t)(li * n2f + iy) * 10, 0);
            }
            B[o * strideB + (size_t)a * ldb + i] = v;
        }
    }
}

// Separable form of the same assembly (the production kernel).  gridD5512C's sum is
//   out[iy][ix] = sum_i wy[iy][i] * ( sum_j wx[ix][j] f[yi[iy]-4+i][xi[ix]-4+j] )        (routine.py:313-336)
// and the inner strip depends on the table row and on ix only.  Output pixels are 1/dscale < 3 table samples
// apart, so the 10-row windows of consecutive iy overlap heavily: per input pixel the CTA first forms the strips
// S[r][ix] for every table row r its output column needs (phase 1: ~(n2f/dscale + 10) * n2f strips, global
// gathers with lanes along ix), then every output is ten fused multiply-adds on S (phase 2, shared memory only).
// Both phases keep the reference's operation order (strip: j = 0..9, then i = 0..9), so the result is bit-identical
// to the direct form above while doing ~2.5x fewer loads and multiply-adds, none of them redundant gathers.
// TI input pixels per CTA so that each output row receives TI contiguous doubles (a full 32-byte sector for TI=4).
template <int TI>
__global__ void __launch_bounds__(512) k_build_B_sep(const double* __restrict__ px, const double* __restrict__ py,
                                                     const int* __restrict__ pcode, int n, int npad,
                                                     const double* __restrict__ tables,
                                                     const long long* __restrict__ lut_io, int n_out, int ngrid,
                                                     double dscale, double nc, int n2f, int mpad, double x0out,
                                                     double y0out, double* __restrict__ B, int ldb, size_t strideB,
                                                     int rmax) {
    extern __shared__ double sm[];
    double* wx = sm;                                // [TI][n2f][10]
    double* wy = wx + (size_t)TI * n2f * 10;        // [TI][n2f][10]
    double* S = wy + (size_t)TI * n2f * 10;         // [TI][rmax][n2f]
    int* xi = (int*)(S + (size_t)TI * rmax * n2f);  // [TI][n2f], -1 = off grid
    int* yi = xi + TI * n2f;
    int* r0 = yi + TI * n2f;  // [TI] first table row of the strip cache
    int* nr = r0 + TI;        // [TI] rows held (0: nothing to do, > rmax: direct evaluation)
    const int i0 = blockIdx.x * TI;
    for (int t = threadIdx.x; t < TI * n2f * 2; t += blockDim.x) {
        const int isy = t / (TI * n2f);
        const int u = t - isy * TI * n2f;
        const int li = u / n2f, k = u - li * n2f;
        const int i = i0 + li;
        double w[10];
        int vi = -1;
        if (i < n) {
            const double pin = isy ? py[i] : px[i];
            const double pout = (isy ? y0out : x0out) + (double)k;  // integer output grid (coadd.py:879-882)
            const double v = __dadd_rn(__dadd_rn(__ddiv_rn(__dadd_rn(pin, -pout), dscale), nc), 6.0);
            vi = (int)v;
            if (d5512_on_grid(vi, ngrid))
                d5512_getw(w, v - vi - 0.5);
            else
                vi = -1;
        }
        double* dst = (isy ? wy : wx) + (size_t)u * 10;
#pragma unroll
        for (int q = 0; q < 10; q++) dst[q] = vi >= 0 ? w[q] : 0.0;
        (isy ? yi : xi)[u] = vi;
    }
    __syncthreads();
    if (threadIdx.x < TI) {
        const int li = threadIdx.x;
        int lo = 1 << 30, hi = -1;
        for (int k = 0; k < n2f; k++) {
            const int v = yi[li * n2f + k];
            if (v >= 0) {
                lo = min(lo, v - 4);
                hi = max(hi, v + 5);
            }
        }
        r0[li] = lo;
        nr[li] = hi >= 0 ? hi - lo + 1 : 0;
    }
    __syncthreads();
    const int m = n2f * n2f;
    for (int o = 0; o < n_out; o++) {
        // phase 1: strips.  A thread owns one (input pixel, output column) pair -- its ten x-weights stay in registers
        // -- and walks down every RS-th table row; the ten taps of a row come from five (even start) or six aligned
        // 16-byte loads.  The stage is bound by L1 wavefronts (a warp's 32 columns spread over ~7 cache lines per load
        // instruction): 6 vector loads per strip entry instead of 10 scalar ones plus 10 shared-memory weight reads.
        // Same ten FMAs in the same order as before, so the values are unchanged.
        const int npair = TI * n2f;
        const int RS = max(1, (int)blockDim.x / npair);
        for (int it = threadIdx.x; it < npair * RS; it += blockDim.x) {
            const int rs = it / npair;
            const int u = it - rs * npair;  // li * n2f + ix: a warp holds consecutive columns of one pixel
            const int li = u / n2f, ix = u - li * n2f;
            const int i = i0 + li;
            const int nrl = nr[li];
            if (i < n && nrl > 0 && nrl <= rmax) {
                const int x = xi[u];
                const long long off = lut_io[(size_t)pcode[i] * n_out + o];
                double* Sp = S + (size_t)li * rmax * n2f + ix;
                if (x >= 0 && off >= 0) {
                    double w[10];
#pragma unroll
                    for (int j = 0; j < 10; j++) w[j] = wx[(size_t)u * 10 + j];
                    long long e = off + (long long)(r0[li] + rs) * ngrid + (x - 4);
#pragma unroll 2
                    for (int r = rs; r < nrl; r += RS, e += (long long)RS * ngrid) {
                        const bool odd = e & 1;
                        const double2* p = reinterpret_cast<const double2*>(tables + (e - (odd ? 1 : 0)));
                        const double2 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3),
                                      v4 = __ldg(p + 4);
                        double2 v5 = make_double2(0.0, 0.0);
                        if (odd) v5 = __ldg(p + 5);
                        double strip = 0.0;
                        strip = fma(w[0], odd ? v0.y : v0.x, strip);
                        strip = fma(w[1], odd ? v1.x : v0.y, strip);
                        strip = fma(w[2], odd ? v1.y : v1.x, strip);
                        strip = fma(w[3], odd ? v2.x : v1.y, strip);
                        strip = fma(w[4], odd ? v2.y : v2.x, strip);
                        strip = fma(w[5], odd ? v3.x : v2.y, strip);
                        strip = fma(w[6], odd ? v3.y : v3.x, strip);
                        strip = fma(w[7], odd ? v4.x : v3.y, strip);
                        strip = fma(w[8], odd ? v4.y : v4.x, strip);
                        strip = fma(w[9], odd ? v5.x : v4.y, strip);
                        Sp[(size_t)r * n2f] = strip;
                    }
                } else {
                    for (int r = rs; r < nrl; r += RS) Sp[(size_t)r * n2f] = 0.0;
                }
            }
        }
        __syncthreads();
        // phase 2: outputs, TI contiguous doubles per output row
        for (int t = threadIdx.x; t < mpad * TI; t += blockDim.x) {
            const int a = t / TI, li = t - a * TI;
            const int i = i0 + li;
            if (i >= npad) continue;
            double v = 0.0;
            if (i < n && a < m) {
                const int iy = a / n2f, ix = a - iy * n2f;
                const int y = yi[li * n2f + iy], x = xi[li * n2f + ix];
                if (y >= 0 && x >= 0) {
                    if (nr[li] <= rmax) {
                        const double* sp = S + ((size_t)li * rmax + (y - 4 - r0[li])) * n2f + ix;
                        const double* w = wy + (size_t)(li * n2f + iy) * 10;
#pragma unroll
                        for (int q = 0; q < 10; q++) v = fma(sp[q * n2f], w[q], v);
                    } else {
                        const long long off = lut_io[(size_t)pcode[i] * n_out + o];
                        if (off >= 0)
                            v = d5512_taps(tables + off, ngrid, ngrid, y, x, wx + (size_t)(li * n2f + ix) * 10,
                                           wy + (size_t)(li * n2f + iy) * 10, 0);
                    }
                }
            }
            B[o * strideB + (size_t)a * ldb + i] = v;
        }
        __syncthreads();
    }
}

// Gather of the selected input pixels of one output stamp (coadd.py:969-977): positions, table codes and
// the n_inframe float32 layers, in the reference's concatenation order given by idx.  Columns n..npad-1
// of the layer block are zero-filled (they multiply zero columns of T).
__global__ void __launch_bounds__(256) k_gather_stamp(const int* __restrict__ idx, int n, int npad,
                                                      const double* __restrict__ sx, const double* __restrict__ sy,
                                                      const int* __restrict__ scode, const float* __restrict__ sdata,
                                                      long src_ld, int n_inframe, double* __restrict__ px,
                                                      double* __restrict__ py, int* __restrict__ pcode,
                                                      float* __restrict__ indata, int ldi) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npad) return;
    if (k < n) {
        const int g = idx[k];
        px[k] = sx[g];
        py[k] = sy[g];
        if (scode) pcode[k] = scode[g];
        for (int f = 0; f < n_inframe; f++) indata[(size_t)f * ldi + k] = sdata[(size_t)f * src_ld + g];
    } else {
        for (int f = 0; f < n_inframe; f++) indata[(size_t)f * ldi + k] = 0.0f;
    }
}

// PSF-overlap tables as uploaded (ntab, ns, ns) -> the arena layout: zero-padded by `pad` on every side
// (np.pad(ovl, 6), psfutil.py:1471, 1580, 1696) to ngrid = ns + 2 pad, row-major (P == 0) or polyphase of period P.
__global__ void __launch_bounds__(256) k_layout_tables(const double* __restrict__ src, int ntab, int ns, int pad,
                                                       int ngrid, int P, double* __restrict__ dst) {
    const int ncell = P > 0 ? (ngrid + P - 1) / P : 0;
    const long long per = P > 0 ? (long long)P * P * ncell * ncell : (long long)ngrid * ngrid;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per * ntab) return;
    const int t = (int)(e / per);
    long long r = e - (long long)t * per;
    int y, x;
    if (P > 0) {
        const int cx = (int)(r % ncell);
        r /= ncell;
        const int cy = (int)(r % ncell);
        r /= ncell;
        const int pxx = (int)(r % P), pyy = (int)(r / P);
        y = cy * P + pyy;
        x = cx * P + pxx;
    } else {
        y = (int)(r / ngrid);
        x = (int)(r - (long long)y * ngrid);
    }
    const int sy = y - pad, sx = x - pad;
    dst[e] = (sy >= 0 && sy < ns && sx >= 0 && sx < ns) ? src[((size_t)t * ns + sy) * ns + sx] : 0.0;
}

// PSF-overlap spectra: G = F1 * conj(F2) element by element on split real / imaginary planes.
// F2 may be a single spectrum broadcast over the n1 spectra of F1 (stride2 == 0) or a matching stack.
__global__ void __launch_bounds__(256) k_cmul_conj(const double* __restrict__ ar, const double* __restrict__ ai,
                                                   const double* __restrict__ br, const double* __restrict__ bi,
                                                   long long per, long long stride2, long long total, double im_sign,
                                                   double* __restrict__ gr, double* __restrict__ gi) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long t = e / per, r = e - t * per;
    const double xr = ar[e], xi = ai[e];
    const double yr = br[t * stride2 + r], yi = bi[t * stride2 + r];
    gr[e] = fma(xr, yr, xi * yi);
    gi[e] = im_sign * fma(xi, yr, -(xr * yi));
}

__global__ void k_getw(double* __restrict__ w, double fh) {
    double t[10];
    d5512_getw(t, fh);
    for (int k = 0; k < 10; k++) w[k] = t[k];
}

// ---- launchers ---------------------------------------------------------------------------------
int launch_iD5512C(const double* f, int nlayer, int ngy, int ngx, const double* x, const double* y, long nout,
                   double* out, cudaStream_t s) {
    if (nout <= 0) return 0;
    k_iD5512C<<<(unsigned)((nout + 127) / 128), 128, 0, s>>>(f, nlayer, ngy, ngx, x, y, nout, out);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_iD5512C_sym(const double* f, int nlayer, int ngy, int ngx, const double* x, const double* y, long nout,
                       double* out, cudaStream_t s) {
    if (nout <= 0) return 0;
    const int sq = (int)sqrt((double)(nout + 1));  // routine.py:214
    const long nt = (long)sq * sq;
    if (nt <= 0) return 0;
    k_iD5512C_sym<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(f, nlayer, ngy, ngx, x, y, nout, sq, out);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_gridD5512C(const double* f, int ngy, int ngx, const double* x, const double* y, long npi, int nxo, int nyo,
                      double* out, cudaStream_t s) {
    if (npi <= 0 || nxo <= 0 || nyo <= 0) return 0;
    const size_t smem = sizeof(double) * 10 * ((size_t)nxo + nyo) + sizeof(int) * ((size_t)nxo + nyo);
    B200_REQUIRE(smem <= 200 * 1024, "gridD5512C: nxo+nyo too large for one CTA's shared memory");
    if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(k_gridD5512C, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gridD5512C<<<(unsigned)npi, 256, smem, s>>>(f, ngy, ngx, x, y, nxo, nyo, out);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_layout_tables(const double* src, int ntab, int ns, int pad, int ngrid, int poly, double* dst,
                         cudaStream_t s) {
    if (ntab <= 0 || ns <= 0) return 0;
    B200_REQUIRE(ngrid == ns + 2 * pad && poly >= 0, "layout_tables: ngrid must be ns + 2 pad");
    const long long ncell = poly > 0 ? (ngrid + poly - 1) / poly : 0;
    const long long per = poly > 0 ? (long long)poly * poly * ncell * ncell : (long long)ngrid * ngrid;
    const long long tot = per * ntab;
    k_layout_tables<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(src, ntab, ns, pad, ngrid, poly, dst);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_cmul_conj(const double* ar, const double* ai, const double* br, const double* bi, long long per,
                     long long stride2, long long n1, double im_sign, double* gr, double* gi, cudaStream_t s) {
    const long long total = per * n1;
    if (total <= 0) return 0;
    k_cmul_conj<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(ar, ai, br, bi, per, stride2, total, im_sign, gr, gi);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_getw(double* w, double fh, cudaStream_t s) {
    k_getw<<<1, 1, 0, s>>>(w, fh);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_gather_stamp(const int* idx, int n, int npad, const double* src_x, const double* src_y, const int* src_code,
                        const float* src_data, long src_ld, int n_inframe, double* px, double* py, int* pcode,
                        float* indata, int ldi, cudaStream_t s) {
    if (npad <= 0) return 0;
    k_gather_stamp<<<(npad + 255) / 256, 256, 0, s>>>(idx, n, npad, src_x, src_y, src_code, src_data, src_ld,
                                                      n_inframe, px, py, pcode, indata, ldi);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_build_A(const double* px, const double* py, const int* pcode, int n, int npad, const double* tables,
                   const TableRef* lut, int nimg, int ncode, int ngrid, double dscale, double nc, double flat_penalty,
                   double* A, int lda, double diag_add, int poly, cudaStream_t s) {
    if (npad <= 0) return 0;
    B200_REQUIRE(npad % 32 == 0 && npad >= n && lda >= npad, "build_A: npad must be a multiple of 32, >= n, <= lda");
    const long nt = npad / 32;
    const long ntri = nt * (nt + 1) / 2;
    prof_begin(PROF_BUILD_A, s);
#define B200_BUILD_A(PP)                                                                                           \
    k_build_A<PP><<<(unsigned)ntri, 256, 0, s>>>(px, py, pcode, n, npad, tables, lut, nimg, ncode, ngrid, dscale, nc, \
                                                 flat_penalty, A, lda, diag_add)
    switch (poly) {
        case 0: B200_BUILD_A(0); break;
        case 2: B200_BUILD_A(2); break;
        case 3: B200_BUILD_A(3); break;
        case 4: B200_BUILD_A(4); break;
        case 5: B200_BUILD_A(5); break;
        case 6: B200_BUILD_A(6); break;
        case 8: B200_BUILD_A(8); break;
        case 10: B200_BUILD_A(10); break;
        case 12: B200_BUILD_A(12); break;
        case 16: B200_BUILD_A(16); break;
        default:
            set_error("build_A: no polyphase instantiation for period %d (0, 2-6, 8, 10, 12, 16)", poly);
            return -1;
    }
#undef B200_BUILD_A
    prof_end(8.0 * npad * (double)npad + 20.0 * n, s);  // bytes: the matrix written once + positions/codes read
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_pair_blocks(const double* gx, const double* gy, const int* gimg, const PairDesc* descs,
                       const int* tile_prefix, int npair, int ntiles, const double* tables, const TableRef* lut,
                       int nimg, int ngrid, double dscale, double nc, double flat_penalty, int poly, double* pool,
                       double points, cudaStream_t s) {
    if (npair <= 0 || ntiles <= 0) return 0;
    // experiment knob (tools/coresident_check.py): shared-memory carve-out the kernel asks for, in per cent.  An SM holds
    // one carve-out at a time, so a CTA of this kernel can only sit beside a k_oz_gemm CTA (198 KB of shared memory) when
    // both ask for the large one.
    const char* carve_env = getenv("B200_PAIR_CARVEOUT");
    const int carve = carve_env ? atoi(carve_env) : -1;
    prof_begin(PROF_BUILD_A, s);
    // B200_PAIR_RESIDENT = CTAs per SM of the resident form (0 / unset: one CTA per tile)
    const char* res_env = getenv("B200_PAIR_RESIDENT");
    const int res = res_env ? atoi(res_env) : 0;
    static int n_sm = 0;
    if (res > 0 && n_sm == 0) {
        int dev = 0;
        B200_CUDA(cudaGetDevice(&dev));
        B200_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int res_grid = res > 0 ? (ntiles < res * n_sm ? ntiles : res * n_sm) : 0;
#define B200_PAIR(PP)                                                                                                \
    if (carve >= 0) {                                                                                                \
        B200_CUDA(cudaFuncSetAttribute(k_pair_blocks<PP>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));   \
        B200_CUDA(cudaFuncSetAttribute(k_pair_blocks_resident<PP>, cudaFuncAttributePreferredSharedMemoryCarveout,   \
                                       carve));                                                                      \
    }                                                                                                                \
    if (res_grid > 0)                                                                                                \
        k_pair_blocks_resident<PP><<<(unsigned)res_grid, 128, 0, s>>>(gx, gy, gimg, descs, tile_prefix, npair, ntiles, \
                                                                     tables, lut, nimg, ngrid, dscale, nc,           \
                                                                     flat_penalty, pool);                            \
    else                                                                                                             \
        k_pair_blocks<PP><<<(unsigned)ntiles, 256, 0, s>>>(gx, gy, gimg, descs, tile_prefix, npair, tables, lut, nimg, \
                                                           ngrid, dscale, nc, flat_penalty, pool)
    switch (poly) {
        case 0: B200_PAIR(0); break;
        case 2: B200_PAIR(2); break;
        case 3: B200_PAIR(3); break;
        case 4: B200_PAIR(4); break;
        case 5: B200_PAIR(5); break;
        case 6: B200_PAIR(6); break;
        case 8: B200_PAIR(8); break;
        case 10: B200_PAIR(10); break;
        case 12: B200_PAIR(12); break;
        case 16: B200_PAIR(16); break;
        default:
            set_error("pair_blocks: no polyphase instantiation for period %d (0, 2-6, 8, 10, 12, 16)", poly);
            return -1;
    }
#undef B200_PAIR
    prof_end(8.0 * points, s);  // bytes: every block entry written once
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_assemble_A(const AsmDesc& d, const int* gidx, int n, int npad, const double* pool, double* A, int lda,
                      double diag_add, cudaStream_t s) {
    if (npad <= 0) return 0;
    B200_REQUIRE(npad % 32 == 0 && npad >= n && lda >= npad, "assemble_A: npad must be a multiple of 32, >= n, <= lda");
    const long nt = npad / 32;
    prof_begin(PROF_ASSEMBLE_A, s);
    k_assemble_A<<<(unsigned)(nt * (nt + 1) / 2), 256, 0, s>>>(d, gidx, n, npad, pool, A, lda, diag_add);
    prof_end(8.0 * npad * (double)npad + 8.0 * 0.5 * n * (double)n, s);  // A written once + the upper half read from blocks
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_build_B(const double* px, const double* py, const int* pcode, int n, int npad, const double* tables,
                   const long long* lut_io, int n_out, int ngrid, double dscale, double nc, int n2f, int mpad,
                   double x0out, double y0out, double* B, int ldb, size_t strideB, cudaStream_t s) {
    if (npad <= 0 || mpad <= 0) return 0;
    B200_REQUIRE(npad >= n && mpad >= n2f * n2f && ldb >= npad, "build_B: padded sizes too small");
    static bool attr_done = false;
    if (!attr_done) {
        B200_CUDA(cudaFuncSetAttribute(k_build_B<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        B200_CUDA(cudaFuncSetAttribute(k_build_B_sep<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        B200_CUDA(cudaFuncSetAttribute(k_build_B_sep<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done = true;
    }
    // rows of the strip cache: the n2f output rows span (n2f - 1) / dscale table rows, plus the 10-row window
    const int rmax = (int)ceil((n2f - 1) / dscale) + 12;
    auto sep_smem = [&](int ti) {
        return (size_t)ti * n2f * 2 * 10 * sizeof(double) + (size_t)ti * rmax * n2f * sizeof(double) +
               (size_t)(2 * ti * n2f + 2 * ti) * sizeof(int);
    };
    prof_begin(PROF_BUILD_B, s);
    if (dscale > 0 && sep_smem(4) <= 200 * 1024 && npad % 4 == 0) {
        k_build_B_sep<4><<<(unsigned)((npad + 3) / 4), 512, sep_smem(4), s>>>(px, py, pcode, n, npad, tables, lut_io, n_out,
                                                                             ngrid, dscale, nc, n2f, mpad, x0out, y0out,
                                                                             B, ldb, strideB, rmax);
    } else if (dscale > 0 && sep_smem(2) <= 200 * 1024 && npad % 2 == 0) {
        k_build_B_sep<2><<<(unsigned)((npad + 1) / 2), 512, sep_smem(2), s>>>(px, py, pcode, n, npad, tables, lut_io, n_out,
                                                                             ngrid, dscale, nc, n2f, mpad, x0out, y0out,
                                                                             B, ldb, strideB, rmax);
    } else {
        constexpr int TI = 16;
        const size_t smem = (size_t)TI * n2f * 2 * (10 * sizeof(double) + sizeof(int));
        B200_REQUIRE(smem <= 220 * 1024, "build_B: n2f too large for the shared-memory weight cache");
        k_build_B<TI><<<(unsigned)((npad + TI - 1) / TI), 256, smem, s>>>(px, py, pcode, n, npad, tables, lut_io, n_out,
                                                                           ngrid, dscale, nc, n2f, mpad, x0out, y0out, B,
                                                                           ldb, strideB);
    }
    prof_end(8.0 * n_out * (double)mpad * npad + 20.0 * n, s);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace b200
