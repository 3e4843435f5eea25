// Stage (c): apply the coaddition matrix.
//   k_finalize     lakernel.py:309-317, 390-393 (D, N, node combination) + coadd.py:1321-1354
//                  (fade T, float32 cast, Tsum per input image, outimage = T . indata)
//   k_stamp_maps   coadd.py:1104-1122, 1339-1350 (clamp, fade the U/S/K maps, Tsum_stamp/inpix, Neff)
//   k_accumulate   coadd.py:1976-1994 (overlap-add of one stamp into the block cube / maps)
//   k_unfade_crop  coadd.py:2156-2176, 1284-1292 (block output assembly, SURVEY 8f row f3: recover the faded block
//                  boundary, crop the fade margin)
//   k_compress_map coadd.py:2086-2137 (log-integer encoding of the quality maps)
//
// k_finalize is the HBM-bound T-apply: it reads the f64 node solutions once (8*nv*m*n B) plus -B/2
// (8*m*n B), writes the float32 T (4*m*n B) and produces everything else from registers/shared memory.
// Eight output pixels per CTA; the  (8 pixels) x (n_inframe <= 8 per pass) x n  product runs on the
// FP64 tensor pipe (DMMA.8x8x4) from a shared-memory tile of the faded float32 row values, accumulated
// in f64 (the reference's einsum is float32; f64 accumulation is at least as accurate).
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int FT = 256;     // threads per CTA
constexpr int ROWS = 8;     // output pixels per CTA
constexpr int CH = 512;     // columns per chunk held in shared memory
constexpr int LDT = CH + 4; // smem row stride (doubles): 516 = 4 mod 16 -> conflict-free DMMA fragment loads

// trapezoid weight sequence applied to one float32 value exactly as numpy does it (coadd.py:1269-1282):
// each "*=" promotes to float64, multiplies, and rounds back to float32.
__device__ __forceinline__ float fade32(float v, int iy, int ix, int ny, int nx, int fk2, const double* __restrict__ s) {
    if (fk2 <= 0) return v;
    if (iy < fk2) v = (float)((double)v * s[iy]);
    if (iy > ny - 1 - fk2) v = (float)((double)v * s[ny - 1 - iy]);
    if (ix < fk2) v = (float)((double)v * s[ix]);
    if (ix > nx - 1 - fk2) v = (float)((double)v * s[nx - 1 - ix]);
    return v;
}
__device__ __forceinline__ double fade64(double v, int iy, int ix, int ny, int nx, int fk2, const double* __restrict__ s) {
    if (fk2 <= 0) return v;
    if (iy < fk2) v = v * s[iy];
    if (iy > ny - 1 - fk2) v = v * s[ny - 1 - iy];
    if (ix < fk2) v = v * s[ix];
    if (ix > nx - 1 - fk2) v = v * s[nx - 1 - ix];
    return v;
}

__global__ void __launch_bounds__(FT) k_finalize(FinalizeArgs A) {
    extern __shared__ __align__(16) double sm[];
    double* tile = sm;                         // [ROWS][LDT] faded f32 values (as doubles) of this chunk
    double* din = tile + ROWS * LDT;           // [8][LDT] indata chunk (as doubles), frames of the current pass
    double* red = din + 8 * LDT;               // [ROWS][8 warps][2] D,N partials ; then [8 warps][64] outimage
    double* segsum = red + 8 * 64 + ROWS * 8 * 2;  // [ROWS][nseg]
    __shared__ double sfade[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a0 = blockIdx.x * ROWS;
    const int fk2 = 2 * A.fade;
    for (int k = tid; k < fk2 && k < 64; k += FT) sfade[k] = A.fade_w[k];
    for (int t = tid; t < ROWS * A.nseg; t += FT) segsum[t] = 0.0;
    __syncthreads();

    // warp w owns output pixel a0 + w for the streaming part
    const int a = a0 + warp;
    const bool live = a < A.m;
    const int iy = live ? a / A.n2f : 0, ix = live ? a - iy * A.n2f : 0;
    double wnode[16];
#pragma unroll
    for (int p = 0; p < 16; p++) wnode[p] = (A.w && live && p < A.nv) ? A.w[(size_t)a * A.nv + p] : 0.0;
    double dsum = 0.0, nsum = 0.0;
    const int npass = (A.n_inframe + 7) / 8;
    double oacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // DMMA accumulators: pass 0/1, (pixel g, frames 2q,2q+1)
    int seg = 0;  // warp-uniform running segment pointer (columns are visited in increasing order)

    for (int c0 = 0; c0 < A.n; c0 += CH) {
        const int cw = min(CH, A.n - c0);
        // ---- stream: combine nodes, D/N partials, fade + float32 cast, T32 store, tile fill ----
        // U independent 8-byte loads per operand and lane are issued before any of them is consumed: the stage is a
        // pure HBM/L2 stream and needs bytes in flight, not arithmetic
        constexpr int U = 8;
        for (int jb = 0; jb < CH; jb += 32 * U) {
            double tiv[U], bv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = jb + u * 32 + lane;
                const int col = c0 + j;
                const bool ok = live && j < cw;
                double ti = 0.0;
                if (ok) {
                    if (A.w) {
#pragma unroll
                        for (int p = 0; p < 16; p++)
                            if (p < A.nv) ti += __ldg(A.Tpi + p * A.strideT + (size_t)a * A.ldt + col) * wnode[p];
                    } else {
                        ti = __ldg(A.Tpi + (size_t)a * A.ldt + col);
                    }
                }
                tiv[u] = ti;
                bv[u] = (ok && A.mB) ? __ldg(A.mB + (size_t)a * A.ldb + col) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = jb + u * 32 + lane;
                const int col = c0 + j;
                double tv = 0.0;
                if (live && j < cw) {
                    const double ti = tiv[u];
                    if (A.Ti64) A.Ti64[(size_t)a * A.ldt64 + col] = ti;
                    dsum += bv[u] * ti;
                    nsum += ti * ti;
                    const float t32 = fade32((float)ti, iy, ix, A.n2f, A.n2f, fk2, sfade);
                    if (A.T32) A.T32[(size_t)a * A.ldt32 + col] = t32;
                    tv = (double)t32;
                }
                tile[warp * LDT + j] = tv;
            }
        }
        __syncwarp();
        // per-(instamp,image) segment sums of the faded float32 T (coadd.py:1327-1337), from the tile row
        while (live && seg < A.nseg) {
            const int sb = seg ? A.seg_end[seg - 1] : 0, se = A.seg_end[seg];
            if (sb >= c0 + cw) break;
            const int lo = max(sb, c0), hi = min(se, c0 + cw);
            double sp = 0.0;
            for (int j = lo + lane; j < hi; j += 32) sp += tile[warp * LDT + j - c0];
            sp = warp_sum(sp);
            if (lane == 0) segsum[warp * A.nseg + seg] += sp;
            if (se <= c0 + cw) seg++;
            else break;
        }
        // ---- outimage on the tensor pipe: (8 pixels) x (8 frames) += tile(8 x cw) * din(8 x cw)^T ----
#pragma unroll
        for (int ps = 0; ps < 2; ps++) {
            if (ps < npass) {
                __syncthreads();
                for (int e = tid; e < 8 * CH; e += FT) {
                    const int f = e / CH, j = e - f * CH;
                    const int fr = ps * 8 + f;
                    din[f * LDT + j] =
                        (fr < A.n_inframe && j < cw) ? (double)A.indata[(size_t)fr * A.ldi + c0 + j] : 0.0;
                }
                __syncthreads();
                // warp w takes the k-range [w*64, w*64+64) of the chunk
                const int g = lane >> 2, q = lane & 3;
                const double* ap = tile + g * LDT + warp * 64 + q;
                const double* bp = din + g * LDT + warp * 64 + q;
#pragma unroll
                for (int kk = 0; kk < 16; kk++) dmma884(oacc[ps][0], oacc[ps][1], ap[kk * 4], bp[kk * 4]);
            }
        }
        __syncthreads();
    }
    // ---- reductions ----
    dsum = warp_sum(dsum);
    nsum = warp_sum(nsum);
    if (live && lane == 0) {
        if (A.D) A.D[a] = dsum;
        if (A.N) A.N[a] = nsum;
    }
    // outimage: sum the 8 warps' partial 8x8 blocks (fixed order -> deterministic)
#pragma unroll
    for (int ps = 0; ps < 2; ps++) {
        if (ps >= npass) break;
        __syncthreads();
        const int g = lane >> 2, q = lane & 3;
        red[warp * 64 + g * 8 + q * 2] = oacc[ps][0];
        red[warp * 64 + g * 8 + q * 2 + 1] = oacc[ps][1];
        __syncthreads();
        if (tid < 64) {
            double s = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < 8; w8++) s += red[w8 * 64 + tid];
            const int pg = tid >> 3, f = ps * 8 + (tid & 7);
            if (a0 + pg < A.m && f < A.n_inframe) A.outimage[(size_t)f * A.m + a0 + pg] = (float)s;
        }
    }
    __syncthreads();
    // Tsum_image[a][img] = sum of this pixel's segment sums belonging to image img, in segment order
    for (int t = tid; t < ROWS * A.n_img; t += FT) {
        const int r = t / A.n_img, img = t - r * A.n_img;
        if (a0 + r >= A.m) continue;
        double s = 0.0;
        for (int sg = 0; sg < A.nseg; sg++)
            if (A.seg_img[sg] == img) s += segsum[r * A.nseg + sg];
        A.Tsum_image[(size_t)(a0 + r) * A.n_img + img] = s;
    }
}

// ---- per-stamp maps --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stamp_maps(const double* __restrict__ kappa, const double* __restrict__ Sigma,
                                                    const double* __restrict__ UC, int m, int n2f, int fade,
                                                    int clamp_iter, const double* __restrict__ fade_w,
                                                    float* __restrict__ kappa32, float* __restrict__ Sigma32,
                                                    float* __restrict__ UC32, const double* __restrict__ Tsum_image,
                                                    int n_img, int n2, double* __restrict__ Tsum_stamp,
                                                    double* __restrict__ Tsum_inpix, double* __restrict__ Neff) {
    const int fk2 = 2 * fade;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < m; a += gridDim.x * blockDim.x) {
        const int iy = a / n2f, ix = a - iy * n2f;
        float k = (float)kappa[a], s = (float)Sigma[a], u = (float)UC[a];
        if (clamp_iter) {  // coadd.py:1104-1107
            u = fmaxf(u, 1e-32f);
            s = fmaxf(s, 1e-32f);
        }
        kappa32[a] = fade32(k, iy, ix, n2f, n2f, fk2, fade_w);
        Sigma32[a] = fade32(s, iy, ix, n2f, n2f, fk2, fade_w);
        UC32[a] = fade32(u, iy, ix, n2f, n2f, fk2, fade_w);
        if (Tsum_image) {
            double tot = 0.0, tabs = 0.0;
            for (int g = 0; g < n_img; g++) {
                const double v = Tsum_image[(size_t)a * n_img + g];
                tot += v;
                tabs += fabs(v);
            }
            double sq = 0.0;
            for (int g = 0; g < n_img; g++) {
                const double v = Tsum_image[(size_t)a * n_img + g] / tabs;
                sq += v * v;
            }
            Tsum_inpix[a] = tot;
            Neff[a] = fade64(1.0 / sq, iy, ix, n2f, n2f, fk2, fade_w);
        }
    }
    if (Tsum_image && blockIdx.x == 0) {  // Tsum_stamp[img] = sum_a Tsum_image[a][img] / n2^2
        for (int g = threadIdx.x; g < n_img; g += blockDim.x) {
            double s = 0.0;
            for (int a = 0; a < m; a++) s += Tsum_image[(size_t)a * n_img + g];
            Tsum_stamp[g] = s / ((double)n2 * (double)n2);
        }
    }
}

// ---- overlap-add of one stamp into the block maps (coadd.py:1976-1994) ------------------------------
__global__ void __launch_bounds__(256) k_accumulate(const float* __restrict__ src, int nlayer, int n2f,
                                                    float* __restrict__ dst, int side, int y0, int x0) {
    const int m = n2f * n2f;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < (long)nlayer * m; t += (long)gridDim.x * blockDim.x) {
        const int l = (int)(t / m), a = (int)(t - (long)l * m);
        const int iy = a / n2f, ix = a - iy * n2f;
        float* d = dst + ((size_t)l * side + (y0 + iy)) * side + x0 + ix;
        *d = *d + src[t];
    }
}
__global__ void __launch_bounds__(256) k_accumulate64(const double* __restrict__ src, int nlayer, int n2f,
                                                      float* __restrict__ dst, int side, int y0, int x0) {
    const int m = n2f * n2f;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < (long)nlayer * m; t += (long)gridDim.x * blockDim.x) {
        const int l = (int)(t / m), a = (int)(t - (long)l * m);
        const int iy = a / n2f, ix = a - iy * n2f;
        float* d = dst + ((size_t)l * side + (y0 + iy)) * side + x0 + ix;
        *d = *d + (float)src[t];
    }
}
// All maps of one stamp in one launch (Block._output_stamp_wrapper, coadd.py:1976-1994): the n_inframe coadded layers,
// the U/C, Sigma, kappa maps (float32), the Tsum_inpix and Neff maps (float64 -> float32 as the reference's float32 maps
// take them) and the per-image weight sums of T_weightmap.  Layer index: 0..nfr-1 image, then U, S, K, Tsum, Neff.
struct AccumStamp {
    const float *outimage, *UC, *Sigma, *kappa;
    const double *Tsum_inpix, *Neff, *Tsum_stamp;
    float *out_map, *UC_map, *Sigma_map, *kappa_map, *Tsum_map, *Neff_map, *T_weight;
    int nfr, n2f, side, y0, x0, n_img, tw_stride;
};
__global__ void __launch_bounds__(256) k_accumulate_stamp(AccumStamp a) {
    const int m = a.n2f * a.n2f;
    const long tot = (long)(a.nfr + 5) * m;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
        const int l = (int)(t / m), px = (int)(t - (long)l * m);
        const int iy = px / a.n2f, ix = px - iy * a.n2f;
        const size_t off = (size_t)(a.y0 + iy) * a.side + a.x0 + ix;
        if (l < a.nfr) {
            float* d = a.out_map + (size_t)l * a.side * a.side + off;
            *d = *d + a.outimage[(size_t)l * m + px];
        } else {
            const int q = l - a.nfr;
            float* d = (q == 0 ? a.UC_map : q == 1 ? a.Sigma_map : q == 2 ? a.kappa_map : q == 3 ? a.Tsum_map : a.Neff_map) + off;
            const float v = q == 0 ? a.UC[px] : q == 1 ? a.Sigma[px] : q == 2 ? a.kappa[px]
                          : q == 3 ? (float)a.Tsum_inpix[px] : (float)a.Neff[px];
            *d = *d + v;
        }
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < a.n_img)
        a.T_weight[(size_t)threadIdx.x * a.tw_stride] = (float)a.Tsum_stamp[threadIdx.x];
}


// ---- block output assembly (coadd.py:2086-2328) -----------------------------------
// Block.build_output_file(is_final=True) divides the trapezoid weights back out of the block boundary
// (OutStamp.trapezoid(recover_mode=True): sides B, T, L, R in that order, each "arr /= s" computed in float64 and rounded
// back to float32, coadd.py:1284-1292) and stores the maps without the fade margin [fk : side - fk].  One pass: read
// the (nlayer, side, side) map, write the cropped (nlayer, side - 2 fk, side - 2 fk) map; the block map itself is not
// modified.  recover = 0 only crops (intermediate outputs).
__global__ void __launch_bounds__(256) k_unfade_crop(const float* __restrict__ in, int nlayer, int side, int fk,
                                                     int recover, int pb, int pt, int pl, int pr,
                                                     const double* __restrict__ s, float* __restrict__ out) {
    const int so = side - 2 * fk, fk2 = 2 * fk;
    const int it = side - pt - 1, ir = side - pr - 1;
    const long tot = (long)nlayer * so * so;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
        const int l = (int)(t / ((long)so * so));
        const int a = (int)(t - (long)l * so * so);
        const int iy = a / so + fk, ix = a - (a / so) * so + fk;
        float v = in[((size_t)l * side + iy) * side + ix];
        if (recover && fk2 > 0) {
            if (iy >= pb && iy < pb + fk2) v = (float)((double)v / s[iy - pb]);
            if (iy <= it && iy > it - fk2) v = (float)((double)v / s[it - iy]);
            if (ix >= pl && ix < pl + fk2) v = (float)((double)v / s[ix - pl]);
            if (ix <= ir && ix > ir - fk2) v = (float)((double)v / s[ir - ix]);
        }
        out[t] = v;
    }
}

// Block.compress_map (coadd.py:2129-2131):  clip(floor(coef * log10(clip(x, 1e-32, None)) + 0.5), lo, hi).astype(int16 or
// uint16), every step in float32 as NumPy evaluates it for a float32 map.  log10 is taken in float64 and rounded to
// float32 (the correctly rounded float32 logarithm); NumPy's own float32 log10 is whatever libm / SVML kernel the host
// dispatches to (< 1 ulp, not always correctly rounded), so codes can differ from a given host by one count on the few
// pixels whose scaled logarithm lies within a float32 ulp of a rounding boundary -- the tests bound both the size (1)
// and the rate of such differences.
__global__ void __launch_bounds__(256) k_compress_map(const float* __restrict__ in, long n, float coef, float lo, float hi,
                                                      int is_unsigned, void* __restrict__ out) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        const float x = fmaxf(in[t], 1e-32f);
        const float lg = (float)log10((double)x);
        float v = floorf(__fadd_rn(__fmul_rn(coef, lg), 0.5f));
        v = fminf(fmaxf(v, lo), hi);
        if (is_unsigned)
            ((unsigned short*)out)[t] = (unsigned short)(int)v;
        else
            ((short*)out)[t] = (short)(int)v;
    }
}

}  // namespace

size_t finalize_smem(int nseg) {
    return sizeof(double) * ((size_t)ROWS * LDT + 8 * LDT + 8 * 64 + ROWS * 8 * 2 + (size_t)ROWS * (nseg > 0 ? nseg : 1));
}

int launch_finalize(const FinalizeArgs& a, cudaStream_t s) {
    if (a.m <= 0 || a.n <= 0) return 0;
    B200_REQUIRE(a.n_inframe <= 16, "finalize handles at most 16 input layers per launch");
    B200_REQUIRE(a.nv <= 16 && a.fade <= 32, "finalize: nv <= 16, fade <= 32");
    const size_t smem = finalize_smem(a.nseg);
    B200_REQUIRE(smem <= 200 * 1024, "finalize: too many (instamp,image) segments");
    static bool done = false;
    if (!done) {
        B200_CUDA(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        done = true;
    }
    prof_begin(PROF_FINALIZE, s);
    k_finalize<<<(a.m + ROWS - 1) / ROWS, FT, smem, s>>>(a);
    // bytes: node solutions + (-B/2) read once, float32 T written once, layers read once
    prof_end((double)a.m * a.n * (8.0 * a.nv + (a.mB ? 8.0 : 0.0) + (a.T32 ? 4.0 : 0.0) + (a.Ti64 ? 8.0 : 0.0)) +
                 4.0 * a.n_inframe * (double)a.n,
             s);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_stamp_maps(const double* kappa, const double* Sigma, const double* UC, int m, int n2f, int fade,
                      int clamp_iter, const double* fade_w, float* kappa32, float* Sigma32, float* UC32,
                      const double* Tsum_image, int n_img, int n2, double* Tsum_stamp, double* Tsum_inpix, double* Neff,
                      cudaStream_t s) {
    if (m <= 0) return 0;
    k_stamp_maps<<<(m + 255) / 256, 256, 0, s>>>(kappa, Sigma, UC, m, n2f, fade, clamp_iter, fade_w, kappa32, Sigma32,
                                                 UC32, Tsum_image, n_img, n2, Tsum_stamp, Tsum_inpix, Neff);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_accumulate(const void* src, int src_is_f64, int nlayer, int n2f, float* dst, int side, int y0, int x0,
                      cudaStream_t s) {
    const long tot = (long)nlayer * n2f * n2f;
    if (tot <= 0) return 0;
    B200_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + n2f <= side && x0 + n2f <= side, "stamp outside the block canvas");
    const int grid = (int)((tot + 255) / 256);
    if (src_is_f64)
        k_accumulate64<<<grid, 256, 0, s>>>((const double*)src, nlayer, n2f, dst, side, y0, x0);
    else
        k_accumulate<<<grid, 256, 0, s>>>((const float*)src, nlayer, n2f, dst, side, y0, x0);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_accumulate_stamp(const float* outimage, int nfr, const float* UC, const float* Sigma, const float* kappa,
                            const double* Tsum_inpix, const double* Neff, const double* Tsum_stamp, int n_img, int n2f,
                            float* out_map, float* UC_map, float* Sigma_map, float* kappa_map, float* Tsum_map,
                            float* Neff_map, int side, int y0, int x0, float* T_weight, int tw_stride, cudaStream_t s) {
    B200_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + n2f <= side && x0 + n2f <= side, "stamp outside the block canvas");
    B200_REQUIRE(n_img <= 256 && nfr >= 0, "accumulate_stamp: at most 256 input images");
    AccumStamp a{outimage, UC, Sigma, kappa, Tsum_inpix, Neff, Tsum_stamp, out_map, UC_map, Sigma_map, kappa_map,
                 Tsum_map, Neff_map, T_weight, nfr, n2f, side, y0, x0, n_img, tw_stride};
    const long tot = (long)(nfr + 5) * n2f * n2f;
    k_accumulate_stamp<<<(int)((tot + 255) / 256), 256, 0, s>>>(a);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_unfade_crop(const float* in, int nlayer, int side, int fk, int recover, int pb, int pt, int pl, int pr,
                       const double* fade_w, float* out, cudaStream_t s) {
    const int so = side - 2 * fk;
    if (nlayer <= 0 || so <= 0) return 0;
    B200_REQUIRE(fk >= 0 && pb >= 0 && pt >= 0 && pl >= 0 && pr >= 0, "unfade_crop: negative widths");
    B200_REQUIRE(!recover || fk == 0 || (fade_w != nullptr && side > 4 * fk + pb + pt && side > 4 * fk + pl + pr),
                 "unfade_crop: map too small for its fade and padding widths");
    const long tot = (long)nlayer * so * so;
    const long grid = (tot + 255) / 256;
    k_unfade_crop<<<(unsigned)(grid < 148 * 32 ? grid : 148 * 32), 256, 0, s>>>(in, nlayer, side, fk, recover, pb, pt, pl,
                                                                            pr, fade_w, out);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_compress_map(const float* in, long n, int coef, int is_unsigned, void* out, cudaStream_t s) {
    if (n <= 0) return 0;
    const long grid = (n + 255) / 256;
    k_compress_map<<<(unsigned)(grid < 148 * 32 ? grid : 148 * 32), 256, 0, s>>>(
        in, n, (float)coef, is_unsigned ? 0.0f : -32768.0f, is_unsigned ? 65535.0f : 32767.0f, is_unsigned, out);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace b200
