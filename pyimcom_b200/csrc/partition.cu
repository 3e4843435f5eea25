// Input-pixel partitioning: InImage.partition_pixels / extract_layers, coadd.py:333-360, 382-408.
//
// The reference walks the relevant cells of a sparse grid over the detector in raster order and, inside a cell, its
// pixels in raster order, appending every pixel that lands in the block (and is unmasked, and whose postage stamp is in
// use) to the list of its stamp.  The lists are therefore ordered by that traversal.  On the device the traversal
// order is kept exactly, without a sort and without atomics on the lists:
//
//   k_part_count    one CTA per cell: stamp of every pixel, and the pixel's rank among the EARLIER pixels of the same
//                   cell that go to the same stamp (warp match + per-warp counters, chunks of 256 pixels in traversal
//                   order); the stamps a cell touches form a small rectangle of stamp indices, which gives each one a
//                   local slot.  Per cell: that rectangle and the number of pixels per slot.
//   k_part_scan     one CTA walks the cells in traversal order and hands every (cell, slot) the number of pixels the
//                   earlier cells sent to that stamp (exclusive running count per stamp).
//   k_part_scatter  one CTA per cell: list position = stamp base + cell base + rank; writes y_idx, x_idx, y_val, x_val.
//   k_part_extract  data[f][stamp][k] = indata[f][y_idx][x_idx]  (extract_layers).
//
// All of it is integer / index work and copies of doubles: results are identical to the reference's.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int PT = 256;       // threads per CTA
constexpr int MAXSLOT = B200_PART_MAXSLOT;  // stamps one sparse cell may touch

// Python's / NumPy's float floor division for a >= 0, b > 0 (float_divmod / npy_divmod): exact through fmod.
__device__ __forceinline__ int py_floordiv_pos(double a, double b) {
    const double mod = fmod(a, b);
    const double div = (a - mod) / b;
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return (int)fl;
}

__global__ void __launch_bounds__(PT) k_part_count(const PartCell* __restrict__ cells, const double* __restrict__ ox,
                                                   const double* __restrict__ oy,
                                                   const unsigned char* __restrict__ mask, int sca,
                                                   const unsigned char* __restrict__ use, int ns, int n2, double lower,
                                                   double upper, int* __restrict__ sid_out,
                                                   unsigned* __restrict__ rank_out, int* __restrict__ cellmeta,
                                                   unsigned* __restrict__ cellcnt, int* __restrict__ err) {
    __shared__ int s_lim[4];  // jmin, jmax, imin, imax over the cell's accepted pixels
    __shared__ unsigned wcount[PT / 32][MAXSLOT];
    __shared__ unsigned cnt[MAXSLOT];
    const PartCell c = cells[blockIdx.x];
    const int npix = c.h * c.w;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_lim[0] = 1 << 30;
        s_lim[1] = -1;
        s_lim[2] = 1 << 30;
        s_lim[3] = -1;
    }
    for (int t = tid; t < (PT / 32) * MAXSLOT; t += PT) (&wcount[0][0])[t] = 0;
    for (int t = tid; t < MAXSLOT; t += PT) cnt[t] = 0;
    __syncthreads();
    // pass 1: stamp of every pixel (coadd.py:341-352)
    for (int p = tid; p < npix; p += PT) {
        const int j = p / c.w, i = p - j * c.w;
        const double x = ox[c.off + p], y = oy[c.off + p];
        int sid = -1;
        if (lower < x && x < upper && lower < y && y < upper && mask[(size_t)(c.bottom + j) * sca + c.left + i]) {
            const int i_st = py_floordiv_pos(x - lower, (double)n2);
            const int j_st = py_floordiv_pos(y - lower, (double)n2);
            if (i_st >= 0 && i_st < ns && j_st >= 0 && j_st < ns && use[j_st * ns + i_st]) {
                sid = j_st * ns + i_st;
                atomicMin(&s_lim[0], j_st);
                atomicMax(&s_lim[1], j_st);
                atomicMin(&s_lim[2], i_st);
                atomicMax(&s_lim[3], i_st);
            }
        }
        sid_out[c.off + p] = sid;
    }
    __syncthreads();
    const int jmin = s_lim[0], imin = s_lim[2];
    const int nj = s_lim[1] >= 0 ? s_lim[1] - jmin + 1 : 0, ni = s_lim[3] >= 0 ? s_lim[3] - imin + 1 : 0;
    const int nslot = nj * ni;
    if (tid == 0) {
        int* m = cellmeta + 4 * (size_t)blockIdx.x;
        m[0] = jmin;
        m[1] = imin;
        m[2] = nj;
        m[3] = ni;
    }
    if (nslot > MAXSLOT) {
        if (tid == 0) atomicMax(err, 1);
        return;
    }
    if (nslot == 0) return;
    // pass 2: rank among the earlier pixels of this cell with the same stamp, 256 pixels at a time in traversal order
    for (int p0 = 0; p0 < npix; p0 += PT) {
        const int p = p0 + tid;
        int s = -1;
        if (p < npix) {
            const int sid = sid_out[c.off + p];  // written by this very thread in pass 1
            if (sid >= 0) s = (sid / ns - jmin) * ni + (sid % ns - imin);
        }
        const unsigned same = __match_any_sync(0xffffffffu, s);
        const unsigned before = __popc(same & ((1u << lane) - 1u));
        if (s >= 0 && (int)(__ffs(same) - 1) == lane) wcount[warp][s] = __popc(same);
        __syncthreads();
        if (s >= 0) {
            unsigned r = cnt[s] + before;
            for (int w = 0; w < warp; w++) r += wcount[w][s];
            rank_out[c.off + p] = r;
        }
        __syncthreads();
        for (int t = tid; t < nslot; t += PT) {
            unsigned tot = 0;
#pragma unroll
            for (int w = 0; w < PT / 32; w++) {
                tot += wcount[w][t];
                wcount[w][t] = 0;
            }
            cnt[t] += tot;
        }
        __syncthreads();
    }
    for (int t = tid; t < nslot; t += PT) cellcnt[(size_t)blockIdx.x * MAXSLOT + t] = cnt[t];
}

// Exclusive running count per stamp over the cells in traversal order; finally pix_count (coadd.py:354-358).
__global__ void __launch_bounds__(PT) k_part_scan(int ncell, const int* __restrict__ cellmeta,
                                                  const unsigned* __restrict__ cellcnt, unsigned* __restrict__ cellbase,
                                                  unsigned* __restrict__ run, int ns, int npixmax,
                                                  unsigned* __restrict__ pix_count, int* __restrict__ err) {
    const int tid = threadIdx.x;
    for (int c = 0; c < ncell; c++) {
        const int* m = cellmeta + 4 * (size_t)c;
        const int jmin = m[0], imin = m[1], nj = m[2], ni = m[3];
        const int nslot = nj * ni;
        if (nslot > MAXSLOT) continue;  // flagged by k_part_count
        for (int t = tid; t < nslot; t += PT) {  // the slots of one cell are distinct stamps
            const int sid = (jmin + t / ni) * ns + imin + t % ni;
            const unsigned b = run[sid];
            cellbase[(size_t)c * MAXSLOT + t] = b;
            run[sid] = b + cellcnt[(size_t)c * MAXSLOT + t];
        }
        __syncthreads();  // the next cell may continue the same stamps
    }
    for (int s = tid; s < ns * ns; s += PT) {
        const unsigned n = run[s];
        pix_count[s] = n;
        if (n > (unsigned)npixmax) atomicMax(err, 2);  // the reference would raise IndexError here
    }
}

__global__ void __launch_bounds__(PT) k_part_scatter(const PartCell* __restrict__ cells, const double* __restrict__ ox,
                                                     const double* __restrict__ oy, const int* __restrict__ sid_in,
                                                     const unsigned* __restrict__ rank_in,
                                                     const int* __restrict__ cellmeta,
                                                     const unsigned* __restrict__ cellbase, int ns, int npixmax,
                                                     unsigned short* __restrict__ y_idx,
                                                     unsigned short* __restrict__ x_idx, double* __restrict__ y_val,
                                                     double* __restrict__ x_val) {
    const PartCell c = cells[blockIdx.x];
    const int* m = cellmeta + 4 * (size_t)blockIdx.x;
    const int jmin = m[0], imin = m[1], ni = m[3];
    if (m[2] * m[3] > MAXSLOT) return;
    const int npix = c.h * c.w;
    for (int p = threadIdx.x; p < npix; p += PT) {
        const int sid = sid_in[c.off + p];
        if (sid < 0) continue;
        const int slot = (sid / ns - jmin) * ni + (sid % ns - imin);
        const unsigned k = cellbase[(size_t)blockIdx.x * MAXSLOT + slot] + rank_in[c.off + p];
        if (k >= (unsigned)npixmax) continue;  // overflow is reported by k_part_scan
        const int j = p / c.w, i = p - j * c.w;
        const size_t dst = (size_t)sid * npixmax + k;
        y_idx[dst] = (unsigned short)(c.bottom + j);
        x_idx[dst] = (unsigned short)(c.left + i);
        y_val[dst] = oy[c.off + p];
        x_val[dst] = ox[c.off + p];
    }
}

// extract_layers (coadd.py:396-404): data (n_inframe, ns*ns, max_count), zero beyond pix_count.
__global__ void __launch_bounds__(PT) k_part_extract(const float* __restrict__ indata, int n_inframe, int sca,
                                                     const unsigned short* __restrict__ y_idx,
                                                     const unsigned short* __restrict__ x_idx,
                                                     const unsigned* __restrict__ pix_count, int nstamp, int npixmax,
                                                     int max_count, float* __restrict__ data) {
    const long tot = (long)n_inframe * nstamp * max_count;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
        const int k = (int)(t % max_count);
        const long r = t / max_count;
        const int sid = (int)(r % nstamp), f = (int)(r / nstamp);
        float v = 0.0f;
        if ((unsigned)k < pix_count[sid]) {
            const size_t src = (size_t)sid * npixmax + k;
            v = indata[((size_t)f * sca + y_idx[src]) * sca + x_idx[src]];
        }
        data[t] = v;
    }
}

// InStamp.__init__ (coadd.py:682-707) for one input image: the image's per-stamp lists go to their place in the block's
// concatenated pixel arrays (stamp-major, images in order inside a stamp): dst_off[sid] = first slot of this image's
// pixels of stamp sid.
__global__ void __launch_bounds__(PT) k_part_assemble(const double* __restrict__ x_val, const double* __restrict__ y_val,
                                                      const float* __restrict__ data, int n_inframe, int nstamp,
                                                      int npixmax, int max_count, const unsigned* __restrict__ pix_count,
                                                      const long long* __restrict__ dst_off, int image, long long npix,
                                                      double* __restrict__ gx, double* __restrict__ gy,
                                                      int* __restrict__ gimg, float* __restrict__ gdata) {
    const int sid = blockIdx.x;
    const unsigned n = pix_count[sid];
    const long long o = dst_off[sid];
    for (unsigned t = threadIdx.x; t < n; t += PT) {
        gx[o + t] = x_val[(size_t)sid * npixmax + t];
        gy[o + t] = y_val[(size_t)sid * npixmax + t];
        gimg[o + t] = image;
        for (int f = 0; f < n_inframe; f++)
            gdata[(size_t)f * npix + o + t] = data[((size_t)f * nstamp + sid) * max_count + t];
    }
}

}  // namespace

int launch_assemble_instamps(const double* x_val, const double* y_val, const float* data, int n_inframe, int nstamp,
                             int npixmax, int max_count, const unsigned* pix_count, const long long* dst_off, int image,
                             long long npix, double* gx, double* gy, int* gimg, float* gdata, cudaStream_t s) {
    if (nstamp <= 0 || npix <= 0) return 0;
    k_part_assemble<<<nstamp, PT, 0, s>>>(x_val, y_val, data, n_inframe, nstamp, npixmax, max_count, pix_count, dst_off,
                                          image, npix, gx, gy, gimg, gdata);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_partition(const PartCell* cells, int ncell, const double* ox, const double* oy, const unsigned char* mask,
                     int sca, const unsigned char* use, int ns, int n2, double lower, double upper, int npixmax,
                     int* sid_tmp, unsigned* rank_tmp, int* cellmeta, unsigned* cellcnt, unsigned* cellbase,
                     unsigned* run, unsigned* pix_count, unsigned short* y_idx, unsigned short* x_idx, double* y_val,
                     double* x_val, int* err, cudaStream_t s) {
    B200_REQUIRE(ns > 0 && n2 > 0 && npixmax > 0 && sca > 0 && sca <= 65536, "partition: bad sizes");
    B200_CUDA(cudaMemsetAsync(run, 0, sizeof(unsigned) * (size_t)ns * ns, s));
    B200_CUDA(cudaMemsetAsync(err, 0, sizeof(int), s));
    if (ncell > 0) {
        k_part_count<<<ncell, PT, 0, s>>>(cells, ox, oy, mask, sca, use, ns, n2, lower, upper, sid_tmp, rank_tmp, cellmeta,
                                          cellcnt, err);
        B200_LAUNCHED(1);
    }
    k_part_scan<<<1, PT, 0, s>>>(ncell, cellmeta, cellcnt, cellbase, run, ns, npixmax, pix_count, err);
    B200_LAUNCHED(1);
    if (ncell > 0) {
        k_part_scatter<<<ncell, PT, 0, s>>>(cells, ox, oy, sid_tmp, rank_tmp, cellmeta, cellbase, ns, npixmax, y_idx, x_idx,
                                            y_val, x_val);
        B200_LAUNCHED(1);
    }
    B200_CUDA(cudaGetLastError());
    return 0;
}

int launch_extract_layers(const float* indata, int n_inframe, int sca, const unsigned short* y_idx,
                          const unsigned short* x_idx, const unsigned* pix_count, int nstamp, int npixmax, int max_count,
                          float* data, cudaStream_t s) {
    const long tot = (long)n_inframe * nstamp * max_count;
    if (tot <= 0) return 0;
    const long grid = (tot + PT - 1) / PT;
    k_part_extract<<<(unsigned)(grid < 148 * 32 ? grid : 148 * 32), PT, 0, s>>>(indata, n_inframe, sca, y_idx, x_idx,
                                                                            pix_count, nstamp, npixmax, max_count, data);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace b200
