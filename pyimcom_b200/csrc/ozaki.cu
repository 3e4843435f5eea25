// FP64 products on the 5th-generation tensor cores: error-free integer slicing (Ozaki scheme) + INT8 tcgen05.mma.
//
// The long-K updates of the batched Cholesky / triangular solves (lakernel.py:263, 276, 304, 358 -> linalg.cu) are
//     C (M x N, float64)  -=  A (M x K) * B (N x K)^T
// with both operands FINISHED panels of L / Z / T.  sm_100a has no FP64 kind on tcgen05 (FP64 tensor math is the legacy
// DMMA pipe, ~36 TFLOP/s); its INT8 kind runs at 4.5 Pop/s with exact INT32 accumulation.  So each finished panel is
// sliced ONCE, when it is completed, into NS signed radix-128 digit planes per row-scaled element,
//     x[r][k]  =  2^e[r] * sum_t d_t[r][k] * 128^-(t+1)  (+ truncation < 2^e[r] * 128^-NS / 2),   |d_t| <= 64,
// eight int8 planes take exactly the 8 bytes of the double they describe, and a product becomes
//     sum_k A[i][k] B[j][k]  =  2^(eA[i] + eB[j] - 14) * sum_d 128^-d * ( sum_{t+u=d} sum_k dA_t[i][k] dB_u[j][k] ).
// The inner sums are exact INT8 x INT8 -> INT32 tensor-core GEMMs (|sum| <= (d+1) K 2^12 < 2^31 for K < 65 k), one TMEM
// accumulator per significance level d = t + u < NS, all levels of a tile live in TMEM at once (NS x 64 columns), and
// the only rounding happens in the epilogue's float64 Horner sum over the NS levels and in the final subtraction.
//
// k_oz_gemm: one 128 x 64 output tile per CTA, warp-specialised:
//   warp 0    TMA producer: per 64-wide K chunk one cp.async.bulk.tensor (4-D box: 64 B x rows x NS planes) per operand
//             into a 2-stage ring of 64-byte-swizzled shared-memory tiles, completion on mbarriers;
//   warp 1    MMA issuer: NS(NS+1)/2 tcgen05.mma.kind::i8 (M = 128, N = 64, K = 32) per 32-byte K step, accumulators
//             in TMEM, tcgen05.commit releases the stage;
//   warps 2-5 epilogue: tcgen05.ld of the NS accumulators, Horner in float64, row/column scales, C -= result.
// Slice planes live in global memory as S[K/64][NS][rows][64] (int8), so that every TMA box is NS contiguous 8 KB / 4 KB
// pieces.  SASS: UTCIMMA (tcgen05.mma.kind::i8), UTMALDG (TMA), LDTM (tcgen05.ld).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "ozaki.cuh"

namespace b200 {

namespace {

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, INT8 x INT8 -> INT32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of one K-major digit plane: rows of 64 bytes, 64-byte swizzle (8-row atoms of 512 B).
// Bit fields as in cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start address [0,14) and stride byte
// offset [32,46) in 16-byte units, version [46,48) = 1 on sm_100, layout type [61,64): 4 = SWIZZLE_64B.  The leading
// byte offset [16,30) is not used by swizzled K-major layouts (CUTLASS writes 1).
__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(8 * OZ_BK >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format [4,6) = 2 (S32), a_format [7,10) = b_format [10,13) = 1
// (signed 8-bit), both operands K-major, n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4.
__device__ __forceinline__ constexpr uint32_t oz_idesc(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
}

constexpr int OZ_THREADS = 320;  // TMA warp, MMA warp, eight epilogue warps
constexpr int OZ_A_STAGE = OZ_NS * OZ_BM * OZ_BK;  // 65536
constexpr int OZ_B_STAGE = OZ_NS * OZ_BN * OZ_BK;  // 32768
constexpr int OZ_STAGE = OZ_A_STAGE + OZ_B_STAGE;
constexpr int OZ_TMEM_COLS = 512;  // NS accumulators of 64 columns (power of two >= NS * 64)
static_assert(OZ_NS * OZ_BN <= OZ_TMEM_COLS, "the accumulators of all significance levels must fit in TMEM");
constexpr int OZ_EPI_LD = OZ_BN + 1;  // row stride (doubles) of the epilogue's staging tile
constexpr size_t OZ_SMEM = 1024 + (size_t)OZ_STAGES * OZ_STAGE + 1024;

// (B200_OZ_LB = 512 caps the kernel at 128 registers so that a small CTA of another stream can sit beside it: experiment
// knob, tools/coresident_check.py)
#ifndef B200_OZ_LB
#define B200_OZ_LB OZ_THREADS
#endif
__global__ void __launch_bounds__(B200_OZ_LB, 1) k_oz_gemm(const __grid_constant__ OzBatch p) {
    const OzSys& s = p.s[blockIdx.z];
    const int tn = blockIdx.x, tm = blockIdx.y;
    if (tm >= s.m_tiles || tn >= s.n_tiles || s.kb1 <= s.kb0) return;
    // triangular output (Cholesky update of a symmetric matrix): tiles strictly above the diagonal are never read
    if (s.tri && s.colC0 + tn * OZ_BN > s.rowC0 + tm * OZ_BM + OZ_BM - 1) return;
    extern __shared__ uint8_t oz_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)oz_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* tail = base + (size_t)OZ_STAGES * OZ_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);  // [OZ_STAGES]
    uint64_t* empty = full + OZ_STAGES;                    // [OZ_STAGES]
    uint64_t* accum_done = empty + OZ_STAGES;              // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_done + 1);
    double* sB = reinterpret_cast<double*>(tail + 128);  // [OZ_BN] column scales of this tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = s.kb1 - s.kb0;
    const int dbg = p.dbg;  // timing experiments only: 1 = no MMAs, 2 = only the first two TMA fills (results invalid)
    unsigned long long t_start = 0, t_setup = 0, t_main = 0;
    if (p.dbgbuf) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

    if (tid == 0) {
        for (int i = 0; i < OZ_STAGES; i++) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, 1);
        }
        mbar_init(accum_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // TMEM allocation (the same warp frees it)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(OZ_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid >= 64 && tid < 64 + OZ_BN) sB[tid - 64] = s.scaleB[s.rowB0 + tn * OZ_BN + tid - 64];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (p.dbgbuf) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_setup));

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ----
            for (int it = 0; it < nk; it++) {
                const int st = it % OZ_STAGES, ph = (it / OZ_STAGES) & 1;
                if (dbg == 2 && it >= OZ_STAGES) break;
                mbar_wait(empty + st, ph ^ 1);
                mbar_expect_tx(full + st, OZ_STAGE);
                uint8_t* sa = base + (size_t)st * OZ_STAGE;
                tma_load_4d(sa, &s.mapA, full + st, 0, s.rowA0 + tm * OZ_BM, 0, s.kb0 + it);
                tma_load_4d(sa + OZ_A_STAGE, &s.mapB, full + st, 0, s.rowB0 + tn * OZ_BN, 0, s.kb0 + it);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            for (int it = 0; it < nk; it++) {
                const int st = it % OZ_STAGES, ph = (it / OZ_STAGES) & 1;
                if (!(dbg == 2 && it >= OZ_STAGES)) mbar_wait(full + st, ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(base + (size_t)st * OZ_STAGE), sb = sa + OZ_A_STAGE;
#pragma unroll
                for (int ks = 0; ks < (dbg == 1 ? 0 : OZ_BK / 32); ks++) {
                    // Levels are taken four at a time: the digit planes of B are consecutive 64-row pieces of one
                    // K-major array, so planes u_lo .. u_hi form ONE operand of N = 64 (u_hi - u_lo + 1) rows, and
                    // A_t times that window lands on the accumulators of levels t + u_lo .. t + u_hi, which sit side
                    // by side in TMEM.  12 instructions per K step instead of 36 products of N = 64 (an M = 128
                    // instruction costs ~128 clocks whatever its N: measured 133 clk at N = 64).
#pragma unroll
                    for (int g = 0; g < (OZ_NS + 3) / 4; g++) {
                        const int L0 = 4 * g, L1 = (4 * g + 3 < OZ_NS - 1) ? 4 * g + 3 : OZ_NS - 1;
#pragma unroll
                        for (int t = 0; t <= L1; t++) {
                            const int u_lo = L0 - t > 0 ? L0 - t : 0;
                            const int u_hi = L1 - t < OZ_NS - 1 ? L1 - t : OZ_NS - 1;
                            if (u_lo > u_hi) continue;
                            const int nwin = OZ_BN * (u_hi - u_lo + 1);
                            const uint64_t da = oz_smem_desc(sa + t * (OZ_BM * OZ_BK) + ks * 32);
                            const uint64_t db = oz_smem_desc(sb + u_lo * (OZ_BN * OZ_BK) + ks * 32);
                            umma_i8(tmem + (t + u_lo) * OZ_BN, da, db, oz_idesc(nwin), (it | ks | t) != 0);
                        }
                    }
                }
                tc_commit(empty + st);  // the stage may be refilled once these MMAs have read it
            }
            tc_commit(accum_done);
        }
    } else {
        // ---- epilogue: warps 2..9.  A warp can only read the TMEM lanes 32 (warp % 4) .. + 31 (= rows of the tile);
        // two warps share each lane quadrant and split the 64 columns, so that the drain of the 8 accumulators (the
        // only part of a tile's life the tensor core waits for) runs on eight warps.
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        constexpr int HC = OZ_BN / 2;  // columns per warp
        // the piece of C the warp will update is fetched while the tensor core works: two rows of 32 doubles per step,
        // one 16-byte piece per lane (16 independent loads in flight per lane)
        double* Cw = s.C + (size_t)(tm * OZ_BM + quad * 32 + (lane >> 4)) * s.ldc + (size_t)tn * OZ_BN + half * HC +
                     2 * (lane & 15);
        double2 cin[16];
#pragma unroll
        for (int r = 0; r < 16; r++) cin[r] = *reinterpret_cast<const double2*>(Cw + (size_t)(2 * r) * s.ldc);
        mbar_wait(accum_done, 0);
        tc_fence_after();
        if (p.dbgbuf) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_main));
        const double sa = s.scaleA[s.rowA0 + tm * OZ_BM + row];
        // phase 1: own row -> float64 values (Horner over the levels, scaled) parked in shared memory.  The stage ring is
        // free: accum_done says that every MMA has finished reading it.  Row stride 65 doubles: lane l starts in bank 2 l.
        double* stg = reinterpret_cast<double*>(base) + (size_t)(quad * 32) * OZ_EPI_LD + half * HC;
        double* mine = stg + (size_t)lane * OZ_EPI_LD;
#pragma unroll 1
        for (int c0 = 0; c0 < HC; c0 += 16) {
            double acc[16];
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + half * HC + c0;
            tmem_ld16(taddr + (OZ_NS - 1) * OZ_BN, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j++) acc[j] = (double)(int)v[j];
#pragma unroll
            for (int d = OZ_NS - 2; d >= 0; d--) {  // Horner over the significance levels: one rounding per level
                tmem_ld16(taddr + d * OZ_BN, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) acc[j] = fma(acc[j], 0.0078125, (double)(int)v[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; j++) mine[c0 + j] = acc[j] * (sa * sB[half * HC + c0 + j]);  // powers of two: exact
        }
        __syncwarp();
        // phase 2: the warp walks its 32 rows, two at a time
        const double* sp = stg + (size_t)(lane >> 4) * OZ_EPI_LD + 2 * (lane & 15);
#pragma unroll
        for (int r = 0; r < 16; r++) {
            double2 c = cin[r];
            c.x -= sp[(2 * r) * OZ_EPI_LD];
            c.y -= sp[(2 * r) * OZ_EPI_LD + 1];
            *reinterpret_cast<double2*>(Cw + (size_t)(2 * r) * s.ldc) = c;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.dbgbuf && tid == 64) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        atomicAdd(p.dbgbuf + 0, t_setup - t_start);
        atomicAdd(p.dbgbuf + 1, t_main - t_setup);
        atomicAdd(p.dbgbuf + 2, t_end - t_main);
        atomicAdd(p.dbgbuf + 3, 1ull);
        atomicMin(p.dbgbuf + 4, t_start);
        atomicMax(p.dbgbuf + 5, t_end);
    }
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(OZ_TMEM_COLS) : "memory");
    }
}

// ---- slicing ---------------------------------------------------------------------------------------------------
// Row scale from the exact maximum over a column range: scale = 2^(e - 7) with |x| <= 2^(e-1) for every x of the row.
__global__ void __launch_bounds__(256) k_oz_scale_max(OzSliceBatch p) {
    const OzSliceSys& s = p.s[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= s.nrows) return;
    const double* src = s.src + (size_t)(s.row0 + r) * s.ld + (size_t)s.kb0 * OZ_BK;
    const int ncol = s.nkb * OZ_BK;
    const int cfirst = s.tri ? ((s.row0 + r) / NB + 1) * NB - s.kb0 * OZ_BK : 0;
    if (s.colmode == 2) {  // quotients by the column scales are bounded a priori: the row scale is 1
        if (lane == 0) s.scale[s.row0 + r] = 1.0;
        return;
    }
    const double* cs = s.colmode == 1 ? s.colscale + (size_t)s.kb0 * OZ_BK : nullptr;
    double m = 0.0;
    for (int c = (cfirst > 0 ? cfirst : 0) + lane * 2; c < ncol; c += 64) {
        double2 v = *reinterpret_cast<const double2*>(src + c);
        if (cs) {
            v.x *= cs[c];
            v.y *= cs[c + 1];
        }
        m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        int q = 0;
        double sc = 0.0;
        if (m > 0.0 && m < 1.0e300) {
            frexp(m, &q);  // m = f 2^q, f in [0.5, 1)  ->  |x| <= m < 2^q = 2^(e-1)
            sc = ldexp(1.0, q + 1 - 7);
        }
        s.scale[s.row0 + r] = sc;
    }
}

// Row scale of the rows of a Cholesky factor from the diagonal of the matrix BEFORE it is factored: sum_k L_ik^2 = W_ii,
// so |L_ik| <= sqrt(W_ii) for every finished column k (a priori, the same for all K chunks, so that the INT32
// accumulators can run over the whole K range of a left-looking update).
__global__ void k_oz_scale_diag(OzSliceBatch p) {
    const OzSliceSys& s = p.s[blockIdx.y];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= s.nrows) return;
    const double w = fabs(s.src[(size_t)(s.row0 + r) * s.ld + s.row0 + r]);
    double sc = 0.0;
    if (w > 0.0 && w < 1.0e300) {
        int q = 0;
        frexp(sqrt(w) * (1.0 + 1e-9), &q);
        sc = ldexp(1.0, q + 1 - 7);
    }
    s.scale[s.row0 + r] = sc;
}

// Digits: thread = 8 consecutive columns of one row; plane t of K chunk kb sits at S[((kb * NS + t) * rows + row) * 64].
__global__ void __launch_bounds__(256) k_oz_slice(OzSliceBatch p) {
    const OzSliceSys& s = p.s[blockIdx.z];
    const int r = blockIdx.y * 32 + (threadIdx.x >> 3);
    if (r >= s.nrows || (int)blockIdx.x >= s.nkb) return;
    const int kb = s.kb0 + blockIdx.x, c8 = (threadIdx.x & 7) * 8;
    const int row = s.row0 + r;
    if (s.tri && kb < (row / NB + 1) * (NB / OZ_BK)) return;
    const double sc = s.scale[row];
    const double inv = sc > 0.0 ? 1.0 / sc : 0.0;  // power of two: exact
    const double* src = s.src + (size_t)row * s.ld + (size_t)kb * OZ_BK + c8;
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const double2 v = *reinterpret_cast<const double2*>(src + j);
        x[j] = v.x * inv;
        x[j + 1] = v.y * inv;
    }
    if (s.colmode) {  // powers of two: exact
        const double* cs = s.colscale + (size_t)kb * OZ_BK + c8;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double c = cs[j];
            x[j] *= s.colmode == 1 ? c : (c > 0.0 ? 1.0 / c : 0.0);
        }
    }
    int8_t* dst = s.S + (((size_t)kb * OZ_NS) * s.rows_total + row) * OZ_BK + c8;
    // Round-to-nearest-integer by the magic-number addition (|x| < 2^51): the low word of x + 1.5 * 2^52 IS the integer
    // in two's complement, so no float -> int conversion is needed and a digit costs four FP64 additions / products.
    // |x| <= 64 for the first digit (the scale bounds the row; garbage input of a failed factorisation is clamped),
    // |x| <= 64 for every later one by construction (the remainder of a rounding is at most 1/2, times 128).
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = fmin(fmax(x[j], -64.0), 64.0);
#pragma unroll
    for (int t = 0; t < OZ_NS; t++) {
        uint32_t w[2] = {0u, 0u};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double r = x[j] + MAGIC;
            const int di = __double2loint(r);
            const double d = r - MAGIC;
            x[j] = (x[j] - d) * 128.0;
            w[j >> 2] |= ((uint32_t)di & 0xffu) << (8 * (j & 3));
        }
        *reinterpret_cast<uint2*>(dst + (size_t)t * s.rows_total * OZ_BK) = make_uint2(w[0], w[1]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_oz_attr = false;

int oz_init() {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        B200_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    if (!g_oz_attr) {
        B200_CUDA(cudaFuncSetAttribute(k_oz_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM));
        g_oz_attr = true;
    }
    return 0;
}

}  // namespace

// Tensor map over the digit planes S[nkb][NS][rows][64] of `rows` matrix rows; a box is (64 B, box_rows, NS, 1).
int oz_make_map(CUtensorMap* map, const int8_t* S, int rows, int nkb, int box_rows) {
    if (int rc = oz_init()) return rc;
    const cuuint64_t gdim[4] = {(cuuint64_t)OZ_BK, (cuuint64_t)rows, (cuuint64_t)OZ_NS, (cuuint64_t)nkb};
    const cuuint64_t gstr[3] = {(cuuint64_t)OZ_BK, (cuuint64_t)rows * OZ_BK, (cuuint64_t)rows * OZ_BK * OZ_NS};
    const cuuint32_t box[4] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, (cuuint32_t)OZ_NS, 1u};
    const cuuint32_t est[4] = {1u, 1u, 1u, 1u};
    const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<int8_t*>(S), gdim, gstr, box, est,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d (rows %d, nkb %d, box rows %d)", (int)r, rows, nkb, box_rows);
        return -1;
    }
    return 0;
}

int oz_launch_gemm(const OzBatch& b, int nsys, int m_tiles_max, int n_tiles_max, cudaStream_t st) {
    if (nsys <= 0 || m_tiles_max <= 0 || n_tiles_max <= 0) return 0;
    if (int rc = oz_init()) return rc;
    k_oz_gemm<<<dim3(n_tiles_max, m_tiles_max, nsys), OZ_THREADS, OZ_SMEM, st>>>(b);
    B200_LAUNCH_CHECK();
    return 0;
}

int oz_launch_scale_max(const OzSliceBatch& b, int nsys, int nrows_max, cudaStream_t st) {
    if (nsys <= 0 || nrows_max <= 0) return 0;
    k_oz_scale_max<<<dim3((nrows_max + 7) / 8, nsys), 256, 0, st>>>(b);
    B200_LAUNCH_CHECK();
    return 0;
}

int oz_launch_scale_diag(const OzSliceBatch& b, int nsys, int nrows_max, cudaStream_t st) {
    if (nsys <= 0 || nrows_max <= 0) return 0;
    k_oz_scale_diag<<<dim3((nrows_max + 255) / 256, nsys), 256, 0, st>>>(b);
    B200_LAUNCH_CHECK();
    return 0;
}

int oz_launch_slice(const OzSliceBatch& b, int nsys, int nrows_max, int nkb_max, cudaStream_t st) {
    if (nsys <= 0 || nrows_max <= 0 || nkb_max <= 0) return 0;
    k_oz_slice<<<dim3(nkb_max, (nrows_max + 31) / 32, nsys), 256, 0, st>>>(b);
    B200_LAUNCH_CHECK();
    return 0;
}

size_t oz_gemm_work_bytes(int M, int N, int K) {
    return (size_t)(M + N) * K * OZ_NS + (size_t)(M + N) * sizeof(double) + 4096;
}

// C (M x N) -= A (M x K) B (N x K)^T through the sliced INT8 path (M multiple of 128, N of 64, K of 64): the library's
// FP64-equivalent GEMM on the tcgen05 pipe, and the unit test of the scheme against the DMMA GEMM.
int launch_ozaki_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                         void* work, size_t work_bytes, cudaStream_t st) {
    B200_REQUIRE(M > 0 && N > 0 && K > 0 && M % OZ_BM == 0 && N % OZ_BN == 0 && K % OZ_BK == 0 && lda % 2 == 0 &&
                     ldb % 2 == 0 && ldc % 2 == 0 && K <= 60000,
                 "ozaki_gemm_nt wants M % 128 == 0, N % 64 == 0, K % 64 == 0, K <= 60000");
    B200_REQUIRE(work != nullptr && work_bytes >= oz_gemm_work_bytes(M, N, K), "workspace too small");
    if (int rc = oz_init()) return rc;
    uint8_t* wp = static_cast<uint8_t*>(work);
    double* scA = reinterpret_cast<double*>(wp);
    double* scB = scA + M;
    int8_t* SA = reinterpret_cast<int8_t*>(wp + (((size_t)(M + N) * sizeof(double) + 1023) & ~(size_t)1023));
    int8_t* SB = SA + (size_t)M * K * OZ_NS;
    const int nkb = K / OZ_BK;
    OzSliceBatch sb;
    sb.s[0] = OzSliceSys{A, lda, 0, M, 0, nkb, scA, SA, M, 0};
    sb.s[1] = OzSliceSys{B, ldb, 0, N, 0, nkb, scB, SB, N, 0};
    if (int rc = oz_launch_scale_max(sb, 2, M > N ? M : N, st)) return rc;
    if (int rc = oz_launch_slice(sb, 2, M > N ? M : N, nkb, st)) return rc;
    OzBatch gb;
    {
        const char* e = getenv("B200_OZ_DBG");
        gb.dbg = e ? atoi(e) : 0;
        static unsigned long long* dbgbuf = nullptr;
        if (getenv("B200_OZ_TIMERS")) {
            if (!dbgbuf) B200_CUDA(cudaMalloc(&dbgbuf, 64));
            unsigned long long init[8] = {0, 0, 0, 0, ~0ull, 0, 0, 0};
            B200_CUDA(cudaMemcpyAsync(dbgbuf, init, 64, cudaMemcpyHostToDevice, st));
            gb.dbgbuf = dbgbuf;
        }
    }
    OzSys& g = gb.s[0];
    if (int rc = oz_make_map(&g.mapA, SA, M, nkb, OZ_BM)) return rc;
    if (int rc = oz_make_map(&g.mapB, SB, N, nkb, OZ_BN)) return rc;
    g.scaleA = scA;
    g.scaleB = scB;
    g.C = C;
    g.ldc = ldc;
    g.m_tiles = M / OZ_BM;
    g.n_tiles = N / OZ_BN;
    g.rowA0 = g.rowB0 = 0;
    g.rowC0 = g.colC0 = 0;
    g.kb0 = 0;
    g.kb1 = nkb;
    g.tri = 0;
    prof_begin(PROF_GEMM, st);
    const int rc = oz_launch_gemm(gb, 1, g.m_tiles, g.n_tiles, st);
    prof_end(2.0 * M * (double)N * K, st);
    if (gb.dbgbuf) {
        unsigned long long h[8];
        B200_CUDA(cudaMemcpyAsync(h, gb.dbgbuf, 64, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        const double n = h[3] ? (double)h[3] : 1.0;
        fprintf(stderr, "oz timers: %llu CTAs, mean setup %.2f us, mainloop %.2f us, epilogue %.2f us; kernel span %.1f us\n",
                h[3], h[0] / n * 1e-3, h[1] / n * 1e-3, h[2] / n * 1e-3, (h[5] - h[4]) * 1e-3);
    }
    return rc;
}

}  // namespace b200
