// gram.cu -- K3: POD contraction of the tall-skinny snapshot matrix on the FP64 tensor path.
//
// Replaces the O(n m^2) part of np.linalg.svd(X0) (reference sparse_sensing.py:272, LAPACK dgesdd):
// per feature block f, Gf = sum_i (x_i - cnt_i)(x_i - cnt_i)^T over the block's rows, so the
// scaling 1/scl_f^2 (known only after the statistics pass) is applied to the tiny m x m result
// and X0 is never materialised.  DMMA.8x8x4 (mma.sync m8n8k4 f64) -- tcgen05 has no FP64 kind.
// Deterministic: fixed row split per feature, partial tiles reduced in fixed order, no atomics.
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

constexpr int GT = 64;            // output tile edge
constexpr int GK = 32;            // rows (contraction) per shared-memory chunk
constexpr int GLD = GT + 4;       // == 4 (mod 16): conflict-free 64-bit fragment loads
constexpr int G_THREADS = 128;    // 2 x 2 warps, 32 x 32 outputs each

struct GramPlan {
    int T;          // tiles per edge
    int ntiles;     // upper-triangular tile pairs
    int splits;     // row splits per feature
    int64_t rows_per_split;
};

static GramPlan gram_plan(int64_t F, int64_t n_c, int64_t m)
{
    GramPlan p;
    p.T = (int)ceil_div(m, GT);
    p.ntiles = p.T * (p.T + 1) / 2;
    int64_t target = (int64_t)sm_count() * 6;                  // CTAs wanted in flight
    int64_t splits = ceil_div(target, (int64_t)F * p.ntiles);
    int64_t max_splits = ceil_div(n_c, 4 * GK);                 // at least 4 chunks per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 4096) splits = 4096;
    p.rows_per_split = round_up(ceil_div(n_c, splits), GK);
    p.splits = (int)ceil_div(n_c, p.rows_per_split);
    return p;
}

__device__ __forceinline__ void tile_pair(int t, int T, int& ti, int& tj)
{
    // enumerate (ti <= tj) row by row
    ti = 0;
    int rowlen = T;
    while (t >= rowlen) { t -= rowlen; ++ti; --rowlen; }
    tj = ti + t;
}

__global__ void __launch_bounds__(G_THREADS)
gram_tile_kernel(const double* __restrict__ X, int64_t n_c, int m, const double* __restrict__ cnt, int T,
                 int64_t rows_per_split, int splits, double* __restrict__ part)
{
    __shared__ double sA[GK * GLD];
    __shared__ double sB[GK * GLD];
    __shared__ double s_cnt[GK];

    int ti, tj;
    tile_pair(blockIdx.x, T, ti, tj);
    const int split = blockIdx.y, f = blockIdx.z;
    const bool diag = (ti == tj);
    const int64_t row_lo = (int64_t)split * rows_per_split;
    int64_t row_hi = row_lo + rows_per_split;
    if (row_hi > n_c) row_hi = n_c;
    const double* Xf = X + (int64_t)f * n_c * m;
    const double* cf = cnt ? cnt + (int64_t)f * n_c : nullptr;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ib = (warp >> 1) * 32, jb = (warp & 1) * 32;
    const int fr = lane & 3, fc = lane >> 2;

    double c[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) c[a][b][0] = c[a][b][1] = 0.0;

    const double* sBp = diag ? sA : sB;
    for (int64_t k0 = row_lo; k0 < row_hi; k0 += GK) {
        if (threadIdx.x < GK) {
            int64_t row = k0 + threadIdx.x;
            s_cnt[threadIdx.x] = (cf && row < row_hi) ? cf[row] : 0.0;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < GK * GT; e += G_THREADS) {
            const int kr = e / GT, col = e - kr * GT;
            const int64_t row = k0 + kr;
            const bool rok = row < row_hi;
            const int ca = ti * GT + col;
            double va = 0.0;
            if (rok && ca < m) va = ldg_stream(Xf + row * m + ca) - s_cnt[kr];
            sA[kr * GLD + col] = va;
            if (!diag) {
                const int cb = tj * GT + col;
                double vb = 0.0;
                if (rok && cb < m) vb = ldg_stream(Xf + row * m + cb) - s_cnt[kr];
                sB[kr * GLD + col] = vb;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < GK / 4; ++k4) {
            double a[4], b[4];
            const int base = (k4 * 4 + fr) * GLD + fc;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = sA[base + ib + q * 8];
                b[q] = sBp[base + jb + q * 8];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) dmma884(c[p][q][0], c[p][q][1], a[p], b[q]);
        }
        __syncthreads();
    }

    // partial tile: part[f][split][tile][GT][GT]
    double* out = part + (((int64_t)f * splits + split) * gridDim.x + blockIdx.x) * (GT * GT);
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = ib + p * 8 + fc;
            const int j = jb + q * 8 + 2 * fr;
            *reinterpret_cast<double2*>(out + i * GT + j) = make_double2(c[p][q][0], c[p][q][1]);
        }
}

// ---------------------------------------------------------------------------------------------
// Many-snapshot variant (m > 64, even): 128 x 128 output tiles, warp-specialised.  A producer warp
// streams 16-row K-chunks of the two column panels of X through a 5-stage shared-memory ring with
// one bulk (TMA) copy per row and panel (rows land at a padded pitch of 132 doubles, so the
// fragment loads are bank-conflict free); 8 consumer warps own 32 x 64 outputs each (64 DMMA
// accumulators per lane).  A 64 x 64 tile needs 4.6 TB/s of operand traffic to keep the FP64
// tensor pipe busy, a 128 x 128 tile half of that -- it comes from the L2, where the tile pairs of
// a row split meet.  Centring is a DADD on the fragment (the raw tile is what TMA delivers).
// ---------------------------------------------------------------------------------------------
constexpr int GB_T = 128;                         // output tile edge
constexpr int GB_K = 16;                          // rows per stage
constexpr int GB_LD = GB_T + 4;                   // == 4 (mod 16)
constexpr int GB_STAGES = 5;
constexpr int GB_WARPS = 8;
constexpr int GB_THREADS = (GB_WARPS + 1) * 32;
constexpr int GB_STAGE_DOUBLES = 2 * GB_K * GB_LD + GB_K;      // panel A, panel B, centring values

static size_t gram_big_smem() { return sizeof(double) * GB_STAGES * GB_STAGE_DOUBLES; }

__global__ void __launch_bounds__(GB_THREADS, 1)
gram_big_kernel(const double* __restrict__ X, int64_t n_c, int m, const double* __restrict__ cnt, int T,
                int64_t rows_per_split, int splits, double* __restrict__ part)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[GB_STAGES], empty_bar[GB_STAGES];

    int ti, tj;
    tile_pair(blockIdx.x, T, ti, tj);
    const int split = blockIdx.y, f = blockIdx.z;
    const bool diag = (ti == tj);
    const int64_t row_lo = (int64_t)split * rows_per_split;
    int64_t row_hi = row_lo + rows_per_split;
    if (row_hi > n_c) row_hi = n_c;
    const double* Xf = X + (int64_t)f * n_c * m;
    const double* cf = cnt ? cnt + (int64_t)f * n_c : nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunks = (int)ceil_div(row_hi - row_lo, GB_K);
    const int ca0 = ti * GB_T, cb0 = tj * GB_T;
    const int wa = (m - ca0) < GB_T ? (m - ca0) : GB_T;       // valid columns of the two panels (even)
    const int wb = (m - cb0) < GB_T ? (m - cb0) : GB_T;

    // columns beyond m are never written by the copies: zero the whole ring once
    for (int e = threadIdx.x; e < GB_STAGES * GB_STAGE_DOUBLES; e += GB_THREADS) smem[e] = 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GB_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], GB_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncthreads();

    if (warp == GB_WARPS) {
        // ---------------- producer warp: lane l < 16 copies row l of panel A, lane 16 + l row l of panel B
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % GB_STAGES;
            if (c >= GB_STAGES) mbar_wait(&empty_bar[s], ((c / GB_STAGES) - 1) & 1);
            const int64_t k0 = row_lo + (int64_t)c * GB_K;
            const int rows = (int)((row_hi - k0) < GB_K ? (row_hi - k0) : GB_K);
            double* stage = smem + (size_t)s * GB_STAGE_DOUBLES;
            double* sCn = stage + 2 * GB_K * GB_LD;
            if (lane < GB_K) sCn[lane] = (cf && lane < rows) ? cf[k0 + lane] : 0.0;
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(rows * (wa + (diag ? 0 : wb)) * sizeof(double)));
            __syncwarp();
            const int rr = lane & (GB_K - 1);
            if (rr < rows) {
                if (lane < GB_K) tma_load_bulk(stage + rr * GB_LD, Xf + (k0 + rr) * m + ca0, (uint32_t)(wa * sizeof(double)), &full_bar[s]);
                else if (!diag) tma_load_bulk(stage + (GB_K + rr) * GB_LD, Xf + (k0 + rr) * m + cb0, (uint32_t)(wb * sizeof(double)), &full_bar[s]);
            }
        }
    } else {
        // ---------------- consumer warps: warp (wr, wc) owns G rows [32 wr, +32) x columns [64 wc, +64)
        const int fr = lane & 3, fc = lane >> 2;
        const int ib = (warp >> 1) * 32, jb = (warp & 1) * 64;
        double c[4][8][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) c[a][b][0] = c[a][b][1] = 0.0;

        for (int ch = 0; ch < nchunks; ++ch) {
            const int s = ch % GB_STAGES;
            mbar_wait(&full_bar[s], (ch / GB_STAGES) & 1);
            const double* sA = smem + (size_t)s * GB_STAGE_DOUBLES;
            const double* sB = diag ? sA : sA + GB_K * GB_LD;
            const double* sCn = sA + 2 * GB_K * GB_LD;
            const int rows = (int)((row_hi - (row_lo + (int64_t)ch * GB_K)) < GB_K ? (row_hi - (row_lo + (int64_t)ch * GB_K)) : GB_K);
#pragma unroll
            for (int k4 = 0; k4 < GB_K / 4; ++k4) {
                const int kr = k4 * 4 + fr;
                const bool kok = kr < rows;                // stale rows of a ragged last chunk
                const double cv = sCn[kr];
                double a[4], b[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] = sA[kr * GB_LD + ib + q * 8 + fc];
#pragma unroll
                for (int q = 0; q < 8; ++q) b[q] = sB[kr * GB_LD + jb + q * 8 + fc];
                const bool aok[4] = {ib + fc < wa, ib + 8 + fc < wa, ib + 16 + fc < wa, ib + 24 + fc < wa};
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] = (kok && aok[q]) ? a[q] - cv : 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) b[q] = (kok && (jb + q * 8 + fc) < wb) ? b[q] - cv : 0.0;
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 8; ++q) dmma884(c[p][q][0], c[p][q][1], a[p], b[q]);
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
        }

        // partial tile: part[f][split][tile][GB_T][GB_T]
        double* out = part + (((int64_t)f * splits + split) * gridDim.x + blockIdx.x) * (GB_T * GB_T);
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = ib + p * 8 + fc;
                const int j = jb + q * 8 + 2 * fr;
                *reinterpret_cast<double2*>(out + i * GB_T + j) = make_double2(c[p][q][0], c[p][q][1]);
            }
    }
}

static GramPlan gram_big_plan(int64_t F, int64_t n_c, int64_t m)
{
    GramPlan p;
    p.T = (int)ceil_div(m, GB_T);
    p.ntiles = p.T * (p.T + 1) / 2;
    // one CTA per SM (shared memory).  Few tiles: one full wave.  More tile CTAs than SMs: split the rows
    // so that the last wave is (nearly) full -- time ~ ceil(tiles * s / SMs) / s -- e.g. m = 1024, F = 9:
    // 324 tile CTAs = 2.19 waves would cost 3; 5 row splits give 11 waves of a fifth = 2.2
    const int64_t sms = sm_count(), tiles = (int64_t)F * p.ntiles;
    int64_t max_splits = ceil_div(n_c, 8 * GB_K);
    if (max_splits < 1) max_splits = 1;
    int64_t splits = sms / tiles;
    if (splits < 1) {
        double best = 1e30;
        splits = 1;
        for (int64_t sp = 1; sp <= 16 && sp <= max_splits; ++sp) {
            const double cost = (double)ceil_div(tiles * sp, sms) / (double)sp;
            if (cost < best * 0.97) { best = cost; splits = sp; }
        }
    }
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.rows_per_split = round_up(ceil_div(n_c, splits), GB_K);
    p.splits = (int)ceil_div(n_c, p.rows_per_split);
    return p;
}

static bool gram_big_ok(const double* X, int64_t m)
{
    return m > 64 && (m & 1) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
}

// ---------------------------------------------------------------------------------------------
// Few-snapshot variant (m <= 64): the whole m x m Gram fits one warp's accumulators, so every
// warp sweeps its own rows and keeps the NB(NB+1)/2 upper-triangular 8 x 8 blocks (NB = ceil(m/8))
// in registers.  A fragment X[k0 + lane%4][8b + lane/4] serves as both operands of the symmetric
// product.  A producer warp streams chunks of 4 * GS_WARPS rows through a 4-stage shared-memory ring
// with bulk (TMA) copies and mbarriers; consumer warp w takes rows 4w .. 4w+3 of every chunk, and
// -- when asked -- derives np.average(x, axis=1) of those rows from the same fragments (bit-exact:
// the fragment layout is numpy's accumulator layout).  X makes one trip from HBM for both results.
// Warps are combined in a fixed order through shared memory; the CTA's partial goes out in the
// 64 x 64 tile format of the staged kernel and is reduced by the same fixed-order pass.
// ---------------------------------------------------------------------------------------------
constexpr int GS_WARPS = 15;                      // consumer warps: warp w owns k-step w of every chunk
constexpr int GS_THREADS = (GS_WARPS + 1) * 32;   // + one producer warp driving the TMA ring
constexpr int GS_CH = 4 * GS_WARPS;               // rows per chunk
constexpr int GS_STAGES = 4;

template <int NB>
__global__ void __launch_bounds__(GS_THREADS)
gram_small_kernel(const double* __restrict__ X, int64_t n_c, int m, const double* __restrict__ cnt,
                  double* __restrict__ cnt_out, int64_t rows_per_split, int splits, double* __restrict__ part)
{
    // ring of GS_STAGES chunks: [GS_CH rows x m] of X followed by the chunk's GS_CH centring values
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[GS_STAGES], empty_bar[GS_STAGES];
    const int stage_doubles = ((GS_CH * m + 1) & ~1) + GS_CH;      // keeps every stage 16-byte aligned
    double* s_acc = smem + GS_STAGES * stage_doubles;              // [GT][GT] final combination

    const int split = blockIdx.x, f = blockIdx.y;
    const int64_t row_lo = (int64_t)split * rows_per_split;
    int64_t row_hi = row_lo + rows_per_split;
    if (row_hi > n_c) row_hi = n_c;
    const double* Xf = X + (int64_t)f * n_c * m;
    const double* cf = cnt ? cnt + (int64_t)f * n_c : nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunks = (int)ceil_div(row_hi - row_lo, GS_CH);

    if (threadIdx.x == 0) {
        for (int s = 0; s < GS_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], GS_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == GS_WARPS) {
        // ---------------- producer warp ----------------
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % GS_STAGES;
            if (c >= GS_STAGES) mbar_wait(&empty_bar[s], ((c / GS_STAGES) - 1) & 1);
            const int64_t k0 = row_lo + (int64_t)c * GS_CH;
            const int rows = (int)((row_hi - k0) < GS_CH ? (row_hi - k0) : GS_CH);
            double* dstX = smem + s * stage_doubles;
            double* dstC = dstX + ((GS_CH * m + 1) & ~1);
            const double* srcX = Xf + k0 * m;
            const uint32_t bytesX = (uint32_t)(rows * m * sizeof(double));
            const bool bulk_ok = ((reinterpret_cast<uintptr_t>(srcX) & 15) == 0) && ((bytesX & 15) == 0) &&
                                 (!cf || ((reinterpret_cast<uintptr_t>(cf + k0) & 15) == 0 && (rows & 1) == 0));
            if (bulk_ok) {
                if (lane == 0) {
                    const uint32_t bytesC = cf ? (uint32_t)(rows * sizeof(double)) : 0u;
                    mbar_expect_tx(&full_bar[s], bytesX + bytesC);
                    tma_load_bulk(dstX, srcX, bytesX, &full_bar[s]);
                    if (cf) tma_load_bulk(dstC, cf + k0, bytesC, &full_bar[s]);
                }
            } else {
                // ragged or unaligned chunk: the producer warp copies it itself
                for (int e = lane; e < rows * m; e += 32) dstX[e] = srcX[e];
                if (cf) for (int e = lane; e < rows; e += 32) dstC[e] = cf[k0 + e];
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        const int fr = lane & 3, fc = lane >> 2;
        double c[NB][NB][2];
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) c[a][b][0] = c[a][b][1] = 0.0;
        const int rowc = 4 * warp + fr;                             // this lane's row inside every chunk
        const int rem = m & 7;                                      // rem != 0: block NB-1 is the ragged octet
        const bool has_tail = rem != 0;
        const bool lastcol_ok = (8 * (NB - 1) + fc) < m;
        const double dm = (double)m;
        const int64_t myrow0 = row_lo + rowc;
        double* cnt_dst = cnt_out ? cnt_out + (int64_t)f * n_c + myrow0 : nullptr;
        const double* px0 = smem + rowc * m + fc;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int s = ch % GS_STAGES;
            mbar_wait(&full_bar[s], (ch / GS_STAGES) & 1);
            // unconditional loads (columns >= m read the next row / the slack behind the ring: always
            // inside shared memory), masked afterwards: no branches in the chunk loop
            const double* px = px0 + s * stage_doubles;
            double a[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) a[b] = px[8 * b];
            double cv = cf ? smem[s * stage_doubles + ((GS_CH * m + 1) & ~1) + rowc] : 0.0;
            if (cnt_out) {
                // np.average(x, axis=1) from the fragments already in registers: lane fc of a row's
                // 8 lanes holds numpy's accumulator fc (a[0] + a[1] + ... over the full octets), the
                // shuffles are numpy's ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the m % 8 tail.
                double sacc;
                if (NB == 1) {
                    sacc = has_tail ? -0.0 : a[0];
                } else {
                    sacc = a[0];
#pragma unroll
                    for (int b = 1; b < NB - 1; ++b) sacc += a[b];
                    if (!has_tail) sacc += a[NB - 1];
                }
                if (!(NB == 1 && has_tail)) {
                    sacc += __shfl_xor_sync(0xFFFFFFFFu, sacc, 4);
                    sacc += __shfl_xor_sync(0xFFFFFFFFu, sacc, 8);
                    sacc += __shfl_xor_sync(0xFFFFFFFFu, sacc, 16);
                }
                for (int e = 0; e < rem; ++e) sacc += __shfl_sync(0xFFFFFFFFu, a[NB - 1], fr + 4 * e);
                cv = sacc / dm;
                if (fc == 0 && myrow0 + (int64_t)ch * GS_CH < row_hi) cnt_dst[(int64_t)ch * GS_CH] = cv;
            }
#pragma unroll
            for (int b = 0; b < NB - 1; ++b) a[b] -= cv;
            a[NB - 1] = lastcol_ok ? a[NB - 1] - cv : 0.0;
            if (ch == nchunks - 1 && !(myrow0 + (int64_t)ch * GS_CH < row_hi)) {      // rows past the split's end
#pragma unroll
                for (int b = 0; b < NB; ++b) a[b] = 0.0;
            }
            __syncwarp();                                          // every lane holds its (used) values
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
#pragma unroll
            for (int bi = 0; bi < NB; ++bi)
#pragma unroll
                for (int bj = bi; bj < NB; ++bj) dmma884(c[bi][bj][0], c[bi][bj][1], a[bi], a[bj]);
        }

        // fixed-order combination of the consumer warps (named barrier 1: consumers only)
        for (int e = threadIdx.x; e < GT * GT; e += GS_WARPS * 32) s_acc[e] = 0.0;
        asm volatile("bar.sync 1, %0;" ::"n"(GS_WARPS * 32) : "memory");
        for (int w = 0; w < GS_WARPS; ++w) {
            if (warp == w) {
#pragma unroll
                for (int bi = 0; bi < NB; ++bi)
#pragma unroll
                    for (int bj = bi; bj < NB; ++bj) {
                        const int i = bi * 8 + fc, j = bj * 8 + 2 * fr;
                        s_acc[i * GT + j] += c[bi][bj][0];
                        s_acc[i * GT + j + 1] += c[bi][bj][1];
                    }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(GS_WARPS * 32) : "memory");
        }
        double* out = part + ((int64_t)f * splits + split) * (GT * GT);
        for (int e = threadIdx.x; e < GT * GT; e += GS_WARPS * 32) out[e] = s_acc[e];
    }
}

static size_t gram_small_smem(int m)
{
    const int stage_doubles = ((GS_CH * m + 1) & ~1) + GS_CH;
    return sizeof(double) * ((size_t)GS_STAGES * stage_doubles + GT * GT);
}

typedef void (*GramSmallFn)(const double*, int64_t, int, const double*, double*, int64_t, int, double*);
static GramSmallFn pick_gram_small(int m)
{
    switch ((m + 7) / 8) {
        case 1: return gram_small_kernel<1>;
        case 2: return gram_small_kernel<2>;
        case 3: return gram_small_kernel<3>;
        case 4: return gram_small_kernel<4>;
        case 5: return gram_small_kernel<5>;
        case 6: return gram_small_kernel<6>;
        case 7: return gram_small_kernel<7>;
        case 8: return gram_small_kernel<8>;
        default: return nullptr;
    }
}

static GramPlan gram_small_plan(int64_t F, int64_t n_c)
{
    GramPlan p;
    p.T = 1;
    p.ntiles = 1;
    // one CTA per SM fits (registers): a single full wave, never a straggler CTA
    int64_t splits = (int64_t)sm_count() / F;
    if (splits < 1) splits = 1;
    const int64_t unit = GS_CH;                                 // rows per chunk
    int64_t max_splits = ceil_div(n_c, 4 * unit);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.rows_per_split = round_up(ceil_div(n_c, splits), unit);
    p.splits = (int)ceil_div(n_c, p.rows_per_split);
    return p;
}

// Gf[f][i][j] = sum over splits (fixed order) of the tile holding (min-tile, max-tile); mirrored.
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const double* __restrict__ part, int m, int T, int ntiles, int splits, int gt,
                   double* __restrict__ Gf)
{
    const int f = blockIdx.y;
    const int64_t mm = (int64_t)m * m;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < mm;
         e += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(e / m), j = (int)(e - (int64_t)i * m);
        if (i > j) { int t = i; i = j; j = t; }          // read the upper triangle, mirror
        const int ti = i / gt, tj = j / gt;
        const int tile = ti * T - ti * (ti - 1) / 2 + (tj - ti);
        const double* p = part + ((int64_t)f * splits * ntiles + tile) * (gt * gt) + (i % gt) * gt + (j % gt);
        double s = 0.0;
        for (int sp = 0; sp < splits; ++sp) s += p[(int64_t)sp * ntiles * (gt * gt)];
        Gf[(int64_t)f * mm + e] = s;
    }
}

__global__ void __launch_bounds__(256)
gram_combine_kernel(const double* __restrict__ Gf, int F, int64_t mm, const double* __restrict__ scl,
                    double* __restrict__ G)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < mm;
         e += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int f = 0; f < F; ++f) {
            double w = 1.0;
            if (scl) { double sc = scl[f]; w = sc * sc; }
            s += Gf[(int64_t)f * mm + e] / w;
        }
        G[e] = s;
    }
}

}  // namespace omb

using namespace omb;

extern "C" int64_t omb_gram_ws_bytes(int64_t F, int64_t n_c, int64_t m)
{
    if (F <= 0 || n_c <= 0 || m <= 0) return 0;
    if (m <= 64) {
        GramPlan p = gram_small_plan(F, n_c);
        return (int64_t)sizeof(double) * F * p.splits * p.ntiles * GT * GT;
    }
    // the larger of the two many-snapshot layouts (which one runs depends on the alignment of X)
    GramPlan p = gram_plan(F, n_c, m), q = gram_big_plan(F, n_c, m);
    const int64_t a = (int64_t)sizeof(double) * F * p.splits * p.ntiles * GT * GT;
    const int64_t b = (int64_t)sizeof(double) * F * q.splits * q.ntiles * GB_T * GB_T;
    return a > b ? a : b;
}

namespace omb {
static int gram_impl(const double* d_X, int64_t F, int64_t n_c, int64_t m, const double* d_cnt, double* d_cnt_out,
                     double* d_Gf, void* d_ws, void* stream);
}

extern "C" int omb_gram(const double* d_X, int64_t F, int64_t n_c, int64_t m, const double* d_cnt,
                        double* d_Gf, void* d_ws, void* stream)
{
    return gram_impl(d_X, F, n_c, m, d_cnt, nullptr, d_Gf, d_ws, stream);
}

// Row means AND the centred per-feature Grams from one read of X (m <= 64: the Gram kernel derives
// np.average(x, axis=1) from the fragments it has in registers); larger m: two kernels.
extern "C" int omb_gram_rowmeans(const double* d_X, int64_t F, int64_t n_c, int64_t m, double* d_cnt_out,
                                 double* d_Gf, void* d_ws, void* stream)
{
    OMB_CHECK_ARG(d_cnt_out, "null pointer");
    if (m > 64) {
        int rc = omb_row_means(d_X, F * n_c, m, d_cnt_out, stream);
        if (rc) return rc;
        return gram_impl(d_X, F, n_c, m, d_cnt_out, nullptr, d_Gf, d_ws, stream);
    }
    return gram_impl(d_X, F, n_c, m, nullptr, d_cnt_out, d_Gf, d_ws, stream);
}

int omb::gram_impl(const double* d_X, int64_t F, int64_t n_c, int64_t m, const double* d_cnt, double* d_cnt_out,
                          double* d_Gf, void* d_ws, void* stream)
{
    OMB_CHECK_ARG(d_X && d_Gf && d_ws, "null pointer");
    OMB_CHECK_ARG(F > 0 && n_c > 0 && m > 0, "non-positive size");
    OMB_CHECK_ARG(F <= 65535 && m <= 16384, "F or m too large");
    cudaStream_t st = (cudaStream_t)stream;
    GramPlan p;
    int rc;
    int gt = GT;
    if (m <= 64) {
        p = gram_small_plan(F, n_c);
        dim3 grid((unsigned)p.splits, (unsigned)F);
        GramSmallFn fn = pick_gram_small((int)m);
        const size_t smem = gram_small_smem((int)m);
        OMB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fn<<<grid, GS_THREADS, smem, st>>>(d_X, n_c, (int)m, d_cnt, d_cnt_out, p.rows_per_split, p.splits, (double*)d_ws);
        rc = check_launch("gram_small_kernel");
    } else if (gram_big_ok(d_X, m)) {
        p = gram_big_plan(F, n_c, m);
        gt = GB_T;
        dim3 grid((unsigned)p.ntiles, (unsigned)p.splits, (unsigned)F);
        OMB_CUDA(cudaFuncSetAttribute(gram_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gram_big_smem()));
        gram_big_kernel<<<grid, GB_THREADS, gram_big_smem(), st>>>(d_X, n_c, (int)m, d_cnt, p.T, p.rows_per_split, p.splits,
                                                                  (double*)d_ws);
        rc = check_launch("gram_big_kernel");
    } else {
        p = gram_plan(F, n_c, m);
        dim3 grid((unsigned)p.ntiles, (unsigned)p.splits, (unsigned)F);
        gram_tile_kernel<<<grid, G_THREADS, 0, st>>>(d_X, n_c, (int)m, d_cnt, p.T, p.rows_per_split, p.splits,
                                                      (double*)d_ws);
        rc = check_launch("gram_tile_kernel");
    }
    if (rc) return rc;
    int64_t gx = ceil_div(m * m, 256);
    if (gx > 2048) gx = 2048;
    dim3 rgrid((unsigned)gx, (unsigned)F);
    gram_reduce_kernel<<<rgrid, 256, 0, st>>>((const double*)d_ws, (int)m, p.T, p.ntiles, p.splits, gt, d_Gf);
    return check_launch("gram_reduce_kernel");
}

extern "C" int omb_gram_combine(const double* d_Gf, int64_t F, int64_t m, const double* d_scl, double* d_G,
                                void* stream)
{
    OMB_CHECK_ARG(d_Gf && d_G, "null pointer");
    OMB_CHECK_ARG(F > 0 && m > 0, "non-positive size");
    int64_t gx = ceil_div(m * m, 256);
    if (gx > 2048) gx = 2048;
    gram_combine_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(d_Gf, (int)F, m * m, d_scl, d_G);
    return check_launch("gram_combine_kernel");
}
