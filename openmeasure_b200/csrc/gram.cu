// gram.cu -- K3: POD contraction of the tall-skinny snapshot matrix on the FP64 tensor path.
//
// Replaces the O(n m^2) part of np.linalg.svd(X0) (reference sparse_sensing.py:272, LAPACK dgesdd):
// per feature block f, Gf = sum_i (x_i - cnt_i)(x_i - cnt_i)^T over the block's rows, so the
// scaling 1/scl_f^2 (known only after the statistics pass) is applied to the tiny m x m result
// and the scaled X0 is never materialised (m > 64: the caller passes the centred copy X - cnt and cnt = NULL, so
// that the inner loops hold no FP64 add).  DMMA.8x8x4 (mma.sync m8n8k4 f64) -- tcgen05 has no FP64 kind.
// Deterministic: fixed work decomposition per (F, n_c, m, SM count), partial tiles reduced in fixed order, no atomics.
#include <stdlib.h>
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
// Many-snapshot variant (m > 64, even): 128 x 128 output tiles of the UPPER triangle, FP64 tensor
// path (DMMA.8x8x4) fed by bulk (TMA) copies through a 6-stage shared-memory ring of 16-row chunks
// (one copy per row and column panel; rows land at a pitch of 132 doubles, so the fragment loads
// are bank-conflict free).
//
//   * Work = every (feature, tile, 16-row chunk), linearised [feature][tile][chunk] and weighted by
//     its DMMA cost.  The P CTAs (one per SM) take CONTIGUOUS, equally expensive slices of that list
//     -- a slice may end in the middle of a tile's rows and continue with the next tile -- so no SM
//     idles in a last partial wave (m = 256, F = 9: 27 tiles on 148 SMs).  Every (CTA, tile) segment
//     leaves one partial tile in the workspace (slot = tile + CTA: injective because slices are
//     ordered); the reduction adds a tile's partials in CTA order.  Deterministic: the decomposition
//     depends only on (F, n_c, m, SM count); no atomics.
//   * Off-diagonal tiles: 8 warps x (32 x 64) outputs, 32 DMMAs per k-step and warp.
//     Diagonal tiles compute ONLY the 136 upper 8 x 8 blocks of their 16 x 16 block grid: warp w
//     owns block rows w and 15 - w (16 - w and w + 1 blocks = 17 DMMAs per k-step), i.e. 0.53 of an
//     off-diagonal tile instead of the full square (m = 256: 1.49x fewer DMMAs than square tiles).
//   * Centring happens ONCE per element, in shared memory, two chunks ahead of the multiplication:
//     DADD shares the FP64 pipe with DMMA and costs ~5.4 pipe cycles per warp instruction against
//     16 for a DMMA (tools/dmma_probe.cu: 12 DADDs per 32 DMMAs = -11 %; fragments are re-read by
//     2-4 warps, so centring them in registers cost 25 % of the pipe).  Each warp centres the two
//     rows it issued (raw rows land via TMA -> `full`; centred -> `ready`; consumed -> `empty`), and
//     zero-fills the rows of a ragged last chunk, so the k-step itself is LDS + DMMA, no masks.
//   * No producer warp: a ninth warp caps the register file at 168 per thread (3 warps on one
//     scheduler).  Every warp issues its rows 4 chunks ahead and centres 2 chunks ahead; with 6
//     stages a warp only needs the others to have finished the chunk two before its own.
//   * The first fragments of the next chunk are loaded during the last k-step of the current one.
// ---------------------------------------------------------------------------------------------
constexpr int GB_T = 128;                         // output tile edge
constexpr int GB_K = 16;                          // rows per stage
constexpr int GB_LD = GB_T + 4;                   // == 4 (mod 16)
constexpr int GB_STAGES = 6;
constexpr int GB_MMA_WARPS = 8;                   // two warpgroups of MMA warps ...
constexpr int GB_CENTRE_WARPS = 3;                // ... and a producer warpgroup: one copy warp + three centring warps
constexpr int GB_THREADS = (GB_MMA_WARPS + 1 + GB_CENTRE_WARPS) * 32;
constexpr int GB_META = 2 * GB_K * GB_LD + GB_K;  // rows | wa | wb as ints behind the centring values
constexpr int GB_STAGE_DOUBLES = GB_META + 2;      // panel A, panel B, centring values, meta
constexpr int GB_W_OFF = 256, GB_W_DIAG = 150;    // relative cost of a chunk: 32 vs 17 DMMAs per warp and k-step, and 19 vs 12 fragment loads (measured optimum)

static size_t gram_big_smem() { return sizeof(double) * GB_STAGES * GB_STAGE_DOUBLES; }

struct GbPlan {
    int T, ntiles, F, P;
    int w_off, w_diag;    // relative cost of an off-diagonal / a diagonal chunk
    int64_t nchunks;      // 16-row chunks per feature block
    int64_t Cf;           // cost of one feature block
    int64_t U;            // total cost
};

__host__ __device__ static inline GbPlan gb_plan(int64_t F, int64_t n_c, int64_t m, int sms)
{
    GbPlan p;
    p.T = (int)ceil_div(m, GB_T);
    p.ntiles = p.T * (p.T + 1) / 2;
    p.F = (int)F;
    p.nchunks = ceil_div(n_c, GB_K);
    p.w_off = GB_W_OFF;
    p.w_diag = GB_W_DIAG;
#ifndef __CUDA_ARCH__
    if (const char* e = getenv("OMB_GB_WDIAG")) { int v = atoi(e); if (v > 0 && v < 100000) p.w_diag = v; }
#endif
    p.Cf = p.nchunks * ((int64_t)p.T * p.w_diag + (int64_t)(p.ntiles - p.T) * p.w_off);
    p.U = p.Cf * F;
    int64_t P = sms;
    const int64_t total_chunks = p.nchunks * p.ntiles * F;
    if (P > total_chunks) P = total_chunks;
    if (P > 159) P = 159;                             // gram_big_reduce_kernel's slice table
    p.P = (int)(P < 1 ? 1 : P);
    return p;
}

// position u in the cost-weighted list -> (global tile g = f * ntiles + tile, chunk c); u == U -> (F * ntiles, 0)
__host__ __device__ static inline void gb_locate(const GbPlan& p, int64_t u, int& g, int64_t& c)
{
    const int64_t f = u / p.Cf;
    int64_t v = u - f * p.Cf;
    int t = 0;
    for (int ti = 0; ti < p.T; ++ti) {                // tiles row by row: (ti, ti) then (ti, ti+1 .. T-1)
        const int noff = p.T - 1 - ti;
        const int64_t rowcost = p.nchunks * (p.w_diag + (int64_t)noff * p.w_off);
        if (v >= rowcost) { v -= rowcost; t += 1 + noff; continue; }
        if (v < p.nchunks * p.w_diag) { c = v / p.w_diag; }
        else {
            v -= p.nchunks * p.w_diag;
            const int64_t k = v / (p.nchunks * p.w_off);
            t += 1 + (int)k;
            c = (v - k * p.nchunks * p.w_off) / p.w_off;
        }
        g = (int)f * p.ntiles + t;
        return;
    }
    g = (int)f * p.ntiles;                            // only reached for v == 0 at u == U
    c = 0;
}

__host__ __device__ static inline void gb_slice_start(const GbPlan& p, int b, int& g, int64_t& c)
{
    if (b >= p.P) { g = p.F * p.ntiles; c = 0; return; }
    gb_locate(p, (p.U * b) / p.P, g, c);
}

// one k-step (4 rows) of an off-diagonal tile for this warp: fragments of rows kr = 4 k4 + fr
struct GbFragOff { double a[4], b[8]; };

__device__ __forceinline__ void gb_load_off(GbFragOff& f, const double* sA, int k4, int fr, int fc, int ib, int jb)
{
    const double* ra = sA + (k4 * 4 + fr) * GB_LD + fc;
    const double* rb = ra + GB_K * GB_LD;
#pragma unroll
    for (int q = 0; q < 4; ++q) f.a[q] = ra[ib + q * 8];
#pragma unroll
    for (int q = 0; q < 8; ++q) f.b[q] = rb[jb + q * 8];
}

__device__ __forceinline__ void gb_mma_off(double (&c)[4][8][2], const GbFragOff& f)
{
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) dmma884(c[p][q][0], c[p][q][1], f.a[p], f.b[q]);
}

// diagonal tile: warp w owns block rows w (blocks w..15) and 15 - w (blocks 15-w..15) of the 16 x 16 grid of
// 8 x 8 blocks.  Accumulator t < 16 - w is block (w, w + t), the others block (15 - w, t - 1).
struct GbFragDiag { double alo, ahi, b[17]; };

__device__ __forceinline__ void gb_load_diag(GbFragDiag& f, const double* sA, int k4, int fr, int fc, int warp)
{
    const double* row = sA + (k4 * 4 + fr) * GB_LD + fc;
    const double* rlo = row + 8 * warp;              // column block of accumulator t: w + t (t < 16 - w), else t - 1
    const double* rhi = row - 8;
    const int nlo = 16 - warp;
    f.alo = row[8 * warp];
    f.ahi = row[8 * (15 - warp)];
#pragma unroll
    for (int t = 0; t < 17; ++t) f.b[t] = (t < nlo ? rlo : rhi)[8 * t];
}

__device__ __forceinline__ void gb_mma_diag(double (&c)[17][2], const GbFragDiag& f, int warp)
{
    const int nlo = 16 - warp;
#pragma unroll
    for (int t = 0; t < 17; ++t) dmma884(c[t][0], c[t][1], (t < nlo) ? f.alo : f.ahi, f.b[t]);
}

// cursor over the chunks of this CTA's slice
struct GbCursor {
    int g, gend, f, ti, tj;
    int c, cend_last, cstop;          // chunk cursor; cstop: end of the current segment
    __device__ __forceinline__ bool valid() const { return g < gend || (g == gend && c < cend_last); }
    __device__ __forceinline__ void set_stop(int nchunks) { cstop = (g == gend) ? cend_last : nchunks; }
    __device__ __forceinline__ void next_tile(int T, int nchunks)
    {
        ++g;
        if (++tj == T) { if (++ti == T) { ti = 0; ++f; } tj = ti; }
        c = 0;
        set_stop(nchunks);
    }
};

__device__ __forceinline__ void gb_cursor_init(GbCursor& cur, const GbPlan& plan, int b)
{
    int g0, g1;
    int64_t c0, c1;
    gb_slice_start(plan, b, g0, c0);
    gb_slice_start(plan, b + 1, g1, c1);
    cur.g = g0; cur.c = (int)c0; cur.gend = g1; cur.cend_last = (int)c1;
    cur.f = g0 / plan.ntiles;
    tile_pair(g0 - cur.f * plan.ntiles, plan.T, cur.ti, cur.tj);
    cur.set_stop((int)plan.nchunks);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(GB_THREADS, 1)
gram_big_kernel(const double* __restrict__ X, int64_t n_c, int m, const double* __restrict__ cnt, GbPlan plan,
                double* __restrict__ part)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[GB_STAGES], ready_bar[GB_STAGES], empty_bar[GB_STAGES];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = plan.T;
    const int nchunks = (int)plan.nchunks;

    // columns beyond m are never written by the copies: zero the whole ring once
    for (int e = threadIdx.x; e < GB_STAGES * GB_STAGE_DOUBLES; e += GB_THREADS) smem[e] = 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GB_STAGES; ++s) {
            mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], GB_CENTRE_WARPS); mbar_init(&empty_bar[s], GB_MMA_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncthreads();

    GbCursor cur;                       // every role walks the same chunk sequence: this CTA's slice of the work list
    gb_cursor_init(cur, plan, blockIdx.x);

    if (warp >= GB_MMA_WARPS) {
        // =========================== producer warpgroup (registers handed to the MMA warps) ===========
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == GB_MMA_WARPS) {
            // ---- copy warp: lane l < 16 owns row l of panel A, lane 16 + l row l of panel B
            const int rr = lane & (GB_K - 1);
            auto fetch_cnt = [&]() -> double {       // centring value of this lane's row, one chunk ahead
                if (!cnt || !cur.valid() || lane >= GB_K) return 0.0;
                const int64_t row = (int64_t)cur.c * GB_K + rr;
                return row < n_c ? cnt[(int64_t)cur.f * n_c + row] : 0.0;
            };
            double cnt_next = fetch_cnt();
            int s = 0;
            uint32_t ph = 0;                           // parity of the use of stage s being filled
            bool wrapped = false;
            while (cur.valid()) {
                if (wrapped) mbar_wait(&empty_bar[s], ph ^ 1);
                const bool diag = cur.ti == cur.tj;
                const int ca0 = cur.ti * GB_T, cb0 = cur.tj * GB_T;
                const int wa = (m - ca0) < GB_T ? (m - ca0) : GB_T;
                const int wb = diag ? 0 : ((m - cb0) < GB_T ? (m - cb0) : GB_T);
                const double* Xf = X + (int64_t)cur.f * n_c * m;
                const int64_t k0 = (int64_t)cur.c * GB_K;
                const int rows = (int)((n_c - k0) < GB_K ? (n_c - k0) : GB_K);
                double* stage = smem + (size_t)s * GB_STAGE_DOUBLES;
                if (lane < GB_K) stage[2 * GB_K * GB_LD + lane] = cnt_next;
                if (lane == 0) {
                    int* meta = reinterpret_cast<int*>(stage + GB_META);
                    meta[0] = rows; meta[1] = wa; meta[2] = wb;
                }
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(rows * (wa + wb) * sizeof(double)));
                __syncwarp();
                if (rr < rows) {
                    if (lane < GB_K) tma_load_bulk(stage + rr * GB_LD, Xf + (k0 + rr) * m + ca0, (uint32_t)(wa * sizeof(double)), &full_bar[s]);
                    else if (!diag) tma_load_bulk(stage + (GB_K + rr) * GB_LD, Xf + (k0 + rr) * m + cb0, (uint32_t)(wb * sizeof(double)), &full_bar[s]);
                }
                if (++cur.c >= cur.cstop) cur.next_tile(T, nchunks);
                cnt_next = fetch_cnt();
                if (++s == GB_STAGES) { s = 0; ph ^= 1; wrapped = true; }
            }
        } else {
            // ---- centring warps: x - cnt once per element (rows cw, cw + 3, ...); stale rows of a ragged chunk -> 0
            const int cw = warp - GB_MMA_WARPS - 1;
            int s = 0;
            uint32_t ph = 0;
            while (cur.valid()) {
                mbar_wait(&full_bar[s], ph);
                double* stage = smem + (size_t)s * GB_STAGE_DOUBLES;
                const int* meta = reinterpret_cast<const int*>(stage + GB_META);
                const int rows = meta[0], wa = meta[1], wb = meta[2];
                if (cnt != nullptr || rows < GB_K) {
#pragma unroll 1
                    for (int rr = cw; rr < GB_K; rr += GB_CENTRE_WARPS) {
                        const bool live = rr < rows;
                        const double cv = stage[2 * GB_K * GB_LD + rr];
                        double2* ra = reinterpret_cast<double2*>(stage + rr * GB_LD);
                        double2* rb = reinterpret_cast<double2*>(stage + (GB_K + rr) * GB_LD);
                        double2 v[4];
                        const int j0 = lane, j1 = lane + 32;        // double2 index: columns 2j, 2j + 1 (wa, wb even)
                        const bool a0 = 2 * j0 < wa, a1 = 2 * j1 < wa, b0 = 2 * j0 < wb, b1 = 2 * j1 < wb;
                        if (a0) v[0] = ra[j0];
                        if (a1) v[1] = ra[j1];
                        if (b0) v[2] = rb[j0];
                        if (b1) v[3] = rb[j1];
#pragma unroll
                        for (int t = 0; t < 4; ++t) { v[t].x = live ? v[t].x - cv : 0.0; v[t].y = live ? v[t].y - cv : 0.0; }
                        if (a0) ra[j0] = v[0];
                        if (a1) ra[j1] = v[1];
                        if (b0) rb[j0] = v[2];
                        if (b1) rb[j1] = v[3];
                    }
                    fence_proxy_async();        // generic writes before the stage's next bulk copy
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready_bar[s]);
                if (++cur.c >= cur.cstop) cur.next_tile(T, nchunks);
                if (++s == GB_STAGES) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // =============================== MMA warps: LDS + DMMA only ========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int fr = lane & 3, fc = lane >> 2;
    int s = 0;
    uint32_t ph = 0;
    while (cur.valid()) {
        const bool diag = cur.ti == cur.tj;
        double* out = part + ((int64_t)cur.g + blockIdx.x) * (GB_T * GB_T);
        if (!diag) {
            const int ib = (warp >> 1) * 32, jb = (warp & 1) * 64;
            double c[4][8][2];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) c[a][b][0] = c[a][b][1] = 0.0;
            GbFragOff f[2];
            mbar_wait(&ready_bar[s], ph);
            gb_load_off(f[0], smem + (size_t)s * GB_STAGE_DOUBLES, 0, fr, fc, ib, jb);
#pragma unroll 1
            for (; cur.c < cur.cstop; ++cur.c) {
                const double* sA = smem + (size_t)s * GB_STAGE_DOUBLES;
                const int s0 = s;
                const bool more = cur.c + 1 < cur.cstop;
                if (++s == GB_STAGES) { s = 0; ph ^= 1; }
#pragma unroll
                for (int k4 = 0; k4 < GB_K / 4; ++k4) {
                    if (k4 + 1 < GB_K / 4) gb_load_off(f[(k4 + 1) & 1], sA, k4 + 1, fr, fc, ib, jb);
                    else if (more) {                          // first fragments of the next chunk
                        mbar_wait(&ready_bar[s], ph);
                        gb_load_off(f[0], smem + (size_t)s * GB_STAGE_DOUBLES, 0, fr, fc, ib, jb);
                    }
                    gb_mma_off(c, f[k4 & 1]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s0]);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int i = ib + p * 8 + fc;
                    const int j = jb + q * 8 + 2 * fr;
                    *reinterpret_cast<double2*>(out + i * GB_T + j) = make_double2(c[p][q][0], c[p][q][1]);
                }
        } else {
            double c[17][2];
#pragma unroll
            for (int t = 0; t < 17; ++t) c[t][0] = c[t][1] = 0.0;
            GbFragDiag f[2];
            mbar_wait(&ready_bar[s], ph);
            gb_load_diag(f[0], smem + (size_t)s * GB_STAGE_DOUBLES, 0, fr, fc, warp);
#pragma unroll 1
            for (; cur.c < cur.cstop; ++cur.c) {
                const double* sA = smem + (size_t)s * GB_STAGE_DOUBLES;
                const int s0 = s;
                const bool more = cur.c + 1 < cur.cstop;
                if (++s == GB_STAGES) { s = 0; ph ^= 1; }
#pragma unroll
                for (int k4 = 0; k4 < GB_K / 4; ++k4) {
                    if (k4 + 1 < GB_K / 4) gb_load_diag(f[(k4 + 1) & 1], sA, k4 + 1, fr, fc, warp);
                    else if (more) {
                        mbar_wait(&ready_bar[s], ph);
                        gb_load_diag(f[0], smem + (size_t)s * GB_STAGE_DOUBLES, 0, fr, fc, warp);
                    }
                    gb_mma_diag(c, f[k4 & 1], warp);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s0]);
            }
#pragma unroll
            for (int t = 0; t < 17; ++t) {
                const int i = 8 * (t < 16 - warp ? warp : 15 - warp) + fc;
                const int j = 8 * (t + (t < 16 - warp ? warp : -1)) + 2 * fr;
                *reinterpret_cast<double2*>(out + i * GB_T + j) = make_double2(c[t][0], c[t][1]);
            }
        }
        cur.next_tile(T, nchunks);
    }
}

// Gf[f] = sum over the CTAs whose slice met the tile, in CTA order (fixed), of their partial tiles; only
// elements (i <= j) of the upper triangle are read and mirrored.
__global__ void __launch_bounds__(256)
gram_big_reduce_kernel(const double* __restrict__ part, int m, GbPlan plan, double* __restrict__ Gf)
{
    __shared__ int s_g[160];
    __shared__ int64_t s_c[160];
    __shared__ int s_list[160];
    __shared__ int s_n;
    const int t = blockIdx.x, f = blockIdx.y;
    const int g = f * plan.ntiles + t;
    for (int b = threadIdx.x; b <= plan.P; b += blockDim.x) gb_slice_start(plan, b, s_g[b], s_c[b]);
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int b = 0; b < plan.P; ++b) {
            if (s_g[b] > g || s_g[b + 1] < g) continue;
            const int64_t lo = (s_g[b] == g) ? s_c[b] : 0;
            const int64_t hi = (s_g[b + 1] == g) ? s_c[b + 1] : plan.nchunks;
            if (lo < hi) s_list[n++] = b;
        }
        s_n = n;
    }
    __syncthreads();
    int ti, tj;
    tile_pair(t, plan.T, ti, tj);
    const int nb = s_n;
    const int64_t mm = (int64_t)m * m;
    for (int e = blockIdx.z * blockDim.x + threadIdx.x; e < GB_T * GB_T; e += gridDim.z * blockDim.x) {
        const int il = e / GB_T, jl = e - il * GB_T;
        const int i = ti * GB_T + il, j = tj * GB_T + jl;
        if (i >= m || j >= m || i > j) continue;
        double s = 0.0;
        for (int k = 0; k < nb; ++k) s += part[((int64_t)g + s_list[k]) * (GB_T * GB_T) + e];
        Gf[(int64_t)f * mm + (int64_t)i * m + j] = s;
        Gf[(int64_t)f * mm + (int64_t)j * m + i] = s;
    }
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
    GramPlan p = gram_plan(F, n_c, m);
    GbPlan q = gb_plan(F, n_c, m, sm_count());
    const int64_t a = (int64_t)sizeof(double) * F * p.splits * p.ntiles * GT * GT;
    const int64_t b = (int64_t)sizeof(double) * ((int64_t)F * q.ntiles + q.P) * GB_T * GB_T;
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
        const GbPlan q = gb_plan(F, n_c, m, sm_count());
        OMB_CUDA(cudaFuncSetAttribute(gram_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gram_big_smem()));
        gram_big_kernel<<<(unsigned)q.P, GB_THREADS, gram_big_smem(), st>>>(d_X, n_c, (int)m, d_cnt, q, (double*)d_ws);
        rc = check_launch("gram_big_kernel");
        if (rc) return rc;
        dim3 rgrid((unsigned)q.ntiles, (unsigned)F, 8);
        gram_big_reduce_kernel<<<rgrid, 256, 0, st>>>((const double*)d_ws, (int)m, q, d_Gf);
        return check_launch("gram_big_reduce_kernel");
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
