// backproject.cu -- K5: U_r = X0 * (V_r Sigma_r^-1), centring/scaling fused into the read of X.
//
// Replaces the U = Q * U_R product inside LAPACK dgesdd behind np.linalg.svd (reference
// sparse_sensing.py:272) and the slice U[:, :r] (:336).  X0 = (X - cnt)/scl is formed on the fly
// (never materialised); the result is written tiled mode-major (common.cuh: Ut[tile][q][128]) -- the layout the placement
// kernels stream -- and the initial dgeqp3 column norms ||U_r[i, :]||_2 are produced in the same
// epilogue, summed sequentially over the modes with fma (the oracle's nrm2 order).
// Tensor path: DMMA.8x8x4 (mma.sync m8n8k4 f64).
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

constexpr int BP_RT = 64;        // rows per CTA tile
constexpr int BP_KC = 32;        // snapshots per shared-memory chunk
constexpr int BP_LDA = BP_KC + 4;   // == 4 (mod 16)
constexpr int BP_THREADS = 128;  // 4 warps x 16 rows
constexpr int BP_LDC = BP_RT + 2;   // == 2 (mod 8): conflict-free transposed accumulator spill

template <int QB>
struct BpSmem {
    static constexpr int QC = QB * 8;
    static constexpr int LDW = QC + 4;                       // == 4 (mod 16)
    static constexpr int AW = BP_RT * BP_LDA + BP_KC * LDW;  // doubles, A chunk + W chunk
    static constexpr int CT = QC * BP_LDC;                   // doubles, transposed output tile
    static constexpr int DOUBLES = (AW > CT ? AW : CT) + 2 * BP_RT;
    static constexpr size_t BYTES = sizeof(double) * DOUBLES;
};

template <int QB>
__global__ void __launch_bounds__(BP_THREADS)
backproject_kernel(const double* __restrict__ X, int64_t n, int64_t n_c, int m, const double* __restrict__ cnt,
                   const double* __restrict__ scl, const double* __restrict__ W, int r,
                   double* __restrict__ Ut, double* __restrict__ vn)
{
    using S = BpSmem<QB>;
    extern __shared__ double smem[];
    double* sA = smem;                         // [BP_RT][BP_LDA]
    double* sW = smem + BP_RT * BP_LDA;        // [BP_KC][LDW]
    double* sC = smem;                         // [QC][BP_LDC]   (aliases sA/sW after the k loop)
    double* s_cnt = smem + (S::DOUBLES - 2 * BP_RT);
    double* s_scl = s_cnt + BP_RT;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane & 3, fc = lane >> 2;
    const int ib = warp * 16;

    for (int64_t row0 = (int64_t)blockIdx.x * BP_RT; row0 < n; row0 += (int64_t)gridDim.x * BP_RT) {
        if (threadIdx.x < BP_RT) {
            int64_t row = row0 + threadIdx.x;
            double cv = 0.0, sv = 1.0;
            if (row < n) {
                if (cnt) cv = cnt[row];
                if (scl) sv = scl[row / n_c];
            }
            s_cnt[threadIdx.x] = cv;
            s_scl[threadIdx.x] = sv;
        }
        double nrm = 0.0;   // running sum of squares of row (row0 + threadIdx.x), threads < BP_RT
        for (int q0 = 0; q0 < r; q0 += S::QC) {
            double c[2][QB][2];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < QB; ++b) c[a][b][0] = c[a][b][1] = 0.0;

            for (int k0 = 0; k0 < m; k0 += BP_KC) {
                __syncthreads();   // previous consumers of sA/sW/sC are done; s_cnt visible
                for (int e = threadIdx.x; e < BP_RT * BP_KC; e += BP_THREADS) {
                    const int rr = e / BP_KC, kk = e - rr * BP_KC;
                    const int64_t row = row0 + rr;
                    const int col = k0 + kk;
                    double v = 0.0;
                    if (row < n && col < m) v = ldg_stream(X + row * m + col) - s_cnt[rr];
                    sA[rr * BP_LDA + kk] = v;
                }
                for (int e = threadIdx.x; e < BP_KC * S::QC; e += BP_THREADS) {
                    const int kk = e / S::QC, qq = e - kk * S::QC;
                    const int col = k0 + kk, q = q0 + qq;
                    sW[kk * S::LDW + qq] = (col < m && q < r) ? W[(int64_t)col * r + q] : 0.0;
                }
                __syncthreads();
#pragma unroll
                for (int k4 = 0; k4 < BP_KC / 4; ++k4) {
                    const double a0 = sA[(ib + fc) * BP_LDA + k4 * 4 + fr];
                    const double a1 = sA[(ib + 8 + fc) * BP_LDA + k4 * 4 + fr];
#pragma unroll
                    for (int b = 0; b < QB; ++b) {
                        const double bv = sW[(k4 * 4 + fr) * S::LDW + b * 8 + fc];
                        dmma884(c[0][b][0], c[0][b][1], a0, bv);
                        dmma884(c[1][b][0], c[1][b][1], a1, bv);
                    }
                }
            }
            __syncthreads();   // all warps done with sA/sW before they are overwritten by sC
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int i = ib + a * 8 + fc;
                const double isv = 1.0 / s_scl[i];
#pragma unroll
                for (int b = 0; b < QB; ++b) {
                    const int q = b * 8 + 2 * fr;
                    sC[q * BP_LDC + i] = c[a][b][0] * isv;
                    sC[(q + 1) * BP_LDC + i] = c[a][b][1] * isv;
                }
            }
            __syncthreads();
            const int qn = (r - q0) < S::QC ? (r - q0) : S::QC;
            for (int e = threadIdx.x; e < qn * BP_RT; e += BP_THREADS) {
                const int qq = e / BP_RT, i = e - qq * BP_RT;
                if (row0 + i < n) stg_stream(Ut + basis_index(q0 + qq, row0 + i, r), sC[qq * BP_LDC + i]);
            }
            if (vn && threadIdx.x < BP_RT) {
                for (int qq = 0; qq < qn; ++qq) {
                    const double u = sC[qq * BP_LDC + threadIdx.x];
                    nrm = fma(u, u, nrm);
                }
            }
        }
        if (vn && threadIdx.x < BP_RT && row0 + threadIdx.x < n) vn[row0 + threadIdx.x] = sqrt(nrm);
        __syncthreads();   // s_cnt/s_scl/sC reuse by the next row tile
    }
}

template <int QB>
static int launch_bp(const double* X, int64_t n, int64_t n_c, int m, const double* cnt, const double* scl,
                     const double* W, int r, double* Ut, double* vn, cudaStream_t st)
{
    using S = BpSmem<QB>;
    OMB_CUDA(cudaFuncSetAttribute(backproject_kernel<QB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)S::BYTES));
    int64_t grid = ceil_div(n, BP_RT);
    int64_t cap = (int64_t)sm_count() * 24;
    if (grid > cap) grid = cap;
    backproject_kernel<QB><<<(unsigned)grid, BP_THREADS, S::BYTES, st>>>(X, n, n_c, m, cnt, scl, W, r, Ut, vn);
    return check_launch("backproject_kernel");
}

// ---------------------------------------------------------------------------------------------
// Few-snapshot variant (m <= 64, r <= 64): one CTA = one basis tile of 128 rows, 8 warps x 16 rows.
// The A fragments (X - cnt, 8 rows x 4 snapshots = eight 32-byte sectors per load) are read
// straight from global memory -- a row's 8m bytes stay in L1 across its ceil(m/4) k-steps -- so X
// makes exactly one trip from HBM and there is no shared-memory staging or barrier in the main
// loop.  W is staged once per CTA.  The result leaves straight from the accumulator fragments (64-byte
// runs of the tile layout) and the dgeqp3 norms are reduced by shuffles: no barrier per tile.
// ---------------------------------------------------------------------------------------------
constexpr int BS_THREADS2 = 256;

template <int QB>
__global__ void __launch_bounds__(BS_THREADS2, 2)
backproject_small_kernel(const double* __restrict__ X, int64_t n, int64_t n_c, int m, const double* __restrict__ cnt,
                         const double* __restrict__ scl, const double* __restrict__ W, int r,
                         double* __restrict__ Ut, double* __restrict__ vn)
{
    constexpr int QC = QB * 8;
    constexpr int LDW = QC + 4;                    // == 4 (mod 16)
    extern __shared__ __align__(128) double smem[];
    double* sW = smem;                             // [mp][LDW]
    const int mp = (m + 3) & ~3;
    for (int e = threadIdx.x; e < mp * QC; e += BS_THREADS2) {
        const int k = e / QC, q = e - k * QC;
        sW[k * LDW + q] = (k < m && q < r) ? W[(int64_t)k * r + q] : 0.0;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane & 3, fc = lane >> 2;
    const int64_t ntiles = basis_tiles(n);
    constexpr int KU = 6;                      // k-steps per chunk: 2*KU loads per lane in flight
    const int nch = (mp + 4 * KU - 1) / (4 * KU);

    // software pipeline in registers: the raw X values of the next chunk (possibly of the next
    // tile) are requested before the current chunk is multiplied
    struct Rows { bool ok0, ok1; double c0, c1, s0, s1; const double* x0; const double* x1; };
    auto rows_of = [&](int64_t tile) {
        Rows R;
        const int64_t rowb = tile * OMB_TB + warp * 16;
        const int64_t r0 = rowb + fc, r1 = rowb + 8 + fc;
        R.ok0 = r0 < n; R.ok1 = r1 < n;
        R.c0 = (R.ok0 && cnt) ? cnt[r0] : 0.0; R.c1 = (R.ok1 && cnt) ? cnt[r1] : 0.0;
        R.s0 = (R.ok0 && scl) ? scl[r0 / n_c] : 1.0; R.s1 = (R.ok1 && scl) ? scl[r1 / n_c] : 1.0;
        R.x0 = X + r0 * m + fr; R.x1 = X + r1 * m + fr;
        return R;
    };
    auto fetch = [&](const Rows& R, int ch, double (&v0)[KU], double (&v1)[KU]) {
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            const int k0 = ch * 4 * KU + 4 * u;
            const bool kin = (k0 + fr) < m;
            v0[u] = (R.ok0 && kin) ? R.x0[k0] : R.c0;     // (value - cnt) == 0 outside the matrix
            v1[u] = (R.ok1 && kin) ? R.x1[k0] : R.c1;
        }
    };

    int64_t tile = blockIdx.x;
    Rows cur{};
    double v0[KU], v1[KU];
    if (tile < ntiles) { cur = rows_of(tile); fetch(cur, 0, v0, v1); }
    while (tile < ntiles) {
        double acc[2][QB][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < QB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        const int64_t ntile = tile + gridDim.x;
        Rows nxt = cur;
        for (int ch = 0; ch < nch; ++ch) {
            double w0[KU], w1[KU];
            const bool last = (ch + 1 == nch);
            if (!last) fetch(cur, ch + 1, w0, w1);
            else if (ntile < ntiles) { nxt = rows_of(ntile); fetch(nxt, 0, w0, w1); }
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                const int k0 = ch * 4 * KU + 4 * u;
                if (k0 < mp) {
                    const double a0 = v0[u] - cur.c0, a1 = v1[u] - cur.c1;
#pragma unroll
                    for (int b = 0; b < QB; ++b) {
                        const double bv = sW[(k0 + fr) * LDW + b * 8 + fc];
                        dmma884(acc[0][b][0], acc[0][b][1], a0, bv);
                        dmma884(acc[1][b][0], acc[1][b][1], a1, bv);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < KU; ++u) { v0[u] = w0[u]; v1[u] = w1[u]; }
        }
        // epilogue without staging or barriers: an accumulator fragment is eight consecutive candidates
        // of one mode = one 64-byte run of the tile; the warps never meet, so the epilogue of one warp
        // overlaps the DMMAs of the others (a staged tile + one bulk store cost two CTA barriers and a
        // 40-step norm loop per tile: the tensor pipe idled half of the time)
        {
            const double is0 = 1.0 / cur.s0, is1 = 1.0 / cur.s1;      // one reciprocal per row
            double* tb = Ut + tile * ((int64_t)r * OMB_TB) + warp * 16 + fc;
            double ss0 = 0.0, ss1 = 0.0;
#pragma unroll
            for (int b = 0; b < QB; ++b) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int q = b * 8 + 2 * fr + e;
                    if (q < r) {
                        const double u0 = acc[0][b][e] * is0, u1 = acc[1][b][e] * is1;
                        if (cur.ok0) stg_stream(tb + (int64_t)q * OMB_TB, u0);
                        if (cur.ok1) stg_stream(tb + (int64_t)q * OMB_TB + 8, u1);
                        ss0 = fma(u0, u0, ss0);
                        ss1 = fma(u1, u1, ss1);
                    }
                }
            }
            if (vn) {
                ss0 += __shfl_xor_sync(0xFFFFFFFFu, ss0, 1); ss0 += __shfl_xor_sync(0xFFFFFFFFu, ss0, 2);
                ss1 += __shfl_xor_sync(0xFFFFFFFFu, ss1, 1); ss1 += __shfl_xor_sync(0xFFFFFFFFu, ss1, 2);
                if (fr == 0) {
                    if (cur.ok0) vn[tile * OMB_TB + warp * 16 + fc] = sqrt(ss0);
                    if (cur.ok1) vn[tile * OMB_TB + warp * 16 + 8 + fc] = sqrt(ss1);
                }
            }
        }
        cur = nxt;
        tile = ntile;
    }
}

template <int QB>
static int launch_bp_small(const double* X, int64_t n, int64_t n_c, int m, const double* cnt, const double* scl,
                           const double* W, int r, double* Ut, double* vn, cudaStream_t st)
{
    const int mp = (m + 3) & ~3;
    const size_t bytes = sizeof(double) * ((size_t)mp * (QB * 8 + 4));
    OMB_CUDA(cudaFuncSetAttribute(backproject_small_kernel<QB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int64_t grid = basis_tiles(n);
    int64_t cap = (int64_t)sm_count() * 2;
    if (grid > cap) grid = cap;
    backproject_small_kernel<QB><<<(unsigned)grid, BS_THREADS2, bytes, st>>>(X, n, n_c, m, cnt, scl, W, r, Ut, vn);
    return check_launch("backproject_small_kernel");
}

// ---------------------------------------------------------------------------------------------
// Many-snapshot / many-mode variant (m or r > 64; m, r even): persistent, TMA-pipelined.
// One CTA per SM walks the 128-row basis tiles.  32-snapshot K-chunks of the tile's rows of X (one
// 256-byte bulk copy per row, padded pitch 36 doubles -> conflict-free fragment loads) and of W
// flow through a 3-stage ring; every warp issues its own share of a chunk's copies two chunks
// ahead (also across tile boundaries, so the epilogue of a tile overlaps the loads of the next).
// 8 warps = 4 row groups x 2 column halves, 32 x 8*QH outputs each (up to 128 accumulator
// registers per lane -- hence 256 threads, not a ninth producer warp: 288 threads cap ptxas at
// 168 registers).  The epilogue needs no staging: a DMMA accumulator fragment is eight
// consecutive candidates of one mode, i.e. one 64-byte run of the tiled mode-major layout.
// ---------------------------------------------------------------------------------------------
constexpr int BB_K = 32;                    // snapshots per stage
constexpr int BB_LDA = BB_K + 4;            // == 4 (mod 16)
constexpr int BB_STAGES = 3;
constexpr int BB_WARPS = 8;
constexpr int BB_THREADS = BB_WARPS * 32;

template <int QH>
struct BbCfg {
    static constexpr int QC = 16 * QH;                   // modes per launch
    static constexpr int LDW = QC + 4;                   // == 4 (mod 16)
    static constexpr int STAGE = OMB_TB * BB_LDA + BB_K * LDW;
    static constexpr size_t BYTES = sizeof(double) * ((size_t)BB_STAGES * STAGE + 2 * OMB_TB);
};

template <int QH>
__global__ void __launch_bounds__(BB_THREADS)
backproject_big_kernel(const double* __restrict__ X, int64_t n, int64_t n_c, int m, const double* __restrict__ cnt,
                       const double* __restrict__ scl, const double* __restrict__ W, int r, int q0, int first, int last,
                       double* __restrict__ Ut, double* __restrict__ vn)
{
    using Cfg = BbCfg<QH>;
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[BB_STAGES], empty_bar[BB_STAGES];
    double* s_n = smem + (size_t)BB_STAGES * Cfg::STAGE;         // [2][128] row sums of squares per column half

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = basis_tiles(n);
    const int nch = (m + BB_K - 1) / BB_K;
    const int qv = (r - q0) < Cfg::QC ? (r - q0) : Cfg::QC;      // valid modes of this launch (even)

    for (int e = threadIdx.x; e < BB_STAGES * Cfg::STAGE; e += BB_THREADS) smem[e] = 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < BB_STAGES; ++s) { mbar_init(&full_bar[s], BB_WARPS); mbar_init(&empty_bar[s], BB_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncthreads();

    // ---- issue side: this warp's share (16 rows of X, 4 rows of W) of chunk `gi`
    int64_t it_tile = blockIdx.x, gi = 0;
    int it_c = 0;
    auto issue = [&]() {
        if (it_tile >= ntiles) return;
        const int s = (int)(gi % BB_STAGES);
        if (gi >= BB_STAGES) mbar_wait(&empty_bar[s], (uint32_t)(((gi / BB_STAGES) - 1) & 1));
        const int64_t row0 = it_tile * OMB_TB;
        const int rows = (int)((n - row0) < OMB_TB ? (n - row0) : OMB_TB);
        const int k0 = it_c * BB_K;
        const int kv = (m - k0) < BB_K ? (m - k0) : BB_K;
        int nx = rows - 16 * warp; nx = nx < 0 ? 0 : (nx > 16 ? 16 : nx);
        int nw = kv - 4 * warp; nw = nw < 0 ? 0 : (nw > 4 ? 4 : nw);
        double* sA = smem + (size_t)s * Cfg::STAGE;
        double* sW = sA + OMB_TB * BB_LDA;
        if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(((int64_t)nx * kv + (int64_t)nw * qv) * sizeof(double)));
        __syncwarp();
        if (lane < nx) {
            const int rr = 16 * warp + lane;
            tma_load_bulk(sA + rr * BB_LDA, X + (row0 + rr) * m + k0, (uint32_t)(kv * sizeof(double)), &full_bar[s]);
        } else if (lane >= 16 && lane - 16 < nw) {
            const int kk = 4 * warp + lane - 16;
            tma_load_bulk(sW + kk * Cfg::LDW, W + (int64_t)(k0 + kk) * r + q0, (uint32_t)(qv * sizeof(double)), &full_bar[s]);
        }
        ++gi;
        if (++it_c == nch) { it_c = 0; it_tile += gridDim.x; }
    };
#pragma unroll 1
    for (int u = 0; u < BB_STAGES - 1; ++u) issue();

    // ---- compute side: warp (wr, wc) owns rows [32 wr, +32) x modes [8 QH wc, +8 QH)
    const int fr = lane & 3, fc = lane >> 2;
    const int wr = warp >> 1, wc = warp & 1;
    const int ib = wr * 32, jb = wc * 8 * QH;
    int64_t g = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * OMB_TB;
        double cv[4], sv[4];
        bool rok[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int64_t row = row0 + ib + 8 * p + fc;
            rok[p] = row < n;
            cv[p] = (rok[p] && cnt) ? cnt[row] : 0.0;
            sv[p] = (rok[p] && scl) ? scl[row / n_c] : 1.0;
        }
        double acc[4][QH][2];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < QH; ++q) acc[p][q][0] = acc[p][q][1] = 0.0;

#pragma unroll 1
        for (int c = 0; c < nch; ++c, ++g) {
            issue();                                           // chunk g + 2 into the stage chunk g - 1 used
            const int s = (int)(g % BB_STAGES);
            mbar_wait(&full_bar[s], (uint32_t)((g / BB_STAGES) & 1));
            const double* sA = smem + (size_t)s * Cfg::STAGE;
            const double* sW = sA + OMB_TB * BB_LDA;
            const int kv = (m - c * BB_K) < BB_K ? (m - c * BB_K) : BB_K;
#pragma unroll
            for (int k4 = 0; k4 < BB_K / 4; ++k4) {
                const int kk = k4 * 4 + fr;
                const bool kok = kk < kv;                      // stale snapshots of a ragged last chunk
                double a[4], b[QH];
#pragma unroll
                for (int p = 0; p < 4; ++p) a[p] = sA[(ib + 8 * p + fc) * BB_LDA + kk];
#pragma unroll
                for (int q = 0; q < QH; ++q) b[q] = sW[kk * Cfg::LDW + jb + 8 * q + fc];
#pragma unroll
                for (int p = 0; p < 4; ++p) a[p] = kok ? a[p] - cv[p] : 0.0;
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < QH; ++q) dmma884(acc[p][q][0], acc[p][q][1], a[p], b[q]);
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
        }

        // epilogue: U = acc / scl straight to the tile (64-byte runs), row sums of squares for the norms
        double* tbase = Ut + tile * ((int64_t)r * OMB_TB);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int i = ib + 8 * p + fc;
            const double isv = 1.0 / sv[p];                    // one division per row, not per element
            double ss = 0.0;
#pragma unroll
            for (int q = 0; q < QH; ++q) {
                const int col = jb + 8 * q + 2 * fr;           // mode index inside this launch
                const double u0 = acc[p][q][0] * isv, u1 = acc[p][q][1] * isv;
                if (col < qv) { if (rok[p]) stg_stream(tbase + (int64_t)(q0 + col) * OMB_TB + i, u0); ss = fma(u0, u0, ss); }
                if (col + 1 < qv) { if (rok[p]) stg_stream(tbase + (int64_t)(q0 + col + 1) * OMB_TB + i, u1); ss = fma(u1, u1, ss); }
            }
            ss += __shfl_xor_sync(0xFFFFFFFFu, ss, 1);
            ss += __shfl_xor_sync(0xFFFFFFFFu, ss, 2);
            if (vn && fr == 0) s_n[wc * OMB_TB + i] = ss;
        }
        if (vn) {
            __syncthreads();
            if (threadIdx.x < OMB_TB && row0 + threadIdx.x < n) {
                double t = s_n[threadIdx.x] + s_n[OMB_TB + threadIdx.x];
                double* dst = vn + row0 + threadIdx.x;
                if (!first) t += *dst;
                *dst = last ? sqrt(t) : t;
            }
            __syncthreads();
        }
    }
}

template <int QH>
static int launch_bp_big(const double* X, int64_t n, int64_t n_c, int m, const double* cnt, const double* scl,
                         const double* W, int r, int q0, double* Ut, double* vn, cudaStream_t st)
{
    using Cfg = BbCfg<QH>;
    OMB_CUDA(cudaFuncSetAttribute(backproject_big_kernel<QH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::BYTES));
    int64_t grid = basis_tiles(n);
    if (grid > sm_count()) grid = sm_count();
    const int first = q0 == 0, last = q0 + Cfg::QC >= r;
    backproject_big_kernel<QH><<<(unsigned)grid, BB_THREADS, Cfg::BYTES, st>>>(X, n, n_c, m, cnt, scl, W, r, q0, first, last, Ut, vn);
    return check_launch("backproject_big_kernel");
}

static int bp_big(const double* X, int64_t n, int64_t n_c, int m, const double* cnt, const double* scl, const double* W,
                  int r, double* Ut, double* vn, cudaStream_t st)
{
    // launches of up to 128 modes; the last one takes the narrowest instantiation that fits
    for (int q0 = 0; q0 < r; q0 += 128) {
        const int rem = r - q0;
        const int qh = rem >= 128 ? 8 : (rem + 15) / 16;
        int rc;
        switch (qh) {
#define OMB_BB(QHV) case QHV: rc = launch_bp_big<QHV>(X, n, n_c, m, cnt, scl, W, r, q0, Ut, vn, st); break;
            OMB_BB(1) OMB_BB(2) OMB_BB(3) OMB_BB(4) OMB_BB(5) OMB_BB(6) OMB_BB(7) OMB_BB(8)
#undef OMB_BB
            default: rc = -1; break;
        }
        if (rc) return rc;
    }
    return 0;
}

}  // namespace omb

using namespace omb;

extern "C" int omb_backproject(const double* d_X, int64_t F, int64_t n_c, int64_t m, const double* d_cnt,
                               const double* d_scl, const double* d_W, int64_t r, double* d_Ut, double* d_vn,
                               void* stream)
{
    OMB_CHECK_ARG(d_X && d_W && d_Ut, "null pointer");
    OMB_CHECK_ARG(F > 0 && n_c > 0 && m > 0 && r > 0, "non-positive size");
    const int64_t n = F * n_c;
    OMB_CHECK_ARG(m <= (1 << 20) && r <= (1 << 20), "m or r too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (m <= 64 && r <= 64) {
        const int qb = (int)((r + 7) / 8);
        switch (qb) {
#define OMB_BPS(QBV) case QBV: return launch_bp_small<QBV>(d_X, n, n_c, (int)m, d_cnt, d_scl, d_W, (int)r, d_Ut, d_vn, st);
            OMB_BPS(1) OMB_BPS(2) OMB_BPS(3) OMB_BPS(4) OMB_BPS(5) OMB_BPS(6) OMB_BPS(7) OMB_BPS(8)
#undef OMB_BPS
            default: break;
        }
    }
    if ((m & 1) == 0 && (r & 1) == 0 && ((reinterpret_cast<uintptr_t>(d_X) | reinterpret_cast<uintptr_t>(d_W)) & 15) == 0)
        return bp_big(d_X, n, n_c, (int)m, d_cnt, d_scl, d_W, (int)r, d_Ut, d_vn, st);
    if (r <= 64) return launch_bp<8>(d_X, n, n_c, (int)m, d_cnt, d_scl, d_W, (int)r, d_Ut, d_vn, st);
    return launch_bp<16>(d_X, n, n_c, (int)m, d_cnt, d_scl, d_W, (int)r, d_Ut, d_vn, st);
}
