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
// Many-snapshot / many-mode variant (m or r > 64; m, r even): persistent, warp-specialised, TMA-fed.
// One CTA per SM walks the 128-row basis tiles.  A producer warp streams 32-snapshot K-chunks of the
// tile's rows of X (one 256-byte bulk copy per row, padded pitch 36 doubles -> conflict-free fragment
// loads) and of W through a 3-stage ring; it belongs to a third warpgroup that hands its registers
// to the MMA warps (setmaxnreg), so those hold their accumulators AND double-buffered fragments
// without a producer's issue code on their critical path.
// 8 MMA warps x (16 rows x all 8*QB modes of the launch): every A fragment is loaded by exactly one
// warp (2 + QB fragment loads for 2*QB DMMAs per k-step), a row's modes all live in one warp -- the
// placement norms need no cross-warp step -- and 8*QB >= r is padded to the next multiple of 8, not
// 16 (r = 100: 104 columns computed instead of 112).
// CENTRE = false: X is the centred copy the Gram pass left behind (Engine._X0c), the k-step is
// LDS + DMMA only.  CENTRE = true (no room for the copy): x - cnt on the two A fragments per k-step.
// The epilogue needs no staging: a DMMA accumulator fragment is eight consecutive candidates of one
// mode, i.e. one 64-byte run of the tiled mode-major layout.
// ---------------------------------------------------------------------------------------------
constexpr int BB_K = 32;                    // snapshots per stage
constexpr int BB_LDA = BB_K + 4;            // == 4 (mod 16)
constexpr int BB_STAGES = 3;
constexpr int BB_MMA_WARPS = 8;
constexpr int BB_THREADS = (BB_MMA_WARPS + 4) * 32;

template <int QB>
struct BbCfg {
    static constexpr int QC = 8 * QB;                    // modes per launch
    static constexpr int LDW = QC + 4;                   // == 4 (mod 8): k rows 4 apart never share a bank pair
    static constexpr int STAGE = OMB_TB * BB_LDA + BB_K * LDW;
    static constexpr size_t BYTES = sizeof(double) * (size_t)BB_STAGES * STAGE;
};

template <int QB>
struct BbFrag { double a0, a1, b[QB]; };

template <int QB>
__device__ __forceinline__ void bb_load(BbFrag<QB>& f, const double* sA, const double* sW, int k4, int fr, int fc, int ib)
{
    const int kk = k4 * 4 + fr;
    f.a0 = sA[(ib + fc) * BB_LDA + kk];
    f.a1 = sA[(ib + 8 + fc) * BB_LDA + kk];
    const double* wr = sW + kk * BbCfg<QB>::LDW + fc;
#pragma unroll
    for (int q = 0; q < QB; ++q) f.b[q] = wr[8 * q];
}

template <int QB, bool CENTRE>
__global__ void __launch_bounds__(BB_THREADS, 1)
backproject_big_kernel(const double* __restrict__ X, int64_t n, int64_t n_c, int m, const double* __restrict__ cnt,
                       const double* __restrict__ scl, const double* __restrict__ W, int r, int q0, int first, int last,
                       double* __restrict__ Ut, double* __restrict__ vn)
{
    using Cfg = BbCfg<QB>;
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[BB_STAGES], empty_bar[BB_STAGES];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = basis_tiles(n);
    const int nch = (m + BB_K - 1) / BB_K;
    const int qv = (r - q0) < Cfg::QC ? (r - q0) : Cfg::QC;      // valid modes of this launch (even)

    // rows / snapshots / modes beyond the matrix are never written by the copies: zero the ring once
    for (int e = threadIdx.x; e < BB_STAGES * Cfg::STAGE; e += BB_THREADS) smem[e] = 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < BB_STAGES; ++s) { mbar_init(&full_bar[s], 4); mbar_init(&empty_bar[s], BB_MMA_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncthreads();

    if (warp >= BB_MMA_WARPS) {
        // =========================== producer warpgroup ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        // every bulk copy is its own (warp-serialised) instruction, ~60 cycles each: one warp cannot issue the 160
        // copies of a chunk in the time the MMA warps need for it, four can.  Producer warp p: rows [32p, 32p + 32)
        // of the tile (one per lane) and rows [8p, 8p + 8) of the W chunk.
        const int pw = warp - BB_MMA_WARPS;
        int s = 0;
        uint32_t ph = 0;
        bool wrapped = false;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t row0 = tile * OMB_TB;
            const int rows = (int)((n - row0) < OMB_TB ? (n - row0) : OMB_TB);
            int nx = rows - 32 * pw; nx = nx < 0 ? 0 : (nx > 32 ? 32 : nx);
            for (int c = 0; c < nch; ++c) {
                if (wrapped) mbar_wait(&empty_bar[s], ph ^ 1);
                const int k0 = c * BB_K;
                const int kv = (m - k0) < BB_K ? (m - k0) : BB_K;
                int nw = kv - 8 * pw; nw = nw < 0 ? 0 : (nw > 8 ? 8 : nw);
                double* sA = smem + (size_t)s * Cfg::STAGE;
                double* sW = sA + OMB_TB * BB_LDA;
                if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(((int64_t)nx * kv + (int64_t)nw * qv) * sizeof(double)));
                __syncwarp();
                if (lane < nx) {
                    const int rr = 32 * pw + lane;
                    tma_load_bulk(sA + rr * BB_LDA, X + (row0 + rr) * m + k0, (uint32_t)(kv * sizeof(double)), &full_bar[s]);
                }
                if (lane < nw) {
                    const int kk = 8 * pw + lane;
                    tma_load_bulk(sW + kk * Cfg::LDW, W + (int64_t)(k0 + kk) * r + q0, (uint32_t)(qv * sizeof(double)), &full_bar[s]);
                }
                if (++s == BB_STAGES) { s = 0; ph ^= 1; wrapped = true; }
            }
        }
        return;
    }

    // =============================== MMA warps ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int fr = lane & 3, fc = lane >> 2;
    const int ib = warp * 16;
    int s = 0;
    uint32_t ph = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * OMB_TB;
        const int64_t ra = row0 + ib + fc, rb = ra + 8;
        const bool oka = ra < n, okb = rb < n;
        double ca = 0.0, cb = 0.0;
        if (CENTRE) {
            ca = (oka && cnt) ? cnt[ra] : 0.0;
            cb = (okb && cnt) ? cnt[rb] : 0.0;
        }
        const double sa = (oka && scl) ? scl[ra / n_c] : 1.0, sb = (okb && scl) ? scl[rb / n_c] : 1.0;
        double acc[2][QB][2];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = 0; q < QB; ++q) acc[p][q][0] = acc[p][q][1] = 0.0;

        BbFrag<QB> f[2];
        mbar_wait(&full_bar[s], ph);
        bb_load<QB>(f[0], smem + (size_t)s * Cfg::STAGE, smem + (size_t)s * Cfg::STAGE + OMB_TB * BB_LDA, 0, fr, fc, ib);
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
            const double* sA = smem + (size_t)s * Cfg::STAGE;
            const double* sW = sA + OMB_TB * BB_LDA;
            const int s0 = s;
            const bool more = c + 1 < nch;
            const int kv = (m - c * BB_K) < BB_K ? (m - c * BB_K) : BB_K;
            if (++s == BB_STAGES) { s = 0; ph ^= 1; }
#pragma unroll
            for (int k4 = 0; k4 < BB_K / 4; ++k4) {
                if (k4 + 1 < BB_K / 4) bb_load<QB>(f[(k4 + 1) & 1], sA, sW, k4 + 1, fr, fc, ib);
                else if (more) {                              // first fragments of the next chunk
                    mbar_wait(&full_bar[s], ph);
                    bb_load<QB>(f[0], smem + (size_t)s * Cfg::STAGE, smem + (size_t)s * Cfg::STAGE + OMB_TB * BB_LDA, 0, fr, fc, ib);
                }
                BbFrag<QB>& g = f[k4 & 1];
                if (CENTRE) {
                    const bool kok = (k4 * 4 + fr) < kv;       // stale snapshots of a ragged last chunk (W rows are zero too)
                    g.a0 = kok ? g.a0 - ca : 0.0;
                    g.a1 = kok ? g.a1 - cb : 0.0;
                } else if (k4 * 4 + 3 >= kv) {                 // ragged last chunk: columns kv.. hold stale data
                    const bool kok = (k4 * 4 + fr) < kv;
                    g.a0 = kok ? g.a0 : 0.0;
                    g.a1 = kok ? g.a1 : 0.0;
                }
#pragma unroll
                for (int q = 0; q < QB; ++q) {
                    dmma884(acc[0][q][0], acc[0][q][1], g.a0, g.b[q]);
                    dmma884(acc[1][q][0], acc[1][q][1], g.a1, g.b[q]);
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s0])) : "memory");
        }

        // epilogue: U = acc / scl straight to the tile (64-byte runs); the row's sum of squares for the norms
        double* tbase = Ut + tile * ((int64_t)r * OMB_TB) + ib + fc;
        const double isa = 1.0 / sa, isb = 1.0 / sb;             // one division per row, not per element
        double ssa = 0.0, ssb = 0.0;
#pragma unroll
        for (int q = 0; q < QB; ++q) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = 8 * q + 2 * fr + e;               // mode index inside this launch
                if (col < qv) {
                    const double ua = acc[0][q][e] * isa, ub = acc[1][q][e] * isb;
                    double* dst = tbase + (int64_t)(q0 + col) * OMB_TB;
                    if (oka) stg_stream(dst, ua);
                    if (okb) stg_stream(dst + 8, ub);
                    ssa = fma(ua, ua, ssa);
                    ssb = fma(ub, ub, ssb);
                }
            }
        }
        if (vn) {
            ssa += __shfl_xor_sync(0xFFFFFFFFu, ssa, 1); ssa += __shfl_xor_sync(0xFFFFFFFFu, ssa, 2);
            ssb += __shfl_xor_sync(0xFFFFFFFFu, ssb, 1); ssb += __shfl_xor_sync(0xFFFFFFFFu, ssb, 2);
            if (fr == 0) {
                if (oka) { double t = ssa; if (!first) t += vn[ra]; vn[ra] = last ? sqrt(t) : t; }
                if (okb) { double t = ssb; if (!first) t += vn[rb]; vn[rb] = last ? sqrt(t) : t; }
            }
        }
    }
}

template <int QB, bool CENTRE>
static int launch_bp_big(const double* X, int64_t n, int64_t n_c, int m, const double* cnt, const double* scl,
                         const double* W, int r, int q0, double* Ut, double* vn, cudaStream_t st)
{
    using Cfg = BbCfg<QB>;
    OMB_CUDA(cudaFuncSetAttribute(backproject_big_kernel<QB, CENTRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::BYTES));
    int64_t grid = basis_tiles(n);
    if (grid > sm_count()) grid = sm_count();
    const int first = q0 == 0, last = q0 + Cfg::QC >= r;
    backproject_big_kernel<QB, CENTRE><<<(unsigned)grid, BB_THREADS, Cfg::BYTES, st>>>(X, n, n_c, m, cnt, scl, W, r, q0, first, last, Ut, vn);
    return check_launch("backproject_big_kernel");
}

template <bool CENTRE>
static int bp_big_t(const double* X, int64_t n, int64_t n_c, int m, const double* cnt, const double* scl, const double* W,
                    int r, double* Ut, double* vn, cudaStream_t st)
{
    // launches of up to 128 modes; the last one takes the narrowest instantiation that fits
    for (int q0 = 0; q0 < r; q0 += 128) {
        const int rem = r - q0;
        const int qb = rem >= 128 ? 16 : (rem + 7) / 8;
        int rc;
        switch (qb) {
#define OMB_BB(QBV) case QBV: rc = launch_bp_big<QBV, CENTRE>(X, n, n_c, m, cnt, scl, W, r, q0, Ut, vn, st); break;
            OMB_BB(1) OMB_BB(2) OMB_BB(3) OMB_BB(4) OMB_BB(5) OMB_BB(6) OMB_BB(7) OMB_BB(8)
            OMB_BB(9) OMB_BB(10) OMB_BB(11) OMB_BB(12) OMB_BB(13) OMB_BB(14) OMB_BB(15) OMB_BB(16)
#undef OMB_BB
            default: rc = -1; break;
        }
        if (rc) return rc;
    }
    return 0;
}

static int bp_big(const double* X, int64_t n, int64_t n_c, int m, const double* cnt, const double* scl, const double* W,
                  int r, double* Ut, double* vn, cudaStream_t st)
{
    return cnt ? bp_big_t<true>(X, n, n_c, m, cnt, scl, W, r, Ut, vn, st)
               : bp_big_t<false>(X, n, n_c, m, cnt, scl, W, r, Ut, vn, st);
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
