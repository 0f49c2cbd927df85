// ols.cu -- K8-K11: train (row gather), batched OLS predict, reconstruct with fused unscale,
// and the layout helpers between the mode-major basis and the reference's (n, r) view.
//
//   gather      replaces Theta = C.dot(Ur) and C.dot(X_cnt[:,0]) for a one-hot C
//               (reference sparse_sensing.py:797, :573): a gather of s rows.
//   predict     replaces the per-vector loop  pinv(Theta) @ ((y - cnt)/scl)  (:865-878, W = I)
//               by one (N x s)(s x r) FP64 tensor-core GEMM with the scaling fused on load.
//   reconstruct replaces X_rec = Ur @ Ar.T followed by the per-column unscale_data (:371-373, :235)
//               by a row-sharded GEMM whose epilogue is scl*acc + cnt (two roundings, like the
//               reference's multiply-then-add).
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

__global__ void gather_rows_kernel(const double* __restrict__ Ut, int r, const int64_t* __restrict__ piv,
                                   int s, double* __restrict__ Theta, const double* __restrict__ cnt,
                                   double* __restrict__ cnt_s)
{
    const int total = s * r;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int a = e / r, q = e - a * r;
        Theta[e] = Ut[basis_index(q, piv[a], r)];
    }
    if (cnt && cnt_s)
        for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < s; a += gridDim.x * blockDim.x)
            cnt_s[a] = cnt[piv[a]];
}

// Ur[i][q] = Ut[q][i]   (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256)
modes_to_rows_kernel(const double* __restrict__ Ut, int64_t n, int r, double* __restrict__ Ur)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    const int64_t ntile_i = ceil_div(n, 32);
    const int ntile_q = (r + 31) / 32;
    for (int64_t tI = blockIdx.x; tI < ntile_i * ntile_q; tI += gridDim.x) {
        const int64_t i0 = (tI / ntile_q) * 32;
        const int q0 = (int)(tI % ntile_q) * 32;
        for (int yy = ty; yy < 32; yy += 8) {
            const int q = q0 + yy;
            const int64_t i = i0 + tx;
            tile[yy][tx] = (q < r && i < n) ? Ut[basis_index(q, i, r)] : 0.0;
        }
        __syncthreads();
        for (int yy = ty; yy < 32; yy += 8) {
            const int64_t i = i0 + yy;
            const int q = q0 + tx;
            if (i < n && q < r) Ur[i * r + q] = tile[tx][yy];
        }
        __syncthreads();
    }
}

// Ut[q][i] = Ur[i][q]; optional row norms (sequential fma over q, the oracle's dnrm2 order)
__global__ void __launch_bounds__(256)
rows_to_modes_kernel(const double* __restrict__ Ur, int64_t n, int r, double* __restrict__ Ut)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t ntile_i = ceil_div(n, 32);
    const int ntile_q = (r + 31) / 32;
    for (int64_t tI = blockIdx.x; tI < ntile_i * ntile_q; tI += gridDim.x) {
        const int64_t i0 = (tI / ntile_q) * 32;
        const int q0 = (int)(tI % ntile_q) * 32;
        for (int yy = ty; yy < 32; yy += 8) {
            const int64_t i = i0 + yy;
            const int q = q0 + tx;
            tile[yy][tx] = (i < n && q < r) ? Ur[i * r + q] : 0.0;
        }
        __syncthreads();
        for (int yy = ty; yy < 32; yy += 8) {
            const int q = q0 + yy;
            const int64_t i = i0 + tx;
            if (q < r && i < n) Ut[basis_index(q, i, r)] = tile[tx][yy];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
row_norms_kernel(const double* __restrict__ Ut, int64_t n, int r, double* __restrict__ vn)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < r; ++k) { const double x = Ut[basis_index(k, j, r)]; s = fma(x, x, s); }
        vn[j] = sqrt(s);
    }
}

// ---------------------------------------------------------------------------------------------
// 64 x 64 output tile, 2 x 2 warps of 32 x 32, DMMA.8x8x4, contraction chunks of 32.
//   MODE 0 (predict):     C[i][j] = sum_k ((Y[i][k] - cs[k]) / ss[k]) * B[k][j]        row-major out
//   MODE 1 (reconstruct): C[i][j] = scl[f(i)] * (sum_k Ut[k][i] * Ac[j][k]) + cnt[i]   row-major out
// ---------------------------------------------------------------------------------------------
constexpr int OT = 64, OK_ = 32, O_THREADS = 128;
constexpr int O_LD = OT + 4;      // [k][i] operand tiles: fr*LD + fc distinct (mod 16)
constexpr int O_LDK = OK_ + 4;    // [i][k] operand tiles: fc*LD + fr distinct (mod 16)

template <int MODE>
__global__ void __launch_bounds__(O_THREADS)
ols_gemm_kernel(const double* __restrict__ A, const double* __restrict__ B, int64_t M, int64_t Nn, int K,
                int64_t lda, const double* __restrict__ p0, const double* __restrict__ p1, int64_t n_c,
                int64_t row0, double* __restrict__ out)
{
    __shared__ double sA[(MODE == 0) ? OT * O_LDK : OK_ * O_LD];
    __shared__ double sB[(MODE == 0) ? OK_ * O_LD : OT * O_LDK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane & 3, fc = lane >> 2;
    const int ib = (warp >> 1) * 32, jb = (warp & 1) * 32;
    const int64_t tiles_j = ceil_div(Nn, OT);
    const int64_t tiles = ceil_div(M, OT) * tiles_j;

    for (int64_t tI = blockIdx.x; tI < tiles; tI += gridDim.x) {
        const int64_t i0 = (tI / tiles_j) * OT, j0 = (tI % tiles_j) * OT;
        double c[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) c[a][b][0] = c[a][b][1] = 0.0;

        for (int k0 = 0; k0 < K; k0 += OK_) {
            __syncthreads();
            if (MODE == 0) {
                // sA[i][k] = (Y[i0+i][k0+k] - cs[k]) / ss[k];  sB[k][j] = PinvT[k0+k][j0+j]
                for (int e = threadIdx.x; e < OT * OK_; e += O_THREADS) {
                    const int ii = e / OK_, kk = e - ii * OK_;
                    const int64_t i = i0 + ii;
                    const int k = k0 + kk;
                    double v = 0.0;
                    if (i < M && k < K) {
                        v = A[i * lda + k];
                        if (p0) v = v - p0[k];
                        if (p1) v = v / p1[k];
                    }
                    sA[ii * O_LDK + kk] = v;
                }
                for (int e = threadIdx.x; e < OK_ * OT; e += O_THREADS) {
                    const int kk = e / OT, jj = e - kk * OT;
                    const int k = k0 + kk;
                    const int64_t j = j0 + jj;
                    sB[kk * O_LD + jj] = (k < K && j < Nn) ? B[(int64_t)k * Nn + j] : 0.0;
                }
            } else {
                // sA[k][i] = Ut[k0+k][row0+i0+i];  sB[j][k] = Ac[j0+j][k0+k]
                for (int e = threadIdx.x; e < OK_ * OT; e += O_THREADS) {
                    const int kk = e / OT, ii = e - kk * OT;
                    const int k = k0 + kk;
                    const int64_t i = i0 + ii;
                    sA[kk * O_LD + ii] = (k < K && i < M) ? A[basis_index(k, row0 + i, K)] : 0.0;
                }
                for (int e = threadIdx.x; e < OT * OK_; e += O_THREADS) {
                    const int jj = e / OK_, kk = e - jj * OK_;
                    const int64_t j = j0 + jj;
                    const int k = k0 + kk;
                    sB[jj * O_LDK + kk] = (j < Nn && k < K) ? B[j * K + k] : 0.0;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k4 = 0; k4 < OK_ / 4; ++k4) {
                double a[4], b[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (MODE == 0) {
                        a[q] = sA[(ib + q * 8 + fc) * O_LDK + k4 * 4 + fr];
                        b[q] = sB[(k4 * 4 + fr) * O_LD + jb + q * 8 + fc];
                    } else {
                        a[q] = sA[(k4 * 4 + fr) * O_LD + ib + q * 8 + fc];
                        b[q] = sB[(jb + q * 8 + fc) * O_LDK + k4 * 4 + fr];
                    }
                }
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) dmma884(c[p][q][0], c[p][q][1], a[p], b[q]);
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int64_t i = i0 + ib + p * 8 + fc;
            if (i >= M) continue;
            double sc = 1.0, cn = 0.0;
            if (MODE == 1) {
                if (p1) sc = p1[(row0 + i) / n_c];
                if (p0) cn = p0[row0 + i];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t j = j0 + jb + q * 8 + 2 * fr + e;
                    if (j >= Nn) continue;
                    double v = c[p][q][e];
                    if (MODE == 1) { v = sc * v; v = v + cn; }
                    out[i * Nn + j] = v;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Reconstruct, warp-specialised and TMA-fed (even r, even N, row range aligned to basis tiles): persistent
// CTAs walk 128 x 128 output tiles (one basis tile of candidates x 128 coefficient vectors).  A 32-mode
// K-chunk of the basis tile is 32 contiguous 1 KB rows of the tiled layout (one bulk copy each, pitch
// 132 doubles), the matching chunk of the coefficient rows 128 copies of 256 bytes (pitch 36): both
// operands land bank-conflict free.  The four warps of a producer warpgroup issue the copies of a chunk
// (a bulk copy is a warp-serialised instruction of ~60 cycles: one warp cannot feed the ring) and hand
// their registers to the 8 MMA warps (setmaxnreg), which own 32 x 64 outputs each (64 DMMA accumulators
// per lane, double-buffered fragments) and run LDS + DMMA only; the first fragments of the next chunk are
// loaded during the last k-step of the current one.  Epilogue scl * acc + cnt (two roundings, like the
// reference's multiply-then-add) as 16-byte stores.  FP64-tensor bound: 2 n r N flop against 8 n N
// written bytes (r/4 flop per byte).
// ---------------------------------------------------------------------------------------------
constexpr int RB_K = 32;
constexpr int RB_LDA = OMB_TB + 4;          // [k][i]: == 4 (mod 16)
constexpr int RB_LDB = RB_K + 4;            // [j][k]: == 4 (mod 16)
constexpr int RB_STAGES = 3;
constexpr int RB_MMA_WARPS = 8;
constexpr int RB_THREADS = (RB_MMA_WARPS + 4) * 32;
constexpr int RB_STAGE = RB_K * RB_LDA + OMB_TB * RB_LDB;
static size_t reconstruct_big_smem() { return sizeof(double) * (size_t)RB_STAGES * RB_STAGE; }

struct RbFrag { double a[4], b[8]; };

__device__ __forceinline__ void rb_load(RbFrag& f, const double* sA, const double* sB, int k4, int fr, int fc, int ib, int jb)
{
    const int kk = k4 * 4 + fr;
    const double* ra = sA + kk * RB_LDA + ib + fc;
    const double* rb = sB + (jb + fc) * RB_LDB + kk;
#pragma unroll
    for (int p = 0; p < 4; ++p) f.a[p] = ra[8 * p];
#pragma unroll
    for (int q = 0; q < 8; ++q) f.b[q] = rb[8 * q * RB_LDB];
}

__global__ void __launch_bounds__(RB_THREADS, 1)
reconstruct_big_kernel(const double* __restrict__ Ut, int r, const double* __restrict__ Ac, int64_t N,
                       const double* __restrict__ cnt, const double* __restrict__ scl, int64_t n_c, int64_t row0,
                       int64_t nrows, double* __restrict__ out)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[RB_STAGES], empty_bar[RB_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tiles_i = ceil_div(nrows, (int64_t)OMB_TB), tiles_j = ceil_div(N, (int64_t)OMB_TB);
    const int64_t ntiles = tiles_i * tiles_j;
    const int nch = (r + RB_K - 1) / RB_K;
    const int64_t tile0 = row0 / OMB_TB;                     // row0 is a multiple of 128

    for (int e = threadIdx.x; e < RB_STAGES * RB_STAGE; e += RB_THREADS) smem[e] = 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < RB_STAGES; ++s) { mbar_init(&full_bar[s], 4); mbar_init(&empty_bar[s], RB_MMA_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncthreads();

    if (warp >= RB_MMA_WARPS) {
        // =========================== producer warpgroup: warp p copies mode rows [8p, 8p + 8) of the basis
        // chunk and coefficient rows [32p, 32p + 32) ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int pw = warp - RB_MMA_WARPS;
        int s = 0;
        uint32_t ph = 0;
        bool wrapped = false;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t ti = tile / tiles_j, tj = tile - ti * tiles_j;
            const int64_t j0 = tj * OMB_TB;
            const int jv = (int)((N - j0) < OMB_TB ? (N - j0) : OMB_TB);
            int nb = jv - 32 * pw; nb = nb < 0 ? 0 : (nb > 32 ? 32 : nb);
            const double* tb = Ut + (tile0 + ti) * ((int64_t)r * OMB_TB);
            for (int c = 0; c < nch; ++c) {
                if (wrapped) mbar_wait(&empty_bar[s], ph ^ 1);
                const int k0 = c * RB_K;
                const int kv = (r - k0) < RB_K ? (r - k0) : RB_K;
                int na = kv - 8 * pw; na = na < 0 ? 0 : (na > 8 ? 8 : na);
                double* sA = smem + (size_t)s * RB_STAGE;
                double* sB = sA + RB_K * RB_LDA;
                if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(((int64_t)na * OMB_TB + (int64_t)nb * kv) * sizeof(double)));
                __syncwarp();
                if (lane < na) {
                    const int kk = 8 * pw + lane;
                    tma_load_bulk(sA + kk * RB_LDA, tb + (int64_t)(k0 + kk) * OMB_TB, OMB_TB * sizeof(double), &full_bar[s]);
                }
                if (lane < nb) {
                    const int jj = 32 * pw + lane;
                    tma_load_bulk(sB + jj * RB_LDB, Ac + (j0 + jj) * r + k0, (uint32_t)(kv * sizeof(double)), &full_bar[s]);
                }
                if (++s == RB_STAGES) { s = 0; ph ^= 1; wrapped = true; }
            }
        }
        return;
    }

    // =============================== MMA warps ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int fr = lane & 3, fc = lane >> 2;
    const int ib = (warp >> 1) * 32, jb = (warp & 1) * 64;
    int s = 0;
    uint32_t ph = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t ti = tile / tiles_j, tj = tile - ti * tiles_j;
        double acc[4][8][2];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[p][q][0] = acc[p][q][1] = 0.0;
        RbFrag f[2];
        mbar_wait(&full_bar[s], ph);
        rb_load(f[0], smem + (size_t)s * RB_STAGE, smem + (size_t)s * RB_STAGE + RB_K * RB_LDA, 0, fr, fc, ib, jb);
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
            const double* sA = smem + (size_t)s * RB_STAGE;
            const double* sB = sA + RB_K * RB_LDA;
            const int s0 = s;
            const bool more = c + 1 < nch;
            const int kv = (r - c * RB_K) < RB_K ? (r - c * RB_K) : RB_K;
            if (++s == RB_STAGES) { s = 0; ph ^= 1; }
#pragma unroll
            for (int k4 = 0; k4 < RB_K / 4; ++k4) {
                if (k4 + 1 < RB_K / 4) rb_load(f[(k4 + 1) & 1], sA, sB, k4 + 1, fr, fc, ib, jb);
                else if (more) {
                    mbar_wait(&full_bar[s], ph);
                    rb_load(f[0], smem + (size_t)s * RB_STAGE, smem + (size_t)s * RB_STAGE + RB_K * RB_LDA, 0, fr, fc, ib, jb);
                }
                RbFrag& g = f[k4 & 1];
                if (k4 * 4 + 3 >= kv) {                        // ragged last chunk: modes kv.. hold stale data
                    const bool kok = (k4 * 4 + fr) < kv;
#pragma unroll
                    for (int p = 0; p < 4; ++p) g.a[p] = kok ? g.a[p] : 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) g.b[q] = kok ? g.b[q] : 0.0;
                }
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 8; ++q) dmma884(acc[p][q][0], acc[p][q][1], g.a[p], g.b[q]);
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s0])) : "memory");
        }
        // epilogue: x = scl * acc + cnt, 16-byte stores (N even, j even)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int64_t i = ti * OMB_TB + ib + 8 * p + fc;         // row inside the requested range
            if (i >= nrows) continue;
            const double sc = scl ? scl[(row0 + i) / n_c] : 1.0;
            const double cn = cnt ? cnt[row0 + i] : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t j = tj * OMB_TB + jb + 8 * q + 2 * fr;
                if (j >= N) continue;
                double v0 = sc * acc[p][q][0], v1 = sc * acc[p][q][1];
                v0 = v0 + cn; v1 = v1 + cn;
                stg_stream2(out + i * N + j, make_double2(v0, v1));
            }
        }
    }
}

}  // namespace omb

using namespace omb;

extern "C" int omb_gather_rows(const double* d_Ut, int64_t r, const int64_t* d_piv, int64_t s,
                               double* d_Theta, const double* d_cnt, double* d_cnt_s, void* stream)
{
    OMB_CHECK_ARG(d_Ut && d_piv && d_Theta, "null pointer");
    OMB_CHECK_ARG(r > 0 && s > 0 && s * r < (1ll << 30), "bad size");
    int g = (int)ceil_div(s * r, 256);
    if (g > 1024) g = 1024;
    gather_rows_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(d_Ut, (int)r, d_piv, (int)s, d_Theta, d_cnt, d_cnt_s);
    return check_launch("gather_rows_kernel");
}

extern "C" int omb_modes_to_rows(const double* d_Ut, int64_t n, int64_t r, double* d_Ur, void* stream)
{
    OMB_CHECK_ARG(d_Ut && d_Ur, "null pointer");
    OMB_CHECK_ARG(n > 0 && r > 0, "bad size");
    int64_t g = ceil_div(n, 32) * ceil_div(r, 32);
    if (g > (int64_t)sm_count() * 32) g = (int64_t)sm_count() * 32;
    modes_to_rows_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(d_Ut, n, (int)r, d_Ur);
    return check_launch("modes_to_rows_kernel");
}

extern "C" int omb_rows_to_modes(const double* d_Ur, int64_t n, int64_t r, double* d_Ut, double* d_vn, void* stream)
{
    OMB_CHECK_ARG(d_Ur && d_Ut, "null pointer");
    OMB_CHECK_ARG(n > 0 && r > 0, "bad size");
    int64_t g = ceil_div(n, 32) * ceil_div(r, 32);
    if (g > (int64_t)sm_count() * 32) g = (int64_t)sm_count() * 32;
    rows_to_modes_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(d_Ur, n, (int)r, d_Ut);
    int rc = check_launch("rows_to_modes_kernel");
    if (rc || !d_vn) return rc;
    int64_t g2 = ceil_div(n, 256);
    if (g2 > (int64_t)sm_count() * 8) g2 = (int64_t)sm_count() * 8;
    row_norms_kernel<<<(unsigned)g2, 256, 0, (cudaStream_t)stream>>>(d_Ut, n, (int)r, d_vn);
    return check_launch("row_norms_kernel");
}

extern "C" int omb_ols_predict(const double* d_Y, const double* d_cnt_s, const double* d_scl_s,
                               const double* d_PinvT, int64_t N, int64_t s, int64_t r, double* d_A, void* stream)
{
    OMB_CHECK_ARG(d_Y && d_PinvT && d_A, "null pointer");
    OMB_CHECK_ARG(N > 0 && s > 0 && r > 0 && s < (1 << 24), "bad size");
    int64_t tiles = ceil_div(N, OT) * ceil_div(r, OT);
    if (tiles > (int64_t)sm_count() * 8) tiles = (int64_t)sm_count() * 8;
    ols_gemm_kernel<0><<<(unsigned)tiles, O_THREADS, 0, (cudaStream_t)stream>>>(d_Y, d_PinvT, N, r, (int)s, s, d_cnt_s,
                                                                                 d_scl_s, 1, 0, d_A);
    return check_launch("ols_gemm_kernel<predict>");
}

extern "C" int omb_reconstruct(const double* d_Ut, int64_t n, int64_t r, const double* d_A, int64_t N,
                               const double* d_cnt, const double* d_scl, int64_t n_c, int64_t row0, int64_t nrows,
                               double* d_out, void* stream)
{
    OMB_CHECK_ARG(d_Ut && d_A && d_out, "null pointer");
    OMB_CHECK_ARG(N > 0 && r > 0 && nrows > 0 && row0 >= 0 && n_c > 0 && r < (1 << 24), "bad size");
    OMB_CHECK_ARG(row0 + nrows <= n, "row range exceeds n");
    if ((r & 1) == 0 && (N & 1) == 0 && row0 % OMB_TB == 0 && r >= 16 && N >= 16 &&
        ((reinterpret_cast<uintptr_t>(d_Ut) | reinterpret_cast<uintptr_t>(d_A) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0) {
        OMB_CUDA(cudaFuncSetAttribute(reconstruct_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)reconstruct_big_smem()));
        int64_t g = ceil_div(nrows, (int64_t)OMB_TB) * ceil_div(N, (int64_t)OMB_TB);
        if (g > sm_count()) g = sm_count();
        reconstruct_big_kernel<<<(unsigned)g, RB_THREADS, reconstruct_big_smem(), (cudaStream_t)stream>>>(
            d_Ut, (int)r, d_A, N, d_cnt, d_scl, n_c, row0, nrows, d_out);
        return check_launch("reconstruct_big_kernel");
    }
    int64_t tiles = ceil_div(nrows, OT) * ceil_div(N, OT);
    if (tiles > (int64_t)sm_count() * 8) tiles = (int64_t)sm_count() * 8;
    ols_gemm_kernel<1><<<(unsigned)tiles, O_THREADS, 0, (cudaStream_t)stream>>>(d_Ut, d_A, nrows, N, (int)r, 0, d_cnt,
                                                                                 d_scl, n_c, row0, d_out);
    return check_launch("ols_gemm_kernel<reconstruct>");
}
