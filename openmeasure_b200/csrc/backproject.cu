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
                const double sv = s_scl[i];
#pragma unroll
                for (int b = 0; b < QB; ++b) {
                    const int q = b * 8 + 2 * fr;
                    sC[q * BP_LDC + i] = c[a][b][0] / sv;
                    sC[(q + 1) * BP_LDC + i] = c[a][b][1] / sv;
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
    if (r <= 64) return launch_bp<8>(d_X, n, n_c, (int)m, d_cnt, d_scl, d_W, (int)r, d_Ut, d_vn, st);
    return launch_bp<16>(d_X, n, n_c, (int)m, d_cnt, d_scl, d_W, (int)r, d_Ut, d_vn, st);
}
