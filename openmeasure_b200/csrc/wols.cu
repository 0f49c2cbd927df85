// wols.cu -- weighted OLS predict, batched.
//
// Replaces the per-vector branch of SPR.predict for measurements with non-zero uncertainties (reference
// sparse_sensing.py:871-878):   W = diag(1 / y0[:,1]);   ar = pinv(W Theta) (W y0[:,0]);
//                               ar_sigma = | pinv(W Theta) y0[:,1] |
// The reference takes an SVD-based pseudo-inverse of the s x r matrix W Theta for every vector in a Python
// loop.  Here one CTA per vector factorises W Theta by Householder QR in shared memory (column-major, both
// right-hand sides carried as two extra columns) and back-substitutes: for a full-column-rank W Theta the
// least-squares solution IS pinv(W Theta) b.  A vector whose R has a (numerically) zero diagonal entry is
// flagged instead of solved -- the caller sends those through the pseudo-inverse route.
// Deterministic: fixed thread -> column ownership, fixed-order reductions.
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

constexpr int WO_THREADS = 256;

__global__ void __launch_bounds__(WO_THREADS)
wols_kernel(const double* __restrict__ Theta, int s, int r, const double* __restrict__ y0v, const double* __restrict__ y0s,
            int64_t N, double rank_tol, double* __restrict__ Ar, double* __restrict__ Asig, int* __restrict__ flag)
{
    extern __shared__ double sm[];
    const int lds = s | 1;                           // odd leading dimension: columns start in different banks
    double* A = sm;                                  // [r + 2][lds] column-major: W Theta | W y0v | y0s
    double* red = sm + (size_t)(r + 2) * lds;        // [WO_THREADS / 32 + 4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc = r + 2;

    for (int64_t b = blockIdx.x; b < N; b += gridDim.x) {
        const double* yv = y0v + b * s;
        const double* ys = y0s + b * s;
        __syncthreads();
        for (int e = threadIdx.x; e < s * r; e += WO_THREADS) {
            const int i = e / r, j = e - i * r;
            A[j * lds + i] = Theta[e] / ys[i];                       // (1 / y0s_i) Theta_ij, the reference's W @ Theta
        }
        for (int i = threadIdx.x; i < s; i += WO_THREADS) {
            A[r * lds + i] = yv[i] / ys[i];                          // W y0[:,0]
            A[(r + 1) * lds + i] = ys[i];                            // y0[:,1]
        }
        __syncthreads();
        double dmin = 1e300, dmax = 0.0;
        for (int k = 0; k < r; ++k) {
            // ---- Householder vector of column k (rows k..s-1): warp 0 forms it, everybody waits
            if (warp == 0) {
                double* x = A + k * lds;
                double ss = 0.0;
                for (int i = k + 1 + lane; i < s; i += 32) ss = fma(x[i], x[i], ss);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
                const double alpha = x[k];
                const double nrm = sqrt(fma(alpha, alpha, ss));
                const double beta = alpha >= 0.0 ? -nrm : nrm;
                // v = [1, x[k+1:] / (alpha - beta)], tau = (beta - alpha) / beta     (dlarfg)
                const double denom = alpha - beta;
                const double tau = (nrm == 0.0) ? 0.0 : (beta - alpha) / beta;
                if (nrm != 0.0)
                    for (int i = k + 1 + lane; i < s; i += 32) x[i] = x[i] / denom;
                __syncwarp();
                if (lane == 0) { x[k] = beta; red[0] = tau; }
            }
            __syncthreads();
            const double tau = red[0];
            const double* v = A + k * lds;
            // ---- apply H = I - tau v v^T to the columns right of k (thread = column)
            for (int j = k + 1 + threadIdx.x; j < nc; j += WO_THREADS) {
                double* c = A + j * lds;
                double w = c[k];
                for (int i = k + 1; i < s; ++i) w = fma(v[i], c[i], w);
                w *= tau;
                c[k] -= w;
                for (int i = k + 1; i < s; ++i) c[i] = fma(-w, v[i], c[i]);
            }
            const double d = fabs(A[k * lds + k]);
            dmin = d < dmin ? d : dmin;
            dmax = d > dmax ? d : dmax;
            __syncthreads();
        }
        const bool singular = !(dmin > rank_tol * dmax);
        if (threadIdx.x == 0) flag[b] = singular ? 1 : 0;
        // ---- back substitution R x = (Q^T b)[0:r], both right-hand sides (threads 0, 1 own one each; column sweep)
        if (!singular) {
            for (int k = r - 1; k >= 0; --k) {
                if (threadIdx.x < 2) {
                    double* z = A + (r + threadIdx.x) * lds;
                    z[k] = z[k] / A[k * lds + k];
                }
                __syncthreads();
                const double x1 = A[r * lds + k], x2 = A[(r + 1) * lds + k];
                for (int i = threadIdx.x; i < k; i += WO_THREADS) {
                    const double rik = A[k * lds + i];
                    A[r * lds + i] = fma(-rik, x1, A[r * lds + i]);
                    A[(r + 1) * lds + i] = fma(-rik, x2, A[(r + 1) * lds + i]);
                }
                __syncthreads();
            }
            for (int q = threadIdx.x; q < r; q += WO_THREADS) {
                Ar[b * r + q] = A[r * lds + q];
                Asig[b * r + q] = fabs(A[(r + 1) * lds + q]);
            }
        }
    }
}

}  // namespace omb

using namespace omb;

extern "C" int64_t omb_wols_smem_bytes(int64_t s, int64_t r)
{
    return (int64_t)sizeof(double) * ((r + 2) * (s | 1) + WO_THREADS / 32 + 4);
}

extern "C" int omb_wols_predict(const double* d_Theta, int64_t s, int64_t r, const double* d_y0v, const double* d_y0s,
                                int64_t N, double rank_tol, double* d_Ar, double* d_Asig, int* d_flag, void* stream)
{
    OMB_CHECK_ARG(d_Theta && d_y0v && d_y0s && d_Ar && d_Asig && d_flag, "null pointer");
    OMB_CHECK_ARG(N > 0 && r > 0 && s >= r, "need N > 0 and s >= r > 0");
    const int64_t smem = omb_wols_smem_bytes(s, r);
    OMB_CHECK_ARG(smem <= 227 * 1024, "s x r too large for the shared-memory factorisation");
    OMB_CUDA(cudaFuncSetAttribute(wols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t g = N;
    const int64_t cap = (int64_t)sm_count() * (smem <= 100 * 1024 ? 2 : 1);
    if (g > cap) g = cap;
    wols_kernel<<<(unsigned)g, WO_THREADS, (size_t)smem, (cudaStream_t)stream>>>(d_Theta, (int)s, (int)r, d_y0v, d_y0s, N,
                                                                                   rank_tol, d_Ar, d_Asig, d_flag);
    return check_launch("wols_kernel");
}
