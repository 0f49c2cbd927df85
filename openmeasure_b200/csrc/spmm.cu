// spmm.cu -- general (non-one-hot) measurement / sampling matrices in CSR form.
//
// Replaces Theta = C.dot(Ur) (reference sparse_sensing.py:797), C.dot(X_cnt) (:573) and the sampled
// scale / centre of unscale_data(sampling=) (:233) and reconstruct(sampling=) (:365-368) for the
// line-of-sight matrices of utils.camera.project (scipy CSR, utils.py:318-469).  The matrix is never
// densified: a 1000 x 16.2M line-of-sight matrix is 130 GB as a dense array, its CSR form a few MB.
//   Theta[i][q] = sum_k data[k] * U[idx[k]][q]     k in [indptr[i], indptr[i+1]), CSR order
//   cs[i]       = sum_k data[k] * cnt[idx[k]]       ss[i] = sum_k data[k] * scl[idx[k] / n_c]
// Row-sharded runs hand every rank the columns of C that fall on its rows (local indices); the s x (r+2)
// partial results are combined in rank order like every other small object of the path.
// Deterministic: a row's non-zeros are cut into fixed segments of CSR_SEG, each summed sequentially in CSR
// order (multiply, then add -- scipy's csr_matvecs does not fuse), segments added in order.
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

constexpr int CSR_SEG = 4096;      // non-zeros per segment
constexpr int CSR_CHUNK = 256;     // staged per iteration
constexpr int CSR_THREADS = 128;

__global__ void __launch_bounds__(CSR_THREADS)
csr_segment_kernel(const int64_t* __restrict__ indptr, const int64_t* __restrict__ indices, const double* __restrict__ data,
                   const double* __restrict__ Ut, int r, const double* __restrict__ cnt, const double* __restrict__ scl,
                   int64_t n_c, int nseg, double* __restrict__ part)
{
    __shared__ int64_t s_idx[CSR_CHUNK];
    __shared__ double s_val[CSR_CHUNK];
    const int64_t i = blockIdx.x;
    const int seg = blockIdx.y;
    const int64_t k_lo = indptr[i] + (int64_t)seg * CSR_SEG;
    int64_t k_hi = k_lo + CSR_SEG;
    if (k_hi > indptr[i + 1]) k_hi = indptr[i + 1];
    const int width = r + 2;
    double* out = part + ((int64_t)i * nseg + seg) * width;
    // column q < r: basis mode q; column r: centring value; column r + 1: scale
    for (int q0 = 0; q0 < width; q0 += CSR_THREADS) {
        const int q = q0 + threadIdx.x;
        double acc = 0.0;
        for (int64_t k0 = k_lo; k0 < k_hi; k0 += CSR_CHUNK) {
            __syncthreads();
            for (int e = threadIdx.x; e < CSR_CHUNK; e += CSR_THREADS) {
                const int64_t k = k0 + e;
                s_idx[e] = k < k_hi ? indices[k] : 0;
                s_val[e] = k < k_hi ? data[k] : 0.0;
            }
            __syncthreads();
            const int cnt_k = (int)((k_hi - k0) < CSR_CHUNK ? (k_hi - k0) : CSR_CHUNK);
            if (q < width) {
                for (int e = 0; e < cnt_k; ++e) {
                    const int64_t j = s_idx[e];
                    double u;
                    if (q < r) u = Ut ? Ut[basis_index(q, j, r)] : 0.0;
                    else if (q == r) u = cnt ? cnt[j] : 0.0;
                    else u = scl ? scl[j / n_c] : 1.0;
                    const double t = s_val[e] * u;
                    acc = acc + t;
                }
            }
        }
        if (q < width) out[q] = acc;
    }
}

__global__ void __launch_bounds__(128)
csr_combine_kernel(const double* __restrict__ part, const int64_t* __restrict__ indptr, int64_t s, int r, int nseg,
                   double* __restrict__ Theta, double* __restrict__ cs, double* __restrict__ ss)
{
    const int width = r + 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < s * width; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / width;
        const int q = (int)(e - i * width);
        const int64_t nnz = indptr[i + 1] - indptr[i];
        const int used = (int)ceil_div(nnz, (int64_t)CSR_SEG);
        double acc = 0.0;
        for (int g = 0; g < used && g < nseg; ++g) acc = acc + part[(i * nseg + g) * width + q];
        if (q < r) { if (Theta) Theta[i * r + q] = acc; }
        else if (q == r) { if (cs) cs[i] = acc; }
        else if (ss) ss[i] = acc;
    }
}

}  // namespace omb

using namespace omb;

extern "C" int64_t omb_csr_ws_bytes(int64_t s, int64_t max_row_nnz, int64_t r)
{
    if (s <= 0 || r < 0) return 0;
    int64_t nseg = ceil_div(max_row_nnz > 0 ? max_row_nnz : 1, (int64_t)CSR_SEG);
    return (int64_t)sizeof(double) * s * nseg * (r + 2);
}

extern "C" int omb_csr_times_basis(const int64_t* d_indptr, const int64_t* d_indices, const double* d_data, int64_t s,
                                   int64_t max_row_nnz, const double* d_Ut, int64_t n, int64_t r, const double* d_cnt,
                                   const double* d_scl, int64_t n_c, double* d_Theta, double* d_cnt_s, double* d_scl_s,
                                   void* d_ws, void* stream)
{
    OMB_CHECK_ARG(d_indptr && d_ws, "null pointer");
    OMB_CHECK_ARG(s > 0 && s < (1 << 30) && r >= 0 && r < (1 << 20) && n > 0 && n_c > 0 && max_row_nnz >= 0, "bad size");
    OMB_CHECK_ARG(max_row_nnz == 0 || (d_indices && d_data), "null pointer");
    OMB_CHECK_ARG(r == 0 || !d_Theta || d_Ut, "a basis is needed for Theta");
    const int64_t nseg = ceil_div(max_row_nnz > 0 ? max_row_nnz : 1, (int64_t)CSR_SEG);
    OMB_CHECK_ARG(nseg <= 65535, "row too long");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)s, (unsigned)nseg);
    csr_segment_kernel<<<grid, CSR_THREADS, 0, st>>>(d_indptr, d_indices, d_data, d_Theta ? d_Ut : nullptr, (int)r, d_cnt, d_scl,
                                                     n_c, (int)nseg, (double*)d_ws);
    int rc = check_launch("csr_segment_kernel");
    if (rc) return rc;
    int64_t g = ceil_div(s * (r + 2), 128);
    if (g > 1024) g = 1024;
    csr_combine_kernel<<<(unsigned)g, 128, 0, st>>>((const double*)d_ws, d_indptr, s, (int)r, (int)nseg, d_Theta, d_cnt_s, d_scl_s);
    return check_launch("csr_combine_kernel");
}
