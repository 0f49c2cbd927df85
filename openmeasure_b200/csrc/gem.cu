// gem.cu -- greedy entropy-maximisation sensor placement.
//
// Replaces the candidate loop of SPR.gem (reference sparse_sensing.py:586-698): for every
// candidate row y (a row of the scaled basis, r "observations") and the chosen rows a,
//     sigma2_cond(y) = var(y) - Sigma_ya Sigma_aa^-1 Sigma_ay        (np.cov, ddof = 1; :670-678)
// and the next sensor is its argmax (first index on ties, np.argmax).  The reference evaluates it
// with one np.cov per candidate in a Python loop; here it is ONE streaming pass per step over the
// tiled mode-major basis (thread = candidate, coalesced along the candidates, same traffic shape
// as a pivoted-QR pass: 8 n r bytes): the k cross-covariances are dot products with the centred
// chosen rows (which sum to zero, so the candidate needs no centring), the k x k inverse
// (including the reference's random diagonal jitter, drawn on the host with the same numpy calls)
// arrives as an argument.  The d_min exclusion (:646-649, :685-688) is a byte mask.
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

constexpr int GEM_THREADS = OMB_TB;      // one CTA iteration sweeps one basis tile
constexpr int GEM_KMAX = 64;             // most sensors a placement can ask for
constexpr int GEM_NCAND = 4096;          // per-CTA argmax records

struct GemCand { double val; int64_t idx; };

__device__ __forceinline__ bool gem_better(double v, int64_t i, double bv, int64_t bi)
{
    return v > bv || (v == bv && i < bi);
}

// unscaled sample variance (ddof = 1) of every row over its r modes, two passes over the row
__global__ void __launch_bounds__(GEM_THREADS)
gem_variance_kernel(const double* __restrict__ Ut, int64_t n, int r, double* __restrict__ var)
{
    const int64_t ntiles = basis_tiles(n);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t j = tile * OMB_TB + threadIdx.x;
        if (j >= n) continue;
        const double* col = Ut + tile * ((int64_t)r * OMB_TB) + threadIdx.x;
        double s = 0.0;
        for (int q = 0; q < r; ++q) s += col[(int64_t)q * OMB_TB];
        const double mu = s / (double)r;
        double ss = 0.0;
        for (int q = 0; q < r; ++q) { const double d = col[(int64_t)q * OMB_TB] - mu; ss = fma(d, d, ss); }
        var[j] = ss / (double)(r - 1);
    }
}

// one greedy step: conditional variance of every live candidate given k chosen rows, per-CTA argmax
template <int KB>      // k <= 8 * KB
__global__ void __launch_bounds__(GEM_THREADS)
gem_step_kernel(const double* __restrict__ Ut, int64_t n, int r, double coef, int k, const double* __restrict__ Z,
                const double* __restrict__ B, const double* __restrict__ var, const unsigned char* __restrict__ alive,
                GemCand* __restrict__ cand)
{
    extern __shared__ double sm[];
    double* sZ = sm;                          // [r][8 KB]  centred, scaled chosen rows, mode-major
    double* sB = sm + (size_t)r * 8 * KB;     // [8 KB][8 KB]
    constexpr int K8 = 8 * KB;
    for (int e = threadIdx.x; e < r * K8; e += GEM_THREADS) {
        const int q = e / K8, a = e - q * K8;
        sZ[e] = a < k ? Z[(int64_t)a * r + q] : 0.0;
    }
    for (int e = threadIdx.x; e < K8 * K8; e += GEM_THREADS) {
        const int a = e / K8, b = e - a * K8;
        sB[e] = (a < k && b < k) ? B[a * k + b] : 0.0;
    }
    __syncthreads();

    const double c2 = coef * coef, inv = 1.0 / (double)(r - 1);
    GemCand best;
    best.val = -__longlong_as_double(0x7FF0000000000000LL);
    best.idx = INT64_MAX;
    const int64_t ntiles = basis_tiles(n);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t j = tile * OMB_TB + threadIdx.x;
        if (j >= n || !alive[j]) continue;
        double temp = c2 * var[j];
        if (k > 0) {
            const double* col = Ut + tile * ((int64_t)r * OMB_TB) + threadIdx.x;
            double c[K8];
#pragma unroll
            for (int a = 0; a < K8; ++a) c[a] = 0.0;
            for (int q = 0; q < r; ++q) {
                const double u = coef * ldg_stream(col + (int64_t)q * OMB_TB);
                const double* z = sZ + q * K8;
#pragma unroll
                for (int a = 0; a < K8; ++a) c[a] = fma(z[a], u, c[a]);
            }
            double quad = 0.0;
#pragma unroll
            for (int a = 0; a < K8; ++a) {
                double w = 0.0;
#pragma unroll
                for (int b = 0; b < K8; ++b) w = fma(sB[a * K8 + b], c[b], w);
                quad = fma(c[a], w, quad);
            }
            temp = temp - quad * inv * inv;            // c holds (r - 1) * Sigma_ya
        }
        if (gem_better(temp, j, best.val, best.idx)) { best.val = temp; best.idx = j; }
    }
    __shared__ GemCand s_c[GEM_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v = __shfl_xor_sync(0xFFFFFFFFu, best.val, o);
        const int64_t i = __shfl_xor_sync(0xFFFFFFFFu, best.idx, o);
        if (gem_better(v, i, best.val, best.idx)) { best.val = v; best.idx = i; }
    }
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < GEM_THREADS / 32; ++w)
            if (gem_better(s_c[w].val, s_c[w].idx, best.val, best.idx)) best = s_c[w];
        cand[blockIdx.x] = best;
    }
}

// winner of the step, its (unscaled) basis row, and -- when positions are given -- the d_min exclusion
__global__ void __launch_bounds__(256)
gem_pick_kernel(const GemCand* __restrict__ cand, int ncand, const double* __restrict__ Ut, int r,
                int64_t* __restrict__ out_idx, double* __restrict__ out_val, double* __restrict__ out_row)
{
    __shared__ GemCand s_c[256];
    GemCand best;
    best.val = -__longlong_as_double(0x7FF0000000000000LL);
    best.idx = INT64_MAX;
    for (int e = threadIdx.x; e < ncand; e += 256)
        if (gem_better(cand[e].val, cand[e].idx, best.val, best.idx)) best = cand[e];
    s_c[threadIdx.x] = best;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if (threadIdx.x < h && gem_better(s_c[threadIdx.x + h].val, s_c[threadIdx.x + h].idx, s_c[threadIdx.x].val, s_c[threadIdx.x].idx))
            s_c[threadIdx.x] = s_c[threadIdx.x + h];
        __syncthreads();
    }
    best = s_c[0];
    if (threadIdx.x == 0) { *out_idx = best.idx == INT64_MAX ? -1 : best.idx; *out_val = best.val; }
    if (best.idx != INT64_MAX)
        for (int q = threadIdx.x; q < r; q += 256) out_row[q] = Ut[basis_index(q, best.idx, r)];
}

__global__ void __launch_bounds__(256)
gem_exclude_kernel(const double* __restrict__ xyz, int64_t n_c, int64_t n, const int64_t* __restrict__ sensor,
                   double d_min, unsigned char* __restrict__ alive)
{
    const int64_t p = *sensor;
    if (p < 0) return;
    const int64_t pc = p % n_c;
    const double px = xyz[pc * 3 + 0], py = xyz[pc * 3 + 1], pz = xyz[pc * 3 + 2];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = j % n_c;
        const double dx = px - xyz[c * 3 + 0], dy = py - xyz[c * 3 + 1], dz = pz - xyz[c * 3 + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);          // np.linalg.norm(p - xyz, axis=1)
        if (!(d >= d_min)) alive[j] = 0;
    }
}

// the same exclusion around a point given by its coordinates (multi-rank: the chosen cell may live on a peer)
__global__ void __launch_bounds__(256)
gem_exclude_point_kernel(const double* __restrict__ xyz, int64_t n_c, int64_t n, double px, double py, double pz,
                         double d_min, unsigned char* __restrict__ alive)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = j % n_c;
        const double dx = px - xyz[c * 3 + 0], dy = py - xyz[c * 3 + 1], dz = pz - xyz[c * 3 + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        if (!(d >= d_min)) alive[j] = 0;
    }
}

template <int KB>
static int launch_gem_step(const double* Ut, int64_t n, int r, double coef, int k, const double* Z, const double* B,
                           const double* var, const unsigned char* alive, GemCand* cand, int grid, cudaStream_t st)
{
    const size_t smem = sizeof(double) * ((size_t)r * 8 * KB + 64 * KB * KB);
    OMB_CUDA(cudaFuncSetAttribute(gem_step_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gem_step_kernel<KB><<<grid, GEM_THREADS, smem, st>>>(Ut, n, r, coef, k, Z, B, var, alive, cand);
    return check_launch("gem_step_kernel");
}

}  // namespace omb

using namespace omb;

extern "C" int64_t omb_gem_ws_bytes(void) { return (int64_t)sizeof(GemCand) * GEM_NCAND; }
extern "C" int omb_gem_max_sensors(void) { return GEM_KMAX; }

extern "C" int omb_gem_variance(const double* d_Ut, int64_t n, int64_t r, double* d_var, void* stream)
{
    OMB_CHECK_ARG(d_Ut && d_var, "null pointer");
    OMB_CHECK_ARG(n > 0 && r > 1, "need n > 0 and r > 1");
    int64_t g = basis_tiles(n);
    if (g > (int64_t)sm_count() * 8) g = (int64_t)sm_count() * 8;
    gem_variance_kernel<<<(unsigned)g, GEM_THREADS, 0, (cudaStream_t)stream>>>(d_Ut, n, (int)r, d_var);
    return check_launch("gem_variance_kernel");
}

extern "C" int omb_gem_step(const double* d_Ut, int64_t n, int64_t r, double coef, int64_t k, const double* d_Z,
                            const double* d_B, const double* d_var, const unsigned char* d_alive, void* d_ws,
                            int64_t* d_idx, double* d_val, double* d_row, void* stream)
{
    OMB_CHECK_ARG(d_Ut && d_var && d_alive && d_ws && d_idx && d_val && d_row, "null pointer");
    OMB_CHECK_ARG(n > 0 && r > 1 && r <= 1024, "need n > 0 and 1 < r <= 1024");
    OMB_CHECK_ARG(k >= 0 && k <= GEM_KMAX, "k must be in [0, 64]");
    OMB_CHECK_ARG(k == 0 || (d_Z && d_B), "chosen rows and inverse covariance required for k > 0");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t g = basis_tiles(n);
    if (g > (int64_t)sm_count() * 4) g = (int64_t)sm_count() * 4;
    if (g > GEM_NCAND) g = GEM_NCAND;
    GemCand* cand = (GemCand*)d_ws;
    const int kb = k <= 8 ? 1 : (k <= 16 ? 2 : (k <= 32 ? 4 : 8));
    int rc;
    switch (kb) {
        case 1: rc = launch_gem_step<1>(d_Ut, n, (int)r, coef, (int)k, d_Z, d_B, d_var, d_alive, cand, (int)g, st); break;
        case 2: rc = launch_gem_step<2>(d_Ut, n, (int)r, coef, (int)k, d_Z, d_B, d_var, d_alive, cand, (int)g, st); break;
        case 4: rc = launch_gem_step<4>(d_Ut, n, (int)r, coef, (int)k, d_Z, d_B, d_var, d_alive, cand, (int)g, st); break;
        default: rc = launch_gem_step<8>(d_Ut, n, (int)r, coef, (int)k, d_Z, d_B, d_var, d_alive, cand, (int)g, st); break;
    }
    if (rc) return rc;
    gem_pick_kernel<<<1, 256, 0, st>>>(cand, (int)g, d_Ut, (int)r, d_idx, d_val, d_row);
    return check_launch("gem_pick_kernel");
}

extern "C" int omb_gem_exclude(const double* d_xyz, int64_t n_c, int64_t n, const int64_t* d_sensor, double d_min,
                               unsigned char* d_alive, void* stream)
{
    OMB_CHECK_ARG(d_xyz && d_sensor && d_alive, "null pointer");
    OMB_CHECK_ARG(n_c > 0 && n > 0, "non-positive size");
    int64_t g = ceil_div(n, 256);
    if (g > (int64_t)sm_count() * 8) g = (int64_t)sm_count() * 8;
    gem_exclude_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(d_xyz, n_c, n, d_sensor, d_min, d_alive);
    return check_launch("gem_exclude_kernel");
}

extern "C" int omb_gem_exclude_point(const double* d_xyz, int64_t n_c, int64_t n, double px, double py, double pz,
                                     double d_min, unsigned char* d_alive, void* stream)
{
    OMB_CHECK_ARG(d_xyz && d_alive, "null pointer");
    OMB_CHECK_ARG(n_c > 0 && n > 0, "non-positive size");
    int64_t g = ceil_div(n, 256);
    if (g > (int64_t)sm_count() * 8) g = (int64_t)sm_count() * 8;
    gem_exclude_point_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(d_xyz, n_c, n, px, py, pz, d_min, d_alive);
    return check_launch("gem_exclude_point_kernel");
}
