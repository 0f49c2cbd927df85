// eigh.cu -- S3: the m x m symmetric eigensolve of the POD Gram matrix for few snapshots (m <= 64).
//
// Replaces the bidiagonal divide-and-conquer on R inside LAPACK dgesdd (np.linalg.svd, reference
// sparse_sensing.py:272).  One CTA, matrix and eigenvectors in shared memory, cyclic two-sided
// Jacobi with the round-robin parallel ordering (M/2 disjoint rotations per step, M-1 steps per
// sweep).  For an m = 41 Gram this takes ~0.2 ms, where a library syevd spends ~0.9 ms mostly in
// host synchronisation; Jacobi also resolves the small eigenvalues of a positive semi-definite
// matrix to high relative accuracy.  Output: eigenvalues in DESCENDING order, V[i][k] = component
// i of eigenvector k.
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

constexpr int EJ_MAX = 64;
constexpr int EJ_LD = EJ_MAX + 1;
constexpr int EJ_THREADS = 512;
constexpr int EJ_MAX_SWEEPS = 30;

__global__ void __launch_bounds__(EJ_THREADS)
eigh_jacobi_kernel(const double* __restrict__ G, int m, double* __restrict__ w_out, double* __restrict__ V_out,
                   int* __restrict__ info)
{
    extern __shared__ double sm[];
    double* A = sm;                         // [EJ_MAX][EJ_LD]
    double* V = sm + EJ_MAX * EJ_LD;        // [EJ_MAX][EJ_LD]
    __shared__ double s_c[EJ_MAX / 2], s_s[EJ_MAX / 2];
    __shared__ int s_p[EJ_MAX / 2], s_q[EJ_MAX / 2];
    __shared__ int s_rot;                   // rotations applied in the current sweep
    __shared__ int s_order[EJ_MAX];
    __shared__ unsigned short s_sched[(EJ_MAX - 1) * (EJ_MAX / 2)];

    const int M = (m + 1) & ~1;             // even number of players; index m (if any) is a dummy
    const int npair = M / 2;
    for (int e = threadIdx.x; e < M * M; e += EJ_THREADS) {
        const int i = e / M, j = e - i * M;
        // symmetrise from the upper triangle so that the iteration starts exactly symmetric; the
        // dummy player of an odd m is a zero row/column that only ever meets identity rotations
        double a = 0.0;
        if (i < m && j < m) a = (i <= j) ? G[i * m + j] : G[j * m + i];
        A[i * EJ_LD + j] = a;
        V[i * EJ_LD + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    // absolute floor for rotations: entries below eps^1.25 * max|a_ii| cannot move any eigenvalue
    // by more than that (the Gram of row-centred data is exactly rank deficient, and its null
    // direction would otherwise keep the relative criterion busy with rounding noise forever)
    __shared__ double s_floor;
    if (threadIdx.x == 0) {
        double dmax = 0.0;
        for (int i = 0; i < m; ++i) dmax = fmax(dmax, fabs(A[i * EJ_LD + i]));
        s_floor = dmax * 1.0e-20;
    }
    __syncthreads();
    const double floor_abs = s_floor, big_abs = s_floor * 1.0e7;      // 1e-13 * max|a_ii|

    // round-robin schedule: pair i of step k, stored once (no integer division in the sweeps)
    for (int e = threadIdx.x; e < (M - 1) * npair; e += EJ_THREADS) {
        const int step = e / npair, i = e - step * npair;
        int p, q;
        if (i == 0) { p = M - 1; q = step; }
        else { p = (step + i) % (M - 1); q = (step - i + (M - 1)) % (M - 1); }
        if (p > q) { const int tswap = p; p = q; q = tswap; }
        s_sched[e] = (unsigned short)((p << 8) | q);
    }
    __syncthreads();

    int sweep = 0;
    for (; sweep < EJ_MAX_SWEEPS; ++sweep) {
        if (threadIdx.x == 0) s_rot = 0;
        __syncthreads();
        for (int step = 0; step < M - 1; ++step) {
            // 1. the M/2 disjoint pairs of this step and their rotations
            if (threadIdx.x < npair) {
                const int i = threadIdx.x;
                const unsigned short pq = s_sched[step * npair + i];
                const int p = pq >> 8, q = pq & 255;
                double c = 1.0, s = 0.0;
                if (q < m) {
                    const double apq = A[p * EJ_LD + q];
                    const double app = A[p * EJ_LD + p], aqq = A[q * EJ_LD + q];
                    const double rel2 = apq * apq, den = fabs(app * aqq);
                    // rotate unless |apq| <= eps * sqrt(app * aqq) (compared squared) or below the floor
                    if (rel2 > 1.2325951644078309e-32 * den && fabs(apq) > floor_abs) {
                        // t = tan(theta) of the smaller root: b / (d + sign(d) * hypot(d, b))
                        const double d = aqq - app, b = 2.0 * apq;
                        const double h = sqrt(fma(d, d, b * b));
                        const double t = b / (d + copysign(h, d));
                        c = rsqrt(fma(t, t, 1.0));
                        s = t * c;
                        // "big" rotation: convergence is quadratic, so a sweep without any of these
                        // is the last one that can change an eigenvalue at the 1e-16 * lambda_max level
                        if (rel2 > 1.0e-18 * den && fabs(apq) > big_abs) atomicAdd(&s_rot, 1);
                    }
                }
                s_p[i] = p; s_q[i] = q; s_c[i] = c; s_s[i] = s;
            }
            __syncthreads();
            // 2. A <- J^T A J by 2 x 2 blocks (pair i rows x pair j columns), V <- V J by column pairs
            const int nblk = npair * npair;
            for (int e = threadIdx.x; e < nblk + npair * m; e += EJ_THREADS) {
                if (e < nblk) {
                    const int i = e / npair, j = e - i * npair;
                    const double si = s_s[i], sj = s_s[j];
                    if (si == 0.0 && sj == 0.0) continue;
                    const int pi = s_p[i], qi = s_q[i], pj = s_p[j], qj = s_q[j];
                    const double ci = s_c[i], cj = s_c[j];
                    const double a00 = A[pi * EJ_LD + pj], a01 = A[pi * EJ_LD + qj];
                    const double a10 = A[qi * EJ_LD + pj], a11 = A[qi * EJ_LD + qj];
                    // columns (J_j), then rows (J_i^T)
                    const double b00 = cj * a00 - sj * a01, b01 = sj * a00 + cj * a01;
                    const double b10 = cj * a10 - sj * a11, b11 = sj * a10 + cj * a11;
                    A[pi * EJ_LD + pj] = ci * b00 - si * b10;
                    A[pi * EJ_LD + qj] = ci * b01 - si * b11;
                    A[qi * EJ_LD + pj] = si * b00 + ci * b10;
                    A[qi * EJ_LD + qj] = si * b01 + ci * b11;
                } else {
                    const int e2 = e - nblk;
                    const int i = e2 / m, row = e2 - i * m;
                    const double s = s_s[i];
                    if (s != 0.0) {
                        const int p = s_p[i], q = s_q[i];
                        const double c = s_c[i];
                        const double vp = V[row * EJ_LD + p], vq = V[row * EJ_LD + q];
                        V[row * EJ_LD + p] = c * vp - s * vq;
                        V[row * EJ_LD + q] = s * vp + c * vq;
                    }
                }
            }
            __syncthreads();
        }
        if (s_rot == 0) break;
        __syncthreads();
    }

    // descending order by rank counting (ties broken by index)
    if (threadIdx.x < m) {
        const int i = threadIdx.x;
        const double wi = A[i * EJ_LD + i];
        int rank = 0;
        for (int j = 0; j < m; ++j) {
            const double wj = A[j * EJ_LD + j];
            rank += (wj > wi) || (wj == wi && j < i);
        }
        s_order[rank] = i;
    }
    __syncthreads();
    if (threadIdx.x < m) w_out[threadIdx.x] = A[s_order[threadIdx.x] * EJ_LD + s_order[threadIdx.x]];
    // deterministic sign: the largest-magnitude component of every eigenvector is positive
    __shared__ double s_sign[EJ_MAX];
    if (threadIdx.x < m) {
        const int col = s_order[threadIdx.x];
        double best = -1.0, sg = 1.0;
        for (int i = 0; i < m; ++i) {
            const double v = V[i * EJ_LD + col];
            if (fabs(v) > best) { best = fabs(v); sg = (v < 0.0) ? -1.0 : 1.0; }
        }
        s_sign[threadIdx.x] = sg;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m * m; e += EJ_THREADS) {
        const int i = e / m, k = e - i * m;
        V_out[e] = V[i * EJ_LD + s_order[k]] * s_sign[k];
    }
    if (threadIdx.x == 0 && info) *info = sweep;
}

}  // namespace omb

extern "C" int omb_eigh_max_m(void) { return omb::EJ_MAX; }

extern "C" int omb_eigh_jacobi(const double* d_G, int64_t m, double* d_w, double* d_V, int* d_info, void* stream)
{
    using namespace omb;
    OMB_CHECK_ARG(d_G && d_w && d_V, "null pointer");
    OMB_CHECK_ARG(m >= 1 && m <= EJ_MAX, "m must be in [1, 64]");
    const size_t smem = sizeof(double) * 2 * EJ_MAX * EJ_LD;
    OMB_CUDA(cudaFuncSetAttribute(eigh_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eigh_jacobi_kernel<<<1, EJ_THREADS, smem, (cudaStream_t)stream>>>(d_G, (int)m, d_w, d_V, d_info);
    return check_launch("eigh_jacobi_kernel");
}
