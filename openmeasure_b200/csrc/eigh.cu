// eigh.cu -- S3: the m x m symmetric eigensolve of the POD Gram matrix for few snapshots (m <= 64).
//
// Replaces the bidiagonal divide-and-conquer on R inside LAPACK dgesdd (np.linalg.svd, reference
// sparse_sensing.py:272).  One CTA, matrix and eigenvectors in shared memory, cyclic two-sided
// Jacobi with the round-robin parallel ordering (M/2 disjoint rotations per step, M-1 steps per
// sweep), the tournament realised as a physical seat permutation.  A library syevd spends ~0.9 ms on an m = 41 Gram, mostly in
// host synchronisation; Jacobi also resolves the small eigenvalues of a positive semi-definite
// matrix to high relative accuracy.  Output: eigenvalues in DESCENDING order, V[i][k] = component
// i of eigenvector k.
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

constexpr int EJ_MAX = 64;
constexpr int EJ_LD = EJ_MAX + 2;            // even: rows stay 16-byte aligned for the 128-bit block loads
constexpr int EJ_THREADS = 1024;
constexpr int EJ_MAX_SWEEPS = 30;

// Jacobi rotation for the pivot a_pq.  Any exactly orthogonal (c, s) is a valid similarity; how well
// it annihilates a_pq only sets the convergence rate.  FP64 sqrt/div/rsqrt are ~50-instruction
// dependent chains (this step is pure latency: one warp, 21 lanes), so the angle is found in FP32
// on power-of-two-scaled inputs (relative error ~1e-7: a_pq shrinks by that factor instead of to
// zero, which leaves the quadratic convergence intact) and only the normalisation is done in FP64,
// in the rational (half-angle) form  c = (1 - u)/(1 + u), s = 2 tau/(1 + u), u = tau^2, which is
// orthogonal to rounding for ANY tau; 1/(1 + u) is an FP32 seed plus two Newton steps.
// `big`: the rotation can still move an eigenvalue at the 1e-16 * lambda_max level.
__device__ __forceinline__ void jacobi_rot(double apq, double app, double aqq, double floor_abs, double big_abs,
                                           double& c, double& s, bool& big)
{
    c = 1.0; s = 0.0; big = false;
    const double rel2 = apq * apq, den = fabs(app * aqq);
    // rotate unless |apq| <= eps * sqrt(app * aqq) (compared squared) or below the floor
    if (!(rel2 > 1.2325951644078309e-32 * den && fabs(apq) > floor_abs)) return;
    const double d = aqq - app, b = 2.0 * apq;
    const double ad = fabs(d), ab = fabs(b);
    const double mx = ad > ab ? ad : ab;
    int e = (int)((__double_as_longlong(mx) >> 52) & 0x7FF);
    e = e < 1 ? 1 : (e > 2045 ? 2045 : e);
    const double sc = __longlong_as_double((long long)(2046 - e) << 52);      // 2^(1023 - e): max(|d|,|b|) -> [1, 2)
    const float df = (float)(d * sc), bf = (float)(b * sc);
    const float hf = sqrtf(fmaf(df, df, bf * bf));
    const float tf = __fdividef(bf, df + copysignf(hf, df));                 // tan(theta), |theta| <= pi/4
    // (|b| < 1e-38 |d| underflows to tf = 0: such a rotation is the identity to FP64 rounding anyway)
    const double tau = (double)__fdividef(tf, 1.0f + sqrtf(fmaf(tf, tf, 1.0f)));   // tan(theta / 2)
    const double u = tau * tau, w = 1.0 + u;
    double y = (double)__frcp_rn((float)w);
    double r1 = fma(-w, y, 1.0);
    y = fma(y, r1, y);
    r1 = fma(-w, y, 1.0);
    y = fma(y, r1, y);
    c = (1.0 - u) * y;
    s = (tau + tau) * y;
    big = rel2 > 1.0e-18 * den && fabs(apq) > big_abs;
}

// Round-robin tournament as a PHYSICAL permutation: pair k always sits at positions (2k, 2k+1) of the
// rows/columns of A and the columns of V, and after every step the players move one seat (t_0 fixed,
// t_k -> t_{k+1}, t_{n-1} -> b_{n-1}, b_k -> b_{k-1}, b_0 -> t_1; t_k = 2k, b_k = 2k+1).  The seat map
// is the same every step, so a thread's source block (two 128-bit loads, consecutive threads =
// consecutive columns: no bank conflicts) and its four destination entries are fixed for the whole
// kernel: no schedule tables, no index arithmetic, no scattered gathers inside the sweeps.
__device__ __forceinline__ int seat_next(int x, int n)
{
    if (n == 1 || x == 0) return x;
    if (x == 1) return 2;
    if (x & 1) return x - 2;
    if (x == 2 * n - 2) return 2 * n - 1;
    return x + 2;
}
__device__ __forceinline__ int seat_prev(int y, int n)
{
    if (n == 1 || y == 0) return y;
    if (y == 2) return 1;
    if (!(y & 1)) return y - 2;
    if (y == 2 * n - 1) return 2 * n - 2;
    return y + 2;
}

// the 2 x 2 block (pair i rows) x (pair j columns) of  J^T A J
struct Blk { double b00, b01, b10, b11; };
__device__ __forceinline__ Blk jacobi_block(const double* __restrict__ A, int i, int j, double2 ri, double2 rj)
{
    const double2 u0 = *reinterpret_cast<const double2*>(A + (2 * i) * EJ_LD + 2 * j);
    const double2 u1 = *reinterpret_cast<const double2*>(A + (2 * i + 1) * EJ_LD + 2 * j);
    const double ci = ri.x, si = ri.y, cj = rj.x, sj = rj.y;
    // columns (J_j), then rows (J_i^T)
    const double b00 = cj * u0.x - sj * u0.y, b01 = sj * u0.x + cj * u0.y;
    const double b10 = cj * u1.x - sj * u1.y, b11 = sj * u1.x + cj * u1.y;
    Blk o;
    o.b00 = ci * b00 - si * b10; o.b01 = ci * b01 - si * b11;
    o.b10 = si * b00 + ci * b10; o.b11 = si * b01 + ci * b11;
    return o;
}

// ONE barrier per step.  While every thread applies the step's rotations to its fixed work items
//   e <  npair^2 : block (pair i) x (pair j) of  A <- J^T A J, written to its next seats
//   e >= npair^2 : two rows of  V <- V J  for one pair, written to the pair's next column seats
// (A and V are ping-ponged: a step rewrites both completely), npair look-ahead threads -- the last
// ones of the CTA, idle otherwise -- each re-derive the three entries a_pp, a_pq, a_qq that the pair
// of the NEXT step will find at its seats (the same expressions, hence the same bits, as the owners
// of those blocks compute) and from them the next rotation.  The rotation chain (~1000 cycles of
// dependent arithmetic) thereby runs beside the update instead of after it.
__global__ void __launch_bounds__(EJ_THREADS)
eigh_jacobi_kernel(const double* __restrict__ G, int m, double* __restrict__ w_out, double* __restrict__ V_out,
                   int* __restrict__ info)
{
    extern __shared__ __align__(16) double sm[];
    double* Ain = sm;                               // [EJ_MAX][EJ_LD] x 2 (A), x 2 (V)
    double* Aout = sm + EJ_MAX * EJ_LD;
    double* Vin = sm + 2 * EJ_MAX * EJ_LD;
    double* Vout = sm + 3 * EJ_MAX * EJ_LD;
    __shared__ double2 s_rot[2][EJ_MAX / 2];        // (c, s) of every pair, double buffered
    __shared__ short s_perm[2][EJ_MAX];             // seat -> original index
    __shared__ int s_order[EJ_MAX];
    __shared__ double s_sign[EJ_MAX];
    __shared__ double s_floor;

    const int M = (m + 1) & ~1;             // even number of players; original index m (if any) is a dummy
    const int npair = M / 2;
    for (int e = threadIdx.x; e < M * M; e += EJ_THREADS) {
        const int i = e / M, j = e - i * M;
        // symmetrise from the upper triangle so that the iteration starts exactly symmetric; the
        // dummy player of an odd m is a zero row/column that only ever meets identity rotations
        double a = 0.0;
        if (i < m && j < m) a = (i <= j) ? G[i * m + j] : G[j * m + i];
        Ain[i * EJ_LD + j] = a;
        Vin[i * EJ_LD + j] = (i == j) ? 1.0 : 0.0;
    }
    if (threadIdx.x < M) s_perm[0][threadIdx.x] = (short)threadIdx.x;
    __syncthreads();
    // absolute floor for rotations: entries below eps^1.25 * max|a_ii| cannot move any eigenvalue
    // by more than that (the Gram of row-centred data is exactly rank deficient, and its null
    // direction would otherwise keep the relative criterion busy with rounding noise forever)
    if (threadIdx.x == 0) {
        double dmax = 0.0;
        for (int i = 0; i < m; ++i) dmax = fmax(dmax, fabs(Ain[i * EJ_LD + i]));
        s_floor = dmax * 1.0e-20;
    }
    __syncthreads();
    const double floor_abs = s_floor, big_abs = s_floor * 1.0e7;      // 1e-13 * max|a_ii|

    // fixed work items of this thread (at most two: 2 * npair^2 <= 2048 items, 1024 threads)
    const int nblk = npair * npair, nitem = 2 * nblk;
    int it_i[2], it_j[2], it_d0[2], it_d1[2], it_d2[2], it_d3[2];
    for (int k = 0; k < 2; ++k) {
        const int e = threadIdx.x + k * EJ_THREADS;
        it_i[k] = it_j[k] = -1;
        it_d0[k] = it_d1[k] = it_d2[k] = it_d3[k] = 0;
        if (e < nblk) {                       // A item: block (i, j) -> seats (next(2i), next(2i+1)) x (next(2j), next(2j+1))
            const int i = e / npair, j = e - i * npair;
            it_i[k] = i; it_j[k] = j;
            const int r0 = seat_next(2 * i, npair), r1 = seat_next(2 * i + 1, npair);
            const int c0 = seat_next(2 * j, npair), c1 = seat_next(2 * j + 1, npair);
            it_d0[k] = r0 * EJ_LD + c0; it_d1[k] = r0 * EJ_LD + c1;
            it_d2[k] = r1 * EJ_LD + c0; it_d3[k] = r1 * EJ_LD + c1;
        } else if (e < nitem) {               // V item: rows (2 rp, 2 rp + 1), pair i -> column seats of pair i
            const int u = e - nblk, rp = u / npair, i = u - rp * npair;
            it_i[k] = npair + i; it_j[k] = rp;
            const int c0 = seat_next(2 * i, npair), c1 = seat_next(2 * i + 1, npair);
            it_d0[k] = (2 * rp) * EJ_LD + c0; it_d1[k] = (2 * rp) * EJ_LD + c1;
            it_d2[k] = (2 * rp + 1) * EJ_LD + c0; it_d3[k] = (2 * rp + 1) * EJ_LD + c1;
        }
    }
    // look-ahead: pair `la` of the next step sits at seats (2 la, 2 la + 1) then, i.e. at x, y now
    const int la = EJ_THREADS - 1 - (int)threadIdx.x;
    const bool is_la = la < npair;
    const int lx = seat_prev(2 * (is_la ? la : 0), npair), ly = seat_prev(2 * (is_la ? la : 0) + 1, npair);
    const int kx = lx >> 1, rx = lx & 1, ky = ly >> 1, ry = ly & 1;

    // rotations of the very first step
    int b0 = 0;
    if (is_la) {
        double c = 1.0, s = 0.0;
        bool big = false;
        if (2 * la + 1 < m)                 // seats == original indices before the first move; index m is the dummy
            jacobi_rot(Ain[(2 * la) * EJ_LD + 2 * la + 1], Ain[(2 * la) * EJ_LD + 2 * la],
                       Ain[(2 * la + 1) * EJ_LD + 2 * la + 1], floor_abs, big_abs, c, s, big);
        s_rot[0][la] = make_double2(c, s);
        b0 = big ? 1 : 0;
    }
    int big_next = __syncthreads_or(b0);    // a big rotation among those prepared for the coming step (uniform)

    int sweep = 0, par = 0;
    for (; sweep < EJ_MAX_SWEEPS; ++sweep) {
        int sweep_big = 0;
        for (int step = 0; step < M - 1; ++step) {
            sweep_big |= big_next;
            const double2* rot = s_rot[par];
            // 1. apply this step's rotations
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (it_i[k] < 0) continue;
                if (it_i[k] < npair) {
                    const Blk o = jacobi_block(Ain, it_i[k], it_j[k], rot[it_i[k]], rot[it_j[k]]);
                    Aout[it_d0[k]] = o.b00; Aout[it_d1[k]] = o.b01;
                    Aout[it_d2[k]] = o.b10; Aout[it_d3[k]] = o.b11;
                } else {
                    const int i = it_i[k] - npair, r0 = 2 * it_j[k];
                    const double2 cs = rot[i];
                    const double2 v0 = *reinterpret_cast<const double2*>(Vin + r0 * EJ_LD + 2 * i);
                    const double2 v1 = *reinterpret_cast<const double2*>(Vin + (r0 + 1) * EJ_LD + 2 * i);
                    Vout[it_d0[k]] = cs.x * v0.x - cs.y * v0.y; Vout[it_d1[k]] = cs.y * v0.x + cs.x * v0.y;
                    Vout[it_d2[k]] = cs.x * v1.x - cs.y * v1.y; Vout[it_d3[k]] = cs.y * v1.x + cs.x * v1.y;
                }
            }
            // 2. seats -> original indices follow the players
            if (threadIdx.x < M) s_perm[par ^ 1][seat_next(threadIdx.x, npair)] = s_perm[par][threadIdx.x];
            // 3. look-ahead: the rotation of pair `la` of the NEXT step from the entries it will see
            int my_big = 0;
            if (is_la) {
                double c = 1.0, s = 0.0;
                bool big = false;
                if (s_perm[par][lx] < m && s_perm[par][ly] < m) {
                    const double2 rkx = rot[kx], rky = rot[ky];
                    const Blk bpq = jacobi_block(Ain, kx, ky, rkx, rky);
                    const Blk bpp = jacobi_block(Ain, kx, kx, rkx, rkx);
                    const Blk bqq = jacobi_block(Ain, ky, ky, rky, rky);
                    const double apq = rx ? (ry ? bpq.b11 : bpq.b10) : (ry ? bpq.b01 : bpq.b00);
                    const double app = rx ? bpp.b11 : bpp.b00;
                    const double aqq = ry ? bqq.b11 : bqq.b00;
                    jacobi_rot(apq, app, aqq, floor_abs, big_abs, c, s, big);
                }
                s_rot[par ^ 1][la] = make_double2(c, s);
                my_big = big ? 1 : 0;
            }
            big_next = __syncthreads_or(my_big);
            double* tsw = Ain; Ain = Aout; Aout = tsw;
            tsw = Vin; Vin = Vout; Vout = tsw;
            par ^= 1;
        }
        if (!sweep_big) break;
    }
    const double* A = Ain;
    const double* V = Vin;
    const short* perm = s_perm[par];

    // descending order by rank counting over the seats of real players (ties broken by seat)
    if (threadIdx.x < M && perm[threadIdx.x] < m) {
        const int i = threadIdx.x;
        const double wi = A[i * EJ_LD + i];
        int rank = 0;
        for (int j = 0; j < M; ++j) {
            if (perm[j] >= m) continue;
            const double wj = A[j * EJ_LD + j];
            rank += (wj > wi) || (wj == wi && j < i);
        }
        s_order[rank] = i;
    }
    __syncthreads();
    if (threadIdx.x < m) w_out[threadIdx.x] = A[s_order[threadIdx.x] * EJ_LD + s_order[threadIdx.x]];
    // deterministic sign: the largest-magnitude component of every eigenvector is positive
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

// sigma = sqrt(max(lambda, 0)) and the back-projection weights W = V diag(1/sigma), with a zero weight
// for modes whose singular value is numerically zero (row-centred data has rank m - 1): one launch
// instead of a dozen framework element-wise kernels between the eigensolve and the back-projection.
__global__ void __launch_bounds__(256)
pod_weights_kernel(const double* __restrict__ w, const double* __restrict__ V, int m, double rel_floor,
                   double* __restrict__ S, double* __restrict__ W)
{
    // thread = mode q (its square root and reciprocal are formed once), CTAs stride over the rows of V
    const double s0 = sqrt(fmax(w[0], 0.0));
    for (int q = threadIdx.x; q < m; q += 256) {
        const double sq = sqrt(fmax(w[q], 0.0));
        const bool live = sq > s0 * rel_floor;
        const double inv = live ? 1.0 / sq : 0.0;
        if (blockIdx.x == 0) S[q] = sq;
        for (int row = blockIdx.x; row < m; row += gridDim.x) {
            const int64_t e = (int64_t)row * m + q;
            W[e] = live ? V[e] * inv : 0.0;
        }
    }
}

}  // namespace omb

extern "C" int omb_eigh_max_m(void) { return omb::EJ_MAX; }

extern "C" int omb_eigh_jacobi(const double* d_G, int64_t m, double* d_w, double* d_V, int* d_info, void* stream)
{
    using namespace omb;
    OMB_CHECK_ARG(d_G && d_w && d_V, "null pointer");
    OMB_CHECK_ARG(m >= 1 && m <= EJ_MAX, "m must be in [1, 64]");
    const size_t smem = sizeof(double) * 4 * EJ_MAX * EJ_LD;
    OMB_CUDA(cudaFuncSetAttribute(eigh_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eigh_jacobi_kernel<<<1, EJ_THREADS, smem, (cudaStream_t)stream>>>(d_G, (int)m, d_w, d_V, d_info);
    return check_launch("eigh_jacobi_kernel");
}

extern "C" int omb_pod_weights(const double* d_w, const double* d_V, int64_t m, double rel_floor, double* d_S, double* d_W,
                               void* stream)
{
    using namespace omb;
    OMB_CHECK_ARG(d_w && d_V && d_S && d_W, "null pointer");
    OMB_CHECK_ARG(m >= 1 && m <= 4096, "m must be in [1, 4096]");
    const int grid = (int)(m < 2 * (int64_t)sm_count() ? m : 2 * (int64_t)sm_count());
    pod_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_w, d_V, (int)m, rel_floor, d_S, d_W);
    return check_launch("pod_weights_kernel");
}
