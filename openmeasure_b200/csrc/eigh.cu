// eigh.cu -- S3: the m x m symmetric eigensolve of the POD Gram matrix for few snapshots (m <= 64).
//
// Replaces the bidiagonal divide-and-conquer on R inside LAPACK dgesdd (np.linalg.svd, reference
// sparse_sensing.py:272).  One CTA, matrix and eigenvectors in shared memory, cyclic two-sided
// Jacobi with the round-robin parallel ordering (M/2 disjoint rotations per step, M-1 steps per
// sweep).  A library syevd spends ~0.9 ms on an m = 41 Gram, mostly in
// host synchronisation; Jacobi also resolves the small eigenvalues of a positive semi-definite
// matrix to high relative accuracy.  Output: eigenvalues in DESCENDING order, V[i][k] = component
// i of eigenvector k.
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

constexpr int EJ_MAX = 64;
constexpr int EJ_LD = EJ_MAX + 1;
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

// the 2 x 2 block (rows p_i, q_i) x (columns p_j, q_j) of  J^T A J  for the rotations (ci, si), (cj, sj)
struct Blk { double b00, b01, b10, b11; };
__device__ __forceinline__ Blk jacobi_block(const double* __restrict__ A, int pi, int qi, int pj, int qj, double ci,
                                            double si, double cj, double sj)
{
    const double a00 = A[pi * EJ_LD + pj], a01 = A[pi * EJ_LD + qj];
    const double a10 = A[qi * EJ_LD + pj], a11 = A[qi * EJ_LD + qj];
    // columns (J_j), then rows (J_i^T)
    const double b00 = cj * a00 - sj * a01, b01 = sj * a00 + cj * a01;
    const double b10 = cj * a10 - sj * a11, b11 = sj * a10 + cj * a11;
    Blk o;
    o.b00 = ci * b00 - si * b10; o.b01 = ci * b01 - si * b11;
    o.b10 = si * b00 + ci * b10; o.b11 = si * b01 + ci * b11;
    return o;
}

// ONE barrier per step.  While every thread applies the step's rotations to its fixed work items
//   e <  npair^2 : block (pair i rows) x (pair j columns) of  A <- J^T A J   (A is ping-ponged between
//                  two shared-memory copies: the blocks of a step tile the whole matrix)
//   e >= npair^2 : two rows of  V <- V J  for one pair                       (single owner: in place)
// npair look-ahead threads (the last ones of the CTA, idle otherwise) each re-derive the three entries
// a_pq, a_pp, a_qq their pair of the NEXT step will see -- the same expressions, hence the same bits,
// as the owners of those blocks compute -- and from them the next rotation.  The rotation chain
// (~1000 cycles of dependent arithmetic) thereby runs beside the update instead of after it.
__global__ void __launch_bounds__(EJ_THREADS)
eigh_jacobi_kernel(const double* __restrict__ G, int m, double* __restrict__ w_out, double* __restrict__ V_out,
                   int* __restrict__ info)
{
    extern __shared__ double sm[];
    double* A0 = sm;                          // [EJ_MAX][EJ_LD]
    double* A1 = sm + EJ_MAX * EJ_LD;         // [EJ_MAX][EJ_LD]
    double* V = sm + 2 * EJ_MAX * EJ_LD;      // [EJ_MAX][EJ_LD]
    __shared__ double s_c[2][EJ_MAX / 2], s_s[2][EJ_MAX / 2];
    __shared__ int s_order[EJ_MAX];
    __shared__ unsigned short s_sched[(EJ_MAX - 1) * (EJ_MAX / 2)];
    __shared__ unsigned char s_slot[(EJ_MAX - 1) * EJ_MAX];    // step, index -> pair * 2 + (index is the pair's q)

    const int M = (m + 1) & ~1;             // even number of players; index m (if any) is a dummy
    const int npair = M / 2;
    for (int e = threadIdx.x; e < M * M; e += EJ_THREADS) {
        const int i = e / M, j = e - i * M;
        // symmetrise from the upper triangle so that the iteration starts exactly symmetric; the
        // dummy player of an odd m is a zero row/column that only ever meets identity rotations
        double a = 0.0;
        if (i < m && j < m) a = (i <= j) ? G[i * m + j] : G[j * m + i];
        A0[i * EJ_LD + j] = a;
        V[i * EJ_LD + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    // absolute floor for rotations: entries below eps^1.25 * max|a_ii| cannot move any eigenvalue
    // by more than that (the Gram of row-centred data is exactly rank deficient, and its null
    // direction would otherwise keep the relative criterion busy with rounding noise forever)
    __shared__ double s_floor;
    if (threadIdx.x == 0) {
        double dmax = 0.0;
        for (int i = 0; i < m; ++i) dmax = fmax(dmax, fabs(A0[i * EJ_LD + i]));
        s_floor = dmax * 1.0e-20;
    }
    // round-robin schedule: pair i of step k, stored once (no integer division in the sweeps)
    for (int e = threadIdx.x; e < (M - 1) * npair; e += EJ_THREADS) {
        const int step = e / npair, i = e - step * npair;
        int p, q;
        if (i == 0) { p = M - 1; q = step; }
        else { p = (step + i) % (M - 1); q = (step - i + (M - 1)) % (M - 1); }
        if (p > q) { const int tswap = p; p = q; q = tswap; }
        s_sched[e] = (unsigned short)((p << 8) | q);
        s_slot[step * M + p] = (unsigned char)(2 * i);
        s_slot[step * M + q] = (unsigned char)(2 * i + 1);
    }
    __syncthreads();
    const double floor_abs = s_floor, big_abs = s_floor * 1.0e7;      // 1e-13 * max|a_ii|

    // fixed work items of this thread (at most two: 2 * npair^2 <= 2048 items, 1024 threads)
    const int nblk = npair * npair, nitem = 2 * nblk;
    int it_i[2], it_j[2];                    // A item: (i, j);  V item: (npair + pair i, row pair j)
    for (int k = 0; k < 2; ++k) {
        const int e = threadIdx.x + k * EJ_THREADS;
        it_i[k] = it_j[k] = -1;
        if (e < nblk) { it_i[k] = e / npair; it_j[k] = e - it_i[k] * npair; }
        else if (e < nitem) { const int u = e - nblk; it_j[k] = u / npair; it_i[k] = npair + (u - it_j[k] * npair); }
    }
    const int la = EJ_THREADS - 1 - (int)threadIdx.x;                // look-ahead pair of this thread (if < npair)

    // rotations of the very first step
    if (la < npair) {
        const int p = s_sched[la] >> 8, q = s_sched[la] & 255;
        double c = 1.0, s = 0.0;
        bool big = false;
        if (q < m) jacobi_rot(A0[p * EJ_LD + q], A0[p * EJ_LD + p], A0[q * EJ_LD + q], floor_abs, big_abs, c, s, big);
        s_c[0][la] = c; s_s[0][la] = s;
    }
    __syncthreads();

    double* Ain = A0;
    double* Aout = A1;
    int sweep = 0, par = 0;
    int big_next = 0;            // a big rotation among those prepared for the coming step (uniform)
    {
        // the first step's own "big" flags: recompute cheaply from its rotations being non-trivial is not
        // equivalent, so evaluate the criterion once more for step 0 (only here, outside the loop)
        int b0 = 0;
        if (la < npair) {
            const int p = s_sched[la] >> 8, q = s_sched[la] & 255;
            if (q < m) {
                const double apq = A0[p * EJ_LD + q], den = fabs(A0[p * EJ_LD + p] * A0[q * EJ_LD + q]);
                b0 = (apq * apq > 1.0e-18 * den && fabs(apq) > big_abs) ? 1 : 0;
            }
        }
        big_next = __syncthreads_or(b0);
    }
    for (; sweep < EJ_MAX_SWEEPS; ++sweep) {
        int sweep_big = 0;
        for (int step = 0; step < M - 1; ++step) {
            sweep_big |= big_next;
            const unsigned short* sched = s_sched + step * npair;
            const double* cc = s_c[par];
            const double* ss = s_s[par];
            // 1. apply this step's rotations
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (it_i[k] < 0) continue;
                if (it_i[k] < npair) {
                    const int i = it_i[k], j = it_j[k];
                    const int pi = sched[i] >> 8, qi = sched[i] & 255, pj = sched[j] >> 8, qj = sched[j] & 255;
                    const Blk o = jacobi_block(Ain, pi, qi, pj, qj, cc[i], ss[i], cc[j], ss[j]);
                    Aout[pi * EJ_LD + pj] = o.b00; Aout[pi * EJ_LD + qj] = o.b01;
                    Aout[qi * EJ_LD + pj] = o.b10; Aout[qi * EJ_LD + qj] = o.b11;
                } else {
                    const int i = it_i[k] - npair, r0 = 2 * it_j[k];
                    const double s = ss[i];
                    if (s != 0.0) {
                        const double c = cc[i];
                        const int p = sched[i] >> 8, q = sched[i] & 255;
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int row = r0 + rr;
                            const double vp = V[row * EJ_LD + p], vq = V[row * EJ_LD + q];
                            V[row * EJ_LD + p] = c * vp - s * vq;
                            V[row * EJ_LD + q] = s * vp + c * vq;
                        }
                    }
                }
            }
            // 2. look-ahead: the rotation of pair `la` of the NEXT step from the entries it will see
            int my_big = 0;
            if (la < npair) {
                const int nstep = (step + 1 == M - 1) ? 0 : step + 1;
                const unsigned short pq = s_sched[nstep * npair + la];
                const int p = pq >> 8, q = pq & 255;
                double c = 1.0, s = 0.0;
                bool big = false;
                if (q < m) {
                    const unsigned char* slot = s_slot + step * M;
                    const int ip = slot[p] >> 1, rp = slot[p] & 1, iq = slot[q] >> 1, rq = slot[q] & 1;
                    const int ppi = sched[ip] >> 8, qpi = sched[ip] & 255, ppq = sched[iq] >> 8, qpq = sched[iq] & 255;
                    const double cip = cc[ip], sip = ss[ip], ciq = cc[iq], siq = ss[iq];
                    const Blk bpq = jacobi_block(Ain, ppi, qpi, ppq, qpq, cip, sip, ciq, siq);
                    const Blk bpp = jacobi_block(Ain, ppi, qpi, ppi, qpi, cip, sip, cip, sip);
                    const Blk bqq = jacobi_block(Ain, ppq, qpq, ppq, qpq, ciq, siq, ciq, siq);
                    const double apq = rp ? (rq ? bpq.b11 : bpq.b10) : (rq ? bpq.b01 : bpq.b00);
                    const double app = rp ? bpp.b11 : bpp.b00;
                    const double aqq = rq ? bqq.b11 : bqq.b00;
                    jacobi_rot(apq, app, aqq, floor_abs, big_abs, c, s, big);
                }
                s_c[par ^ 1][la] = c; s_s[par ^ 1][la] = s;
                my_big = big ? 1 : 0;
            }
            big_next = __syncthreads_or(my_big);
            double* tsw = Ain; Ain = Aout; Aout = tsw;
            par ^= 1;
        }
        if (!sweep_big) break;
    }
    const double* A = Ain;

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
    const size_t smem = sizeof(double) * 3 * EJ_MAX * EJ_LD;
    OMB_CUDA(cudaFuncSetAttribute(eigh_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eigh_jacobi_kernel<<<1, EJ_THREADS, smem, (cudaStream_t)stream>>>(d_G, (int)m, d_w, d_V, d_info);
    return check_launch("eigh_jacobi_kernel");
}
