// qrcp.cu -- K6: QR with column pivoting over the n candidate sensor locations.
//
// Replaces scipy.linalg.qr(self.Ur.T, pivoting=True, mode='economic') (reference
// sparse_sensing.py:739), i.e. LAPACK dgeqp3 -> dlaqp2 (+ dlarfg / dlarf) on the r x n matrix
// A = Ur^T, of which the reference keeps only the first r pivots (:741-743).
//
// Layout: A is mode-major (r rows of ld doubles), so the n candidates are the coalesced axis and
// step i only touches rows i..r-1.  One thread owns one candidate column per step; the only
// cross-thread work is the (norm, LAPACK-position) argmax.  Per step:
//
//   panel kernel (1 CTA)   reduce the per-CTA argmax records -> pivot p; LAPACK's column-swap
//                          bookkeeping (position keys for tie-breaking); gather p's trailing
//                          column; dlarfg -> (v, tau, beta); for blocked runs the compact-WY T
//                          and the row vector q = Q e_t.
//   pass kernel (grid)     block == 1 or last step of a block: apply the block's reflectors to
//                          every column (dlarf arithmetic, sequential fma), write the rows below
//                          the block, down-date vn1/vn2 exactly like dlaqp2 (tol3z guard and norm
//                          recomputation), emit the next argmax records.
//                          inside a block: read-only GEMV  R[i, j] = q . A[i0:, j]  + down-date +
//                          argmax: the trailing matrix is streamed once and never written.
//
// With block == 1 the arithmetic is dlaqp2's, operation for operation, in the order fixed by
// oracle/csrc/oracle.c (bit-identical R diagonal, norms and pivots).  block > 1 moves
// ~ (1 + 1/block)/2 of the bytes; pivots are identical on non-degenerate inputs (the degeneracy
// meter d_gap reports how close any decision was).
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

constexpr int QR_RMAX = 256;     // max modes (rows of A)
constexpr int QR_BMAX = 16;      // max steps per block
constexpr int QR_NCAND = 4096;   // max CTAs of a pass kernel (argmax records)
constexpr double QR_TOL3Z = 1.0536712127723509e-08;   // sqrt(2^-53), LAPACK tol3z

struct Cand {
    double best;      // largest partial column norm (-1: none)
    double second;    // second largest (-1: none)
    int64_t idx;      // local column index of best
    int64_t key;      // LAPACK position of best (ties -> lowest position wins, as idamax)
};

struct Panel {
    double V[QR_BMAX][QR_RMAX];   // in-block reflectors over rows i0.., V[t][k] = 0 (k<t), 1 (k==t)
    double T[QR_BMAX][QR_BMAX];   // compact-WY factor, Q = H_0 ... H_t = I - V T V^T
    double tau[QR_BMAX];
    double q[QR_RMAX];            // q = Q e_t   ->  R[i, j] = q . A[i0:, j]
    int64_t posmap[QR_RMAX];      // current LAPACK position of original column c < s
    int64_t col_at_pos[QR_RMAX];  // original column now at position k < s
};

__device__ __forceinline__ bool cand_better(double b1, int64_t k1, double b2, int64_t k2)
{
    return b1 > b2 || (b1 == b2 && k1 < k2);
}
__device__ __forceinline__ void cand_merge(Cand& a, const Cand& b)
{
    if (cand_better(b.best, b.key, a.best, a.key)) {
        double s = fmax(a.best, b.second);
        a.best = b.best; a.idx = b.idx; a.key = b.key; a.second = s;
    } else {
        a.second = fmax(a.second, b.best);
    }
}
__device__ __forceinline__ void cand_push(Cand& a, double v, int64_t idx, int64_t key)
{
    if (cand_better(v, key, a.best, a.key)) { a.second = a.best; a.best = v; a.idx = idx; a.key = key; }
    else a.second = fmax(a.second, v);
}
__device__ __forceinline__ Cand cand_shfl_xor(const Cand& a, int o)
{
    Cand b;
    b.best = __shfl_xor_sync(0xFFFFFFFFu, a.best, o);
    b.second = __shfl_xor_sync(0xFFFFFFFFu, a.second, o);
    b.idx = __shfl_xor_sync(0xFFFFFFFFu, a.idx, o);
    b.key = __shfl_xor_sync(0xFFFFFFFFu, a.key, o);
    return b;
}
// CTA-wide reduction; result valid in thread 0.  s_c must hold blockDim.x/32 records.
__device__ __forceinline__ Cand cand_block_reduce(Cand c, Cand* s_c)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { Cand b = cand_shfl_xor(c, o); cand_merge(c, b); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_c[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) cand_merge(c, s_c[w]);
    return c;
}

// dlaqp2's partial-norm down-date for one column.  rij = R[i, j].  Returns true when LAPACK would
// recompute the norm from scratch (caller supplies it).
__device__ __forceinline__ bool downdate(double rij, double& v1, double v2)
{
    double qv = fabs(rij) / v1;
    double temp = 1.0 - qv * qv;
    temp = fmax(temp, 0.0);
    double q2 = v1 / v2;
    double temp2 = temp * (q2 * q2);
    if (temp2 <= QR_TOL3Z) return true;
    v1 = v1 * sqrt(temp);
    return false;
}

// ---------------------------------------------------------------------------------------------
// initial norms (when the back-projection did not supply them): sequential fma over the modes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
qr_norms_kernel(const double* __restrict__ A, int64_t ld, int64_t n, int r, double* __restrict__ vn)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < r; ++k) { double x = ldg_stream(A + (int64_t)k * ld + j); s = fma(x, x, s); }
        vn[j] = sqrt(s);
    }
}

__global__ void __launch_bounds__(256)
qr_init_kernel(const double* __restrict__ vn, int64_t n, double* __restrict__ vn1, double* __restrict__ vn2,
               Panel* __restrict__ P)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double v = vn[j];
        vn1[j] = v;
        vn2[j] = v;
    }
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < QR_RMAX; k += blockDim.x) { P->posmap[k] = k; P->col_at_pos[k] = k; }
}

// ---------------------------------------------------------------------------------------------
// slow path of the read-only pass: exact trailing norm of one column after reflectors 0..t
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ double recompute_norm(const double* __restrict__ src, int64_t ld, int64_t j, int L, int t,
                                              const double* __restrict__ Vg, const double* __restrict__ taug)
{
    double c[QR_RMAX];
    for (int k = 0; k < L; ++k) c[k] = src[(int64_t)k * ld + j];
    for (int tt = 0; tt <= t; ++tt) {
        const double* v = Vg + tt * QR_RMAX;
        double w = c[tt];
        for (int k = tt + 1; k < L; ++k) w = fma(v[k], c[k], w);
        const double tw = taug[tt] * w;
        c[tt] -= tw;
        for (int k = tt + 1; k < L; ++k) c[k] = fma(-tw, v[k], c[k]);
    }
    double s = 0.0;
    for (int k = t + 1; k < L; ++k) s = fma(c[k], c[k], s);
    return sqrt(s);
}

// ---------------------------------------------------------------------------------------------
// read-only pass: R[i, j] = q . A[i0:, j], down-date, argmax.  L = r - i0 rows (0 = argmax only).
// Two adjacent columns per thread (128-bit loads).
// ---------------------------------------------------------------------------------------------
constexpr int GV_THREADS = 256;

__global__ void __launch_bounds__(GV_THREADS)
qr_gemv_kernel(const double* __restrict__ src, int64_t ld, int64_t n, int L, int t, int last_row,
               const Panel* __restrict__ P, double* __restrict__ vn1, double* __restrict__ vn2,
               int64_t s_total, Cand* __restrict__ cand)
{
    __shared__ double s_q[QR_RMAX];
    __shared__ Cand s_c[GV_THREADS / 32];
    for (int k = threadIdx.x; k < L; k += GV_THREADS) s_q[k] = P->q[k];
    __syncthreads();

    Cand best;
    best.best = -1.0; best.second = -1.0; best.idx = -1; best.key = INT64_MAX;

    const int64_t npairs = (n + 1) >> 1;
    for (int64_t pr = (int64_t)blockIdx.x * GV_THREADS + threadIdx.x; pr < npairs;
         pr += (int64_t)gridDim.x * GV_THREADS) {
        const int64_t j = pr * 2;
        double y0 = 0.0, y1 = 0.0;
        const double* col = src + j;
        int k = 0;
        for (; k + 8 <= L; k += 8) {
            double2 a[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = ldg_stream2(col + (int64_t)(k + u) * ld);
#pragma unroll
            for (int u = 0; u < 8; ++u) { y0 = fma(s_q[k + u], a[u].x, y0); y1 = fma(s_q[k + u], a[u].y, y1); }
        }
        for (; k < L; ++k) {
            double2 a = ldg_stream2(col + (int64_t)k * ld);
            y0 = fma(s_q[k], a.x, y0);
            y1 = fma(s_q[k], a.y, y1);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int64_t jj = j + e;
            if (jj >= n) break;
            double v1 = vn1[jj];
            if (v1 < 0.0) continue;                 // already a pivot
            if (v1 != 0.0 && L > 0) {
                const double v2 = vn2[jj];
                if (downdate(e ? y1 : y0, v1, v2)) {
                    v1 = last_row ? 0.0 : recompute_norm(src, ld, jj, L, t, &P->V[0][0], P->tau);
                    vn2[jj] = v1;
                }
                vn1[jj] = v1;
            }
            const int64_t key = jj < s_total ? P->posmap[jj] : jj;
            cand_push(best, v1, jj, key);
        }
    }
    best = cand_block_reduce(best, s_c);
    if (threadIdx.x == 0) cand[blockIdx.x] = best;
}

// ---------------------------------------------------------------------------------------------
// apply pass: apply reflectors 0..t of the current block to every column (dlarf arithmetic),
// write rows t+1.. of the block-local tail to dst, down-date, argmax.
// One thread per column; the column tail lives in shared memory (per-thread private slice).
// ---------------------------------------------------------------------------------------------
constexpr int AP_THREADS = 128;

__global__ void __launch_bounds__(AP_THREADS)
qr_apply_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t ld, int64_t n, int L, int t,
                const Panel* __restrict__ P, double* __restrict__ vn1, double* __restrict__ vn2,
                int64_t s_total, Cand* __restrict__ cand)
{
    extern __shared__ double sm[];
    double* s_V = sm;                                   // (t+1) x L
    double* s_col = sm + (size_t)(t + 1) * L;           // L x AP_THREADS
    __shared__ double s_tau[QR_BMAX];
    __shared__ Cand s_c[AP_THREADS / 32];
    for (int e = threadIdx.x; e < (t + 1) * L; e += AP_THREADS) {
        const int tt = e / L, k = e - tt * L;
        s_V[e] = P->V[tt][k];
    }
    if (threadIdx.x <= t) s_tau[threadIdx.x] = P->tau[threadIdx.x];
    __syncthreads();

    Cand best;
    best.best = -1.0; best.second = -1.0; best.idx = -1; best.key = INT64_MAX;
    double* c = s_col + threadIdx.x;
    const bool last_row = (t + 1 == L);

    for (int64_t j0 = (int64_t)blockIdx.x * AP_THREADS; j0 < n; j0 += (int64_t)gridDim.x * AP_THREADS) {
        const int64_t j = j0 + threadIdx.x;
        if (j < n) {
            const double* col = src + j;
            int k = 0;
            for (; k + 8 <= L; k += 8) {
                double a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = ldg_stream(col + (int64_t)(k + u) * ld);
#pragma unroll
                for (int u = 0; u < 8; ++u) c[(k + u) * AP_THREADS] = a[u];
            }
            for (; k < L; ++k) c[k * AP_THREADS] = ldg_stream(col + (int64_t)k * ld);

            for (int tt = 0; tt <= t; ++tt) {
                const double* v = s_V + tt * L;
                double w = c[tt * AP_THREADS];
                for (int kk = tt + 1; kk < L; ++kk) w = fma(v[kk], c[kk * AP_THREADS], w);
                const double tw = s_tau[tt] * w;
                c[tt * AP_THREADS] -= tw;
                for (int kk = tt + 1; kk < L; ++kk) c[kk * AP_THREADS] = fma(-tw, v[kk], c[kk * AP_THREADS]);
            }
            double* out = dst + j;
            for (int kk = t + 1; kk < L; ++kk) stg_stream(out + (int64_t)kk * ld, c[kk * AP_THREADS]);

            double v1 = vn1[j];
            if (v1 >= 0.0) {
                if (v1 != 0.0) {
                    const double v2 = vn2[j];
                    if (downdate(c[t * AP_THREADS], v1, v2)) {
                        double s = 0.0;
                        for (int kk = t + 1; kk < L; ++kk) { const double x = c[kk * AP_THREADS]; s = fma(x, x, s); }
                        v1 = last_row ? 0.0 : sqrt(s);
                        vn2[j] = v1;
                    }
                    vn1[j] = v1;
                }
                const int64_t key = j < s_total ? P->posmap[j] : j;
                cand_push(best, v1, j, key);
            }
        }
    }
    best = cand_block_reduce(best, s_c);
    if (threadIdx.x == 0) cand[blockIdx.x] = best;
}

// ---------------------------------------------------------------------------------------------
// panel kernel (1 CTA): pivot selection + reflector for global step i (block-local index t)
// ---------------------------------------------------------------------------------------------
constexpr int PN_THREADS = 256;

__device__ __forceinline__ double block_sum(double x, double* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < PN_THREADS / 32; ++w) s += s_red[w];
    return s;
}

__global__ void __launch_bounds__(PN_THREADS)
qr_panel_kernel(Panel* __restrict__ P, const Cand* __restrict__ cand, int ncand, const double* __restrict__ src,
                int64_t ld, int L, int i, int t, int seq_norm, int64_t s_total, int64_t index_base,
                double* __restrict__ vn1, int64_t* __restrict__ piv, double* __restrict__ rdiag,
                double* __restrict__ gap)
{
    __shared__ Cand s_c[PN_THREADS / 32];
    __shared__ double s_x[QR_RMAX];      // pivot column tail over rows i0..
    __shared__ double s_z[QR_BMAX];
    __shared__ double s_red[PN_THREADS / 32];
    __shared__ int64_t s_p;
    __shared__ double s_beta, s_tau, s_scal;

    // 1. global argmax over the pass kernel's records
    Cand c;
    c.best = -1.0; c.second = -1.0; c.idx = -1; c.key = INT64_MAX;
    for (int e = threadIdx.x; e < ncand; e += PN_THREADS) cand_merge(c, cand[e]);
    c = cand_block_reduce(c, s_c);
    if (threadIdx.x == 0) {
        const int64_t p = c.idx;
        s_p = p;
        piv[i] = p + index_base;
        gap[i] = (c.second < 0.0 || c.best <= 0.0) ? 1.0 : (c.best - c.second) / c.best;
        // LAPACK's swap of positions i <-> pos(p): the column sitting at position i moves to pos(p)
        const int64_t pos_p = c.key;
        if (pos_p != i) {
            const int64_t ci = P->col_at_pos[i];
            P->posmap[ci] = pos_p;
            if (pos_p < s_total) P->col_at_pos[pos_p] = ci;
        }
        P->col_at_pos[i] = p;
        if (p < s_total) P->posmap[p] = i;
        vn1[p] = -1.0;                   // never a candidate again
    }
    __syncthreads();
    const int64_t p = s_p;

    // 2. pivot column tail at block start, then the block's earlier reflectors (compact WY)
    for (int k = threadIdx.x; k < L; k += PN_THREADS) s_x[k] = src[(int64_t)k * ld + p];
    __syncthreads();
    if (t > 0) {
        // z = V^T x  (one warp per reflector), z' = T^T z, x -= V z'
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int tt = warp; tt < t; tt += PN_THREADS / 32) {
            double s = 0.0;
            for (int k = lane; k < L; k += 32) s = fma(P->V[tt][k], s_x[k], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
            if (lane == 0) s_z[tt] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double zz[QR_BMAX];
            for (int a = 0; a < t; ++a) {
                double s = 0.0;
                for (int b = 0; b <= a; ++b) s = fma(P->T[b][a], s_z[b], s);   // (T^T z)_a
                zz[a] = s;
            }
            for (int a = 0; a < t; ++a) s_z[a] = zz[a];
        }
        __syncthreads();
        for (int k = threadIdx.x; k < L; k += PN_THREADS) {
            double x = s_x[k];
            for (int a = 0; a < t; ++a) x = fma(-P->V[a][k], s_z[a], x);
            s_x[k] = x;
        }
        __syncthreads();
    }

    // 3. dlarfg on x[t:]  (alpha = x[t], tail x[t+1:])
    double xn2;
    if (seq_norm) {
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int k = t + 1; k < L; ++k) s = fma(s_x[k], s_x[k], s);
            s_red[0] = s;
        }
        __syncthreads();
        xn2 = s_red[0];
        __syncthreads();
    } else {
        double s = 0.0;
        for (int k = t + 1 + threadIdx.x; k < L; k += PN_THREADS) s = fma(s_x[k], s_x[k], s);
        xn2 = block_sum(s, s_red);
    }
    if (threadIdx.x == 0) {
        const double alpha = s_x[t];
        const double xnorm = sqrt(xn2);
        double beta = alpha, tau = 0.0, scal = 0.0;
        if (L - t > 1 && xnorm != 0.0) {
            // dlapy2(alpha, xnorm)
            const double xa = fabs(alpha), ya = xnorm;
            const double w = fmax(xa, ya), z = fmin(xa, ya);
            double h = w;
            if (z != 0.0) { const double qq = z / w; h = w * sqrt(1.0 + qq * qq); }
            beta = -copysign(h, alpha);
            tau = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
        s_beta = beta; s_tau = tau; s_scal = scal;
        rdiag[i] = beta;
        P->tau[t] = tau;
    }
    __syncthreads();
    const double tau = s_tau, scal = s_scal;
    for (int k = threadIdx.x; k < L; k += PN_THREADS) {
        double v = 0.0;
        if (k == t) v = 1.0;
        else if (k > t) v = (tau != 0.0) ? s_x[k] * scal : 0.0;
        P->V[t][k] = v;
        s_x[k] = v;                      // s_x now holds v_t
    }
    __syncthreads();

    // 4. compact WY column t of T and q = Q e_t = e_t - V T (V^T e_t)   (blocked runs only use q)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int tt = warp; tt < t; tt += PN_THREADS / 32) {
            double s = 0.0;
            for (int k = lane; k < L; k += 32) s = fma(P->V[tt][k], s_x[k], s);     // V[:, tt]^T v_t
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
            if (lane == 0) s_z[tt] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // T[0:t, t] = -tau * T[0:t, 0:t] * s_z ; T[t][t] = tau
            double col[QR_BMAX];
            for (int a = 0; a < t; ++a) {
                double s = 0.0;
                for (int b = a; b < t; ++b) s = fma(P->T[a][b], s_z[b], s);
                col[a] = -tau * s;
            }
            for (int a = 0; a < t; ++a) P->T[a][t] = col[a];
            P->T[t][t] = tau;
            // g = T * (V^T e_t) with (V^T e_t)_a = V[a][t]  (a <= t)
            for (int a = 0; a <= t; ++a) {
                double s = 0.0;
                for (int b = a; b <= t; ++b) s = fma(P->T[a][b], (b == t) ? 1.0 : P->V[b][t], s);
                s_z[a] = s;
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < L; k += PN_THREADS) {
            double qv = (k == t) ? 1.0 : 0.0;
            for (int a = 0; a < t; ++a) qv = fma(-P->V[a][k], s_z[a], qv);
            qv = fma(-s_x[k], s_z[t], qv);
            P->q[k] = qv;
        }
    }
}

struct QrWs {
    double* vn1;
    double* vn2;
    Cand* cand;
    Panel* panel;
    double* vn_tmp;
};

static int64_t qr_ws_layout(int64_t n, QrWs* w, char* base)
{
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off += round_up(bytes, 256); return base ? base + o : (char*)nullptr; };
    char* a = take((int64_t)sizeof(double) * (n + 2));
    char* b = take((int64_t)sizeof(double) * (n + 2));
    char* c = take((int64_t)sizeof(Cand) * QR_NCAND);
    char* d = take((int64_t)sizeof(Panel));
    char* e = take((int64_t)sizeof(double) * (n + 2));
    if (w) { w->vn1 = (double*)a; w->vn2 = (double*)b; w->cand = (Cand*)c; w->panel = (Panel*)d; w->vn_tmp = (double*)e; }
    return off;
}

}  // namespace omb

using namespace omb;

extern "C" int64_t omb_qrcp_ws_bytes(int64_t n, int64_t r)
{
    (void)r;
    if (n <= 0) return 0;
    return qr_ws_layout(n, nullptr, nullptr);
}

extern "C" int omb_qrcp(const double* d_Ut, int64_t ld, int64_t n, int64_t r, int64_t s, const double* d_vn,
                        double* d_work, void* d_ws, int block, int64_t index_base, int64_t* d_piv,
                        double* d_rdiag, double* d_gap, void* stream)
{
    OMB_CHECK_ARG(d_Ut && d_work && d_ws && d_piv && d_rdiag && d_gap, "null pointer");
    OMB_CHECK_ARG(n > 0 && r > 0 && s > 0, "non-positive size");
    OMB_CHECK_ARG(r <= QR_RMAX, "r exceeds the supported number of modes (256)");
    OMB_CHECK_ARG(s <= r && s <= n, "s must be <= min(r, n)");
    OMB_CHECK_ARG(ld >= n && (ld % 2) == 0, "ld must be even and >= n");
    OMB_CHECK_ARG(block >= 1 && block <= QR_BMAX, "block must be in [1, 16]");
    OMB_CHECK_ARG((((uintptr_t)d_Ut | (uintptr_t)d_work) & 15) == 0, "basis pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    QrWs w;
    qr_ws_layout(n, &w, (char*)d_ws);
    const int sms = sm_count();
    int rc;

    const double* vn = d_vn;
    if (!vn) {
        int64_t g = ceil_div(n, 256);
        if (g > (int64_t)sms * 8) g = (int64_t)sms * 8;
        qr_norms_kernel<<<(unsigned)g, 256, 0, st>>>(d_Ut, ld, n, (int)r, w.vn_tmp);
        if ((rc = check_launch("qr_norms_kernel"))) return rc;
        vn = w.vn_tmp;
    }
    {
        int64_t g = ceil_div(n, 256);
        if (g > (int64_t)sms * 8) g = (int64_t)sms * 8;
        qr_init_kernel<<<(unsigned)g, 256, 0, st>>>(vn, n, w.vn1, w.vn2, w.panel);
        if ((rc = check_launch("qr_init_kernel"))) return rc;
    }

    int64_t gv_grid = ceil_div((n + 1) / 2, GV_THREADS);
    if (gv_grid > (int64_t)sms * 8) gv_grid = (int64_t)sms * 8;
    if (gv_grid > QR_NCAND) gv_grid = QR_NCAND;

    // step-0 argmax: a read-only pass over zero rows leaves the norms untouched
    qr_gemv_kernel<<<(unsigned)gv_grid, GV_THREADS, 0, st>>>(d_Ut, ld, n, 0, 0, 0, w.panel, w.vn1, w.vn2, s, w.cand);
    if ((rc = check_launch("qr_gemv_kernel"))) return rc;
    int ncand = (int)gv_grid;

    const double* src = d_Ut;   // trailing matrix as of the block start, rows i0..r-1
    int i0 = 0;
    static bool attr_set = false;
    if (!attr_set) {
        OMB_CUDA(cudaFuncSetAttribute(qr_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    for (int i = 0; i < (int)s; ++i) {
        const int t = i - i0;
        const int L = (int)r - i0;
        qr_panel_kernel<<<1, PN_THREADS, 0, st>>>(w.panel, w.cand, ncand, src + (int64_t)i0 * ld, ld, L, i, t,
                                                   block == 1 ? 1 : 0, s, index_base, w.vn1, d_piv, d_rdiag, d_gap);
        if ((rc = check_launch("qr_panel_kernel"))) return rc;
        if (i == (int)s - 1) break;           // no further pivot needed: skip the last pass
        const bool close_block = (t == block - 1);
        if (close_block) {
            const size_t smem = sizeof(double) * ((size_t)(t + 1) * L + (size_t)L * AP_THREADS);
            OMB_CHECK_ARG(smem <= 220 * 1024, "trailing block too tall for the apply kernel");
            int per_sm = (int)((220 * 1024) / (smem + 1024));
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 8) per_sm = 8;
            int64_t g = ceil_div(n, AP_THREADS);
            if (g > (int64_t)sms * per_sm) g = (int64_t)sms * per_sm;
            if (g > QR_NCAND) g = QR_NCAND;
            qr_apply_kernel<<<(unsigned)g, AP_THREADS, smem, st>>>(src + (int64_t)i0 * ld, d_work + (int64_t)i0 * ld,
                                                                   ld, n, L, t, w.panel, w.vn1, w.vn2, s, w.cand);
            if ((rc = check_launch("qr_apply_kernel"))) return rc;
            ncand = (int)g;
            src = d_work;
            i0 = i + 1;
        } else {
            qr_gemv_kernel<<<(unsigned)gv_grid, GV_THREADS, 0, st>>>(src + (int64_t)i0 * ld, ld, n, L, t,
                                                                    (t + 1 == L) ? 1 : 0, w.panel, w.vn1, w.vn2, s,
                                                                    w.cand);
            if ((rc = check_launch("qr_gemv_kernel"))) return rc;
            ncand = (int)gv_grid;
        }
    }
    return 0;
}
