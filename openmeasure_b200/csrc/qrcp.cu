// qrcp.cu -- K6: QR with column pivoting over the n candidate sensor locations.
//
// Replaces scipy.linalg.qr(self.Ur.T, pivoting=True, mode='economic') (reference
// sparse_sensing.py:739), i.e. LAPACK dgeqp3 -> dlaqp2 (+ dlarfg / dlarf) on the r x n matrix
// A = Ur^T, of which the reference keeps only the first r pivots (:741-743).
//
// Layout: A is tiled mode-major, A[tile][q][128] (common.cuh): the trailing rows [i0, r) of the
// 128 candidates of a tile are one contiguous burst.  One thread (or one DMMA fragment slot) owns
// one candidate column per step; the only cross-thread work is the (norm, LAPACK-position) argmax.
// Per pivot step:
//
//   panel kernel (1 CTA)   reduce the per-CTA argmax records -> pivot p; LAPACK's column-swap
//                          bookkeeping (position keys for tie-breaking); gather p's trailing
//                          column; dlarfg -> (v, tau, beta); compact-WY T and q = Q e_t.
//   pass kernel (grid)     inside a block: read-only GEMV  R[i, j] = q . A[i0:, j]  + dlaqp2's
//                          norm down-date (tol3z guard, exact recomputation) + next argmax: the
//                          trailing matrix is streamed once and never written.
//                          last step of a block: apply the block's reflectors to every column,
//                          write the rows below the block, norms (block > 1: exact), argmax.
//                            block == 1: dlarf arithmetic in registers, the very fma sequence of
//                                        oracle/csrc/oracle.c -> bit-identical to dlaqp2;
//                            block  > 1: compact WY on the FP64 tensor path (DMMA.8x8x4); with 41 .. 104
//                                        trailing rows the tiles arrive as TMA tensor copies
//                                        (qr_apply_tma_kernel), otherwise through registers.
//
// block > 1 moves ~ (1 + 1/block)/2 of the bytes of the unblocked algorithm; pivots are identical
// on non-degenerate inputs (the degeneracy meter d_gap reports how close any decision was).
//
// Lazy norm down-dates (block > 1, omb_qrcp_set_lazy): partial column norms only ever shrink, so a
// column whose norm at the block start is below the norm of the pivot that is finally chosen cannot
// be that pivot -- its row of R and its down-date are not needed until the block closes.  Candidates
// are grouped in SEGMENTS of 64 (half a tile, one warp iteration of the read-only pass).  At a block
// start every norm is exact and seg_max[seg] holds the segment's largest; the panel sets the bound
// theta = alpha * (pivot norm) and the in-block passes skip every segment with seg_max < theta.  A
// pivot found among the remaining segments is THE pivot iff its norm is >= theta (everything skipped
// is strictly below theta).  Otherwise (norm c < theta) the panel raises `retry`: a catch-up pass
// brings the skipped segments with seg_max >= c up to date (all deferred steps of the block, same
// q . a arithmetic as the regular pass), they stay active for the rest of the block, theta drops to
// c and the panel is repeated on both record sets -- exact again, since what is still skipped is
// below c.  The block-closing apply pass reads every column anyway and leaves EVERY column's exact
// trailing norm (sum of squares of the rows it is about to write; dlaqp2's recompute branch taken
// unconditionally, lazy or not): a norm at a block boundary is a function of the column alone, the
// in-block down-dates of the visited segments served the block's pivot decisions only.  Lazy and
// eager runs therefore hold bit-identical norms at every decision, exact ties included.
// The decisions depend on the data only (identical on every rank of a row-sharded run).
#include "common.cuh"
#include <cuda.h>      // CUtensorMap (types only: the encoder comes through cudaGetDriverEntryPoint)
#include <stdlib.h>
#include "../../include/omb200.h"

namespace omb {

constexpr int QR_RMAX = 256;     // max modes (rows of A)
constexpr int QR_BMAX = 8;       // max steps per block
constexpr int QR_NCAND = 4096;   // max CTAs of a pass kernel (argmax records)
constexpr double QR_TOL3Z = 1.0536712127723509e-08;   // sqrt(2^-53), LAPACK tol3z
constexpr int QR_LREG = 100;     // tallest trailing block of the register-resident (bit-exact) pass

// Programmatic dependent launch: the kernels of the placement loop form a strict chain (pass -> panel
// -> pass ...).  Each one lets its successor start launching at once (launch_dependents) and waits for
// its predecessor's results at its own top (wait), so launch latency and CTA scheduling of kernel k+1
// overlap the execution of kernel k instead of following its drain.
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
    double qs[QR_BMAX][QR_RMAX];  // q_t = Q e_t of every in-block step  ->  R[i0 + t, j] = q_t . A[i0:, j]
    int64_t posmap[QR_RMAX];      // current LAPACK position of original column c < s
    int64_t col_at_pos[QR_RMAX];  // original column now at position k < s
    int ncand;                    // argmax records left by the last pass kernel
    int ncand2;                   // ... by the last catch-up pass (records at cand + QR_NCAND / 2)
    // lazy down-dates (see the file header)
    int lazy;                     // 0: every pass visits every segment
    int retry;                    // the last panel could not certify its pivot: catch-up pass + second panel run
    int xseq;                     // cross-rank exchanges done so far (slot parity and tag of the next one)
    int nretry;                   // statistics: catch-up rounds of this placement
    double alpha;                 // theta = alpha * (pivot norm at the block start)
    double theta;                 // every skipped segment's seg_max is < theta
    double cstar;                 // exact active maximum that failed the test (the catch-up threshold)
    long long rest_bits;          // largest block-start norm among the segments still skipped (bit pattern of a
                                  // double >= 0 or of -1: ordered like a signed integer), for the gap meter
    unsigned long long seg_rows;  // statistics: (segment, row) visits of the read-only passes
    unsigned long long seg_visits;// statistics: segment visits of the read-only passes
};

constexpr int QR_REC = 8 + QR_RMAX;   // doubles per cross-rank record: header + pivot column tail

// Row sharding across ranks: every rank holds cells [cell0, cell0 + n_c_loc) of every feature, so
// local row j = f * n_c_loc + c is global row f * n_c + cell0 + c.  world == 1: identity.
struct Shard {
    int64_t n_c_loc, n_c, cell0;
    int rank, world;
};
__device__ __forceinline__ int64_t glob_index(int64_t j, const Shard& sh)
{
    if (sh.world == 1) return j;
    const int64_t f = j / sh.n_c_loc;
    return f * sh.n_c + sh.cell0 + (j - f * sh.n_c_loc);
}

__device__ __forceinline__ bool cand_better(double b1, int64_t k1, double b2, int64_t k2)
{
    return b1 > b2 || (b1 == b2 && k1 < k2);
}
// compare-select max: the norms are never NaN (fmax costs ~10 instructions in FP64 for its NaN rules)
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ void cand_merge(Cand& a, const Cand& b)
{
    if (cand_better(b.best, b.key, a.best, a.key)) {
        double s = dmax(a.best, b.second);
        a.best = b.best; a.idx = b.idx; a.key = b.key; a.second = s;
    } else {
        a.second = dmax(a.second, b.best);
    }
}
__device__ __forceinline__ void cand_push(Cand& a, double v, int64_t idx, int64_t key)
{
    if (cand_better(v, key, a.best, a.key)) { a.second = a.best; a.best = v; a.idx = idx; a.key = key; }
    else a.second = dmax(a.second, v);
}
// push with the LAPACK position key computed only when it can matter (v >= current best)
__device__ __forceinline__ void cand_push_lazy(Cand& a, double v, int64_t j, const Shard& sh, const Panel* P,
                                               int64_t s_total)
{
    if (v >= a.best) {
        const int64_t g = glob_index(j, sh);
        const int64_t key = g < s_total ? P->posmap[g] : g;
        cand_push(a, v, j, key);
    } else {
        a.second = dmax(a.second, v);
    }
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
__device__ __forceinline__ Cand cand_empty()
{
    Cand c;
    c.best = -1.0; c.second = -1.0; c.idx = -1; c.key = INT64_MAX;
    return c;
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
qr_norms_kernel(const double* __restrict__ A, int64_t n, int r, double* __restrict__ vn)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const double* col = A + basis_index(0, j, r);
        double s = 0.0;
        for (int k = 0; k < r; ++k) { double x = ldg_stream(col + (int64_t)k * OMB_TB); s = fma(x, x, s); }
        vn[j] = sqrt(s);
    }
}

__global__ void __launch_bounds__(256)
qr_init_kernel(const double* __restrict__ vn, int64_t n, double* __restrict__ vn1, double* __restrict__ vn2,
               Panel* __restrict__ P, double lazy_alpha)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double v = vn[j];
        vn1[j] = v;
        vn2[j] = v;
    }
    if (blockIdx.x == 0) {
        for (int k = threadIdx.x; k < QR_RMAX; k += blockDim.x) { P->posmap[k] = k; P->col_at_pos[k] = k; }
        if (threadIdx.x == 0) {
            P->ncand = 0; P->ncand2 = 0;
            P->lazy = lazy_alpha > 0.0 ? 1 : 0;
            P->retry = 0; P->xseq = 0; P->nretry = 0;
            P->alpha = lazy_alpha; P->theta = -1.0; P->cstar = -1.0; P->rest_bits = __double_as_longlong(-1.0);
            P->seg_rows = 0ull; P->seg_visits = 0ull;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// slow path of the read-only pass: exact trailing norm of one column after reflectors 0..t
// (col points at the column's first trailing row; rows are OMB_TB apart)
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ double recompute_norm(const double* __restrict__ col, int L, int t,
                                              const double* __restrict__ Vg, const double* __restrict__ taug)
{
    double c[QR_RMAX];
    for (int k = 0; k < L; ++k) c[k] = col[(int64_t)k * OMB_TB];
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
// read-only pass: R[i, j] = q . A[i0:, j], down-date, argmax.  L = r - i0 rows.
// A warp iteration sweeps one SEGMENT = 64 adjacent columns (half a tile): two columns per thread
// (128-bit loads), the trailing rows of the segment are L runs of 512 bytes, 1 KB apart.
//   GV_START    argmax only (no rows read); leaves every segment's largest norm in seg_w
//   GV_PASS     in-block step t.  Lazy: the first pass of a block (t == 0) decides which segments the
//               block skips (seg_r[seg] < theta), clears seg_w for the block-closing apply pass; the
//               later passes follow the flags
//   GV_CATCHUP  only after a panel raised `retry` (exits at once otherwise): the skipped segments
//               with seg_r[seg] >= cstar get the down-dates of steps 0 .. t-1 and become active
// Each lane first looks at the flags of 32 upcoming segments of its warp (one load instead of a
// dependent load per iteration), the warp then walks the set bits.
// ---------------------------------------------------------------------------------------------
constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int QR_SEG = 64;        // candidates per segment
enum { GV_START = 0, GV_PASS = 1, GV_CATCHUP = 2 };

template <int MODE>
__global__ void __launch_bounds__(GV_THREADS, MODE == GV_CATCHUP ? 2 : 3)
qr_gemv_kernel(const double* __restrict__ src, int64_t n, int r, int i0, int L, int t,
               Panel* __restrict__ P, double* __restrict__ vn1, double* __restrict__ vn2,
               int64_t s_total, Shard sh, Cand* __restrict__ cand, const double* __restrict__ seg_r,
               double* __restrict__ seg_w, unsigned char* __restrict__ seg_skip)
{
    // steps whose row of R this launch forms: the pass does step t, the catch-up steps 0 .. t-1
    constexpr int NQ = (MODE == GV_CATCHUP) ? QR_BMAX - 1 : 1;
    __shared__ double s_q[NQ][QR_RMAX];
    __shared__ Cand s_c[GV_WARPS];
    __shared__ unsigned long long s_vis;
    __shared__ long long s_rest;
    pdl_enter();
    const bool lazy = (MODE != GV_START) && P->lazy != 0;
    if (MODE == GV_CATCHUP && (!lazy || P->retry == 0)) return;
    const int t_lo = (MODE == GV_CATCHUP) ? 0 : t;
    const int nq = (MODE == GV_CATCHUP) ? t : 1;
    if (MODE != GV_START)
        for (int a = 0; a < nq; ++a)
            for (int k = threadIdx.x; k < L; k += GV_THREADS) s_q[a][k] = P->qs[t_lo + a][k];
    if (threadIdx.x == 0) { s_vis = 0ull; s_rest = __double_as_longlong(-1.0); }
    const double theta = lazy ? P->theta : -1.0;
    const double cstar = lazy ? P->cstar : -1.0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Cand best = cand_empty();
    const int64_t nseg = basis_tiles(n) * (OMB_TB / QR_SEG);
    const int64_t wstride = (int64_t)gridDim.x * GV_WARPS;
    unsigned visits = 0;
    double rest = -1.0;            // largest block-start norm among the segments this lane leaves skipped
    for (int64_t base = (int64_t)blockIdx.x * GV_WARPS + warp; base < nseg; base += 32 * wstride) {
        const int64_t myseg = base + (int64_t)lane * wstride;
        bool act = myseg < nseg;
        if (lazy && act) {
            if (MODE == GV_CATCHUP) {
                act = false;
                if (seg_skip[myseg] != 0) {
                    const double sm = seg_r[myseg];
                    act = sm >= cstar;
                    if (act) seg_skip[myseg] = 0;
                    else rest = dmax(rest, sm);
                }
            } else if (t == 0) {
                const double sm = seg_r[myseg];
                const bool skip = sm < theta;
                seg_skip[myseg] = skip ? 1 : 0;
                seg_w[myseg] = -1.0;
                act = !skip;
                if (skip) rest = dmax(rest, sm);
            } else {
                act = seg_skip[myseg] == 0;
                if (!act) rest = dmax(rest, seg_r[myseg]);
            }
        }
        unsigned todo = __ballot_sync(0xFFFFFFFFu, act);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t seg = base + (int64_t)l * wstride;
            const int64_t j = seg * QR_SEG + 2 * lane;
            ++visits;
            if (MODE == GV_START) {
                double m0 = -1.0;
                if (j < n) {
                    const double2 pv1 = *reinterpret_cast<const double2*>(vn1 + j);
                    if (pv1.x >= 0.0) { m0 = pv1.x; cand_push_lazy(best, pv1.x, j, sh, P, s_total); }
                    if (j + 1 < n && pv1.y >= 0.0) { m0 = dmax(m0, pv1.y); cand_push_lazy(best, pv1.y, j + 1, sh, P, s_total); }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m0 = dmax(m0, __shfl_xor_sync(0xFFFFFFFFu, m0, o));
                if (lane == 0) seg_w[seg] = m0;
                continue;
            }
            if (j >= n) continue;
            const double* col = src + basis_index(i0, j, r);
            // n is padded to whole tiles in the norm arrays, so the pair load is always in bounds
            const double2 pv1 = *reinterpret_cast<const double2*>(vn1 + j);
            const double2 pv2 = *reinterpret_cast<const double2*>(vn2 + j);
            if (MODE == GV_PASS) {
                const bool last_row = (t + 1 == L);
                double y0 = 0.0, y1 = 0.0;
                int k = 0;
                for (; k + 8 <= L; k += 8) {
                    double2 a[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = ldg_stream2(col + (int64_t)(k + u) * OMB_TB);
#pragma unroll
                    for (int u = 0; u < 8; ++u) { y0 = fma(s_q[0][k + u], a[u].x, y0); y1 = fma(s_q[0][k + u], a[u].y, y1); }
                }
                for (; k < L; ++k) {
                    double2 a = ldg_stream2(col + (int64_t)k * OMB_TB);
                    y0 = fma(s_q[0][k], a.x, y0);
                    y1 = fma(s_q[0][k], a.y, y1);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t jj = j + e;
                    if (jj >= n) break;
                    double v1 = e ? pv1.y : pv1.x;
                    if (v1 < 0.0) continue;                 // already a pivot
                    if (v1 != 0.0) {
                        const double v2 = e ? pv2.y : pv2.x;
                        if (downdate(e ? y1 : y0, v1, v2)) {
                            v1 = last_row ? 0.0 : recompute_norm(col + e, L, t, &P->V[0][0], P->tau);
                            vn2[jj] = v1;
                        }
                        vn1[jj] = v1;
                    }
                    cand_push_lazy(best, v1, jj, sh, P, s_total);
                }
            } else {
                // catch-up: one sweep over the segment's rows forms the rows of R of ALL deferred steps (each
                // accumulator runs over k in the order of the regular pass: same bits), then dlaqp2's
                // down-dates in step order
                double v1[2] = {pv1.x, pv1.y}, v2[2] = {pv2.x, pv2.y};
                bool w2[2] = {false, false};
                double y0[NQ], y1[NQ];
#pragma unroll
                for (int a = 0; a < NQ; ++a) y0[a] = y1[a] = 0.0;
                int k = 0;
                for (; k + 4 <= L; k += 4) {
                    double2 x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) x[u] = ldg_stream2(col + (int64_t)(k + u) * OMB_TB);
#pragma unroll
                    for (int a = 0; a < NQ; ++a)
                        if (a < nq) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                y0[a] = fma(s_q[a][k + u], x[u].x, y0[a]);
                                y1[a] = fma(s_q[a][k + u], x[u].y, y1[a]);
                            }
                        }
                }
                for (; k < L; ++k) {
                    const double2 x = ldg_stream2(col + (int64_t)k * OMB_TB);
#pragma unroll
                    for (int a = 0; a < NQ; ++a)
                        if (a < nq) { y0[a] = fma(s_q[a][k], x.x, y0[a]); y1[a] = fma(s_q[a][k], x.y, y1[a]); }
                }
#pragma unroll
                for (int a = 0; a < NQ; ++a) {
                    if (a >= nq) break;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (j + e >= n || v1[e] <= 0.0) continue;     // past the end, a pivot, or an exact zero
                        if (downdate(e ? y1[a] : y0[a], v1[e], v2[e])) {
                            v1[e] = (a + 1 == L) ? 0.0 : recompute_norm(col + e, L, a, &P->V[0][0], P->tau);
                            v2[e] = v1[e];
                            w2[e] = true;
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t jj = j + e;
                    if (jj >= n || v1[e] < 0.0) continue;
                    if (w2[e]) vn2[jj] = v2[e];
                    vn1[jj] = v1[e];
                    cand_push_lazy(best, v1[e], jj, sh, P, s_total);
                }
            }
        }
    }
    best = cand_block_reduce(best, s_c);          // (contains a CTA barrier)
    if (MODE != GV_START) {
        if (lazy) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rest = dmax(rest, __shfl_xor_sync(0xFFFFFFFFu, rest, o));
            if (lane == 0 && rest >= 0.0) atomicMax(&s_rest, __double_as_longlong(rest));
        }
        if (lane == 0 && visits) atomicAdd(&s_vis, (unsigned long long)visits);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (MODE == GV_CATCHUP) {
            cand[QR_NCAND / 2 + blockIdx.x] = best;
            if (blockIdx.x == 0) P->ncand2 = (int)gridDim.x;
        } else {
            cand[blockIdx.x] = best;
            if (blockIdx.x == 0) P->ncand = (int)gridDim.x;
        }
        if (MODE != GV_START) {
            if (s_vis) { atomicAdd(&P->seg_rows, s_vis * (unsigned long long)(L * nq)); atomicAdd(&P->seg_visits, s_vis); }
            if (lazy && s_rest >= 0) atomicMax(&P->rest_bits, s_rest);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// unblocked apply pass (block == 1), register-resident: one thread owns one column tail of up to
// LMAX rows and applies dlarf:  w = v^T c ; c -= tau w v  -- the fma sequence of the oracle, so
// the trailing matrix, the norms and the pivots are bit-identical to dlaqp2.  v lives in shared
// memory (one broadcast LDS per fma; this pass is HBM-bound at a few fma per 16 bytes).
// ---------------------------------------------------------------------------------------------
constexpr int AR_THREADS = 128;   // == OMB_TB: one CTA iteration sweeps one tile
constexpr int ar_min_blocks(int lmax) { return lmax <= 16 ? 6 : (lmax <= 32 ? 4 : (lmax <= 48 ? 3 : 2)); }

template <int LMAX>
__global__ void __launch_bounds__(AR_THREADS, ar_min_blocks(LMAX))
qr_apply1_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n, int r, int i0, int L,
                 const Panel* __restrict__ P, double* __restrict__ vn1, double* __restrict__ vn2,
                 int64_t s_total, Shard sh, Cand* __restrict__ cand)
{
    __shared__ double s_v[LMAX];
    __shared__ double s_tau;
    __shared__ Cand s_c[AR_THREADS / 32];
    pdl_enter();
    for (int k = threadIdx.x; k < LMAX; k += AR_THREADS) s_v[k] = (k < L) ? P->V[0][k] : 0.0;
    if (threadIdx.x == 0) s_tau = P->tau[0];
    __syncthreads();
    const double tau = s_tau;
    const bool last_row = (L == 1);

    Cand best = cand_empty();
    const int64_t ntiles = basis_tiles(n);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t j = tile * OMB_TB + threadIdx.x;
        if (j >= n) continue;
        const double* col = src + tile * ((int64_t)r * OMB_TB) + (int64_t)i0 * OMB_TB + threadIdx.x;
        double* outp = dst + tile * ((int64_t)r * OMB_TB) + (int64_t)i0 * OMB_TB + threadIdx.x;
        const double pv1 = vn1[j], pv2 = vn2[j];
        double c[LMAX];
#pragma unroll
        for (int k = 0; k < LMAX; ++k) c[k] = (k < L) ? ldg_stream(col + k * OMB_TB) : 0.0;
        // w = v^T c (v[0] = 1; rows k >= L hold exact zeros)
        double w = c[0];
#pragma unroll
        for (int k = 1; k < LMAX; ++k) w = fma(s_v[k], c[k], w);
        const double tw = tau * w;
        const double rij = c[0] - tw;
#pragma unroll
        for (int k = 1; k < LMAX; ++k) {
            c[k] = fma(-tw, s_v[k], c[k]);
            if (k < L) stg_stream(outp + k * OMB_TB, c[k]);
        }
        double v1 = pv1;
        if (v1 >= 0.0) {
            if (v1 != 0.0) {
                const double v2 = pv2;
                if (downdate(rij, v1, v2)) {
                    double sq = 0.0;
#pragma unroll
                    for (int k = 1; k < LMAX; ++k) sq = fma(c[k], c[k], sq);
                    v1 = last_row ? 0.0 : sqrt(sq);
                    vn2[j] = v1;
                }
                vn1[j] = v1;
            }
            cand_push_lazy(best, v1, j, sh, P, s_total);
        }
    }
    best = cand_block_reduce(best, s_c);
    if (threadIdx.x == 0) {
        cand[blockIdx.x] = best;
        if (blockIdx.x == 0) const_cast<Panel*>(P)->ncand = (int)gridDim.x;
    }
}

typedef void (*Apply1Fn)(const double*, double*, int64_t, int, int, int, const Panel*, double*, double*, int64_t,
                         Shard, Cand*);
static Apply1Fn pick_apply1(int L, int* lmax)
{
    if (L <= 16) { *lmax = 16; return qr_apply1_kernel<16>; }
    if (L <= 32) { *lmax = 32; return qr_apply1_kernel<32>; }
    if (L <= 48) { *lmax = 48; return qr_apply1_kernel<48>; }
    if (L <= 64) { *lmax = 64; return qr_apply1_kernel<64>; }
    if (L <= 80) { *lmax = 80; return qr_apply1_kernel<80>; }
    if (L <= QR_LREG) { *lmax = QR_LREG; return qr_apply1_kernel<QR_LREG>; }
    return nullptr;
}

// ---------------------------------------------------------------------------------------------
// blocked apply pass on the FP64 tensor path (up to 8 reflectors per block):
//     C^T <- C^T - ((C^T V) T) V^T        for groups of 8 candidate columns,
// i.e. three small GEMMs per group on DMMA.8x8x4.  A warp owns 8*NG columns; each lane keeps
// C[8G + 2p + e][col] (p = lane%4, e = 0,1, col = lane/4 of its group) -- which is at once the
// A fragment of the first product (rows taken in the order {0,2,4,6},{1,3,5,7} of each group of
// eight) and the accumulator fragment of the last one, so nothing is shuffled or staged: every
// element is read once from HBM and written once.  The reflector fragments come from shared
// memory once per 8-row group and are reused by the NG column groups (one LDS per 4*NG*256 FMAs;
// a DFMA formulation needs one operand fetch per FMA and is issue-bound).
// ---------------------------------------------------------------------------------------------
constexpr int AM_THREADS = 128;
constexpr int AM_SV = 10;     // row stride of sV  [row][refl]   : (2p*10 + c) distinct mod 16
constexpr int AM_ST = 10;     // row stride of sT  [refl][refl']

template <int LG, int NG>
__global__ void __launch_bounds__(AM_THREADS)
qr_apply_mma_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n, int r, int i0, int L, int t,
                    const Panel* __restrict__ P, double* __restrict__ vn1, double* __restrict__ vn2,
                    int64_t s_total, Shard sh, Cand* __restrict__ cand, double* __restrict__ seg_w)
{
    constexpr int LP = LG * 8;                 // padded rows
    constexpr int SVT = LP + 2;                // row stride of sVt [refl][row]: == 2 (mod 8)
    constexpr int WT = 8 * NG;                 // columns per warp tile (divides OMB_TB)
    __shared__ double sV[LP * AM_SV];
    __shared__ double sVt[8 * SVT];
    __shared__ double sT[8 * AM_ST];
    __shared__ Cand s_c[AM_THREADS / 32];
    pdl_enter();
    for (int e = threadIdx.x; e < LP * 8; e += AM_THREADS) {
        const int k = e >> 3, a = e & 7;
        const double v = (k < L && a <= t) ? P->V[a][k] : 0.0;
        sV[k * AM_SV + a] = v;
        sVt[a * SVT + k] = v;
    }
    if (threadIdx.x < 64) {
        const int a = threadIdx.x >> 3, b = threadIdx.x & 7;
        sT[a * AM_ST + b] = (a <= b && b <= t) ? P->T[a][b] : 0.0;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = lane & 3, cq = lane >> 2;
    const int tp = (t & 7) >> 1;                  // t < 8: the rows of R live in row group 0
    const bool last_row = (t + 1 == L);
    const bool lazy = P->lazy != 0;

    Cand best = cand_empty();
    const int64_t nwt = basis_tiles(n) * (OMB_TB / WT);
    for (int64_t wt = (int64_t)blockIdx.x * (AM_THREADS / 32) + warp; wt < nwt;
         wt += (int64_t)gridDim.x * (AM_THREADS / 32)) {
        const int64_t j0 = wt * WT;
        if (j0 >= n) continue;
        const int64_t tbase = (j0 >> 7) * ((int64_t)r * OMB_TB) + (int64_t)i0 * OMB_TB + (j0 & (OMB_TB - 1));
        const double* in = src + tbase + (2 * p) * OMB_TB + cq;
        double* out = dst + tbase + (2 * p) * OMB_TB + cq;
        // norms of the columns this lane owns (p == tp), fetched up front so that the down-date at
        // the end of the tile does not add a dependent round trip to HBM
        double pv1[NG], pv2[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int64_t j = j0 + 8 * g + cq;
            const bool owner = (p == tp) && (j < n);
            pv1[g] = owner ? vn1[j] : -1.0;
            pv2[g] = owner ? vn2[j] : 1.0;
        }
        double c[LG][NG][2];
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = 8 * G + 2 * p + e;
#pragma unroll
                for (int g = 0; g < NG; ++g)
                    c[G][g][e] = (row < L) ? ldg_stream(in + (8 * G + e) * OMB_TB + 8 * g) : 0.0;
            }
        // Z^T = C^T V
        double z[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) z[g][0] = z[g][1] = 0.0;
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double bv = sV[(8 * G + 2 * p + e) * AM_SV + cq];
#pragma unroll
                for (int g = 0; g < NG; ++g) dmma884(z[g][0], z[g][1], c[G][g][e], bv);
            }
        // Z'^T = Z^T T
        double zp[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) zp[g][0] = zp[g][1] = 0.0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double bv = sT[(2 * p + e) * AM_ST + cq];
#pragma unroll
            for (int g = 0; g < NG; ++g) dmma884(zp[g][0], zp[g][1], z[g][e], bv);
        }
        // C^T -= Z'^T V^T
#pragma unroll
        for (int g = 0; g < NG; ++g) { zp[g][0] = -zp[g][0]; zp[g][1] = -zp[g][1]; }
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double bv = sVt[(2 * p + e) * SVT + 8 * G + cq];
#pragma unroll
                for (int g = 0; g < NG; ++g) dmma884(c[G][g][0], c[G][g][1], zp[g][e], bv);
            }
        // rows below the block go back to HBM; R[i, j] = row t
        double rij[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) rij[g] = 0.0;
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = 8 * G + 2 * p + e;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (G == 0 && 2 * p + e == t) rij[g] = c[G][g][e];          // t < 8: row group 0
                    if (row > t && row < L) stg_stream(out + (8 * G + e) * OMB_TB + 8 * g, c[G][g][e]);
                }
            }
        // dlaqp2's down-date of row t: the lanes holding row t (p == tp) own their column's norms
        double wmax = -1.0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int64_t j = j0 + 8 * g + cq;
            const bool owner = (p == tp) && (j < n);
            double v1 = pv1[g];
            bool redo = false;
            if (owner && v1 > 0.0) redo = downdate(rij[g], v1, pv2[g]);
            if (__any_sync(0xFFFFFFFFu, redo)) {
                // exact trailing norm: each of the column's 4 lanes sums its rows, then combine
                double sq = 0.0;
#pragma unroll
                for (int G = 0; G < LG; ++G)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int row = 8 * G + 2 * p + e;
                        if (row > t && row < L) sq = fma(c[G][g][e], c[G][g][e], sq);
                    }
                sq += __shfl_xor_sync(0xFFFFFFFFu, sq, 1);
                sq += __shfl_xor_sync(0xFFFFFFFFu, sq, 2);
                if (redo) {
                    v1 = last_row ? 0.0 : sqrt(sq);
                    vn2[j] = v1;
                }
            }
            if (owner && v1 >= 0.0) {
                vn1[j] = v1;
                wmax = dmax(wmax, v1);
                cand_push_lazy(best, v1, j, sh, P, s_total);
            }
        }
        if (lazy) {
            // the segment's largest norm at the start of the next block (norms are >= 0 or -1: their
            // bit patterns order like signed integers)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wmax = dmax(wmax, __shfl_xor_sync(0xFFFFFFFFu, wmax, o));
            if (lane == 0)
                atomicMax(reinterpret_cast<long long*>(seg_w + j0 / QR_SEG), __double_as_longlong(wmax));
        }
    }
    best = cand_block_reduce(best, s_c);
    if (threadIdx.x == 0) {
        cand[blockIdx.x] = best;
        if (blockIdx.x == 0) const_cast<Panel*>(P)->ncand = (int)gridDim.x;
    }
}

// ---------------------------------------------------------------------------------------------
// block-closing apply pass of a blocked schedule (block > 1): same compact-WY update, and EVERY
// column leaves with its EXACT trailing norm (sum of squares of the rows written back; dlaqp2's
// recompute branch taken unconditionally, vn2 following as there).  A norm at a block boundary is
// then a function of the column alone, whether or not the block's read-only passes visited it
// (lazy down-dates): duplicated candidates stay bit-identical and ties fall as in LAPACK.
// The pass is bound by the length of a warp's per-tile chain (loads -> 3 dependent products ->
// stores -> norms) at 8-12 resident warps per SM, not by the DMMA pipe or by HBM as such (measured:
// loads alone 6.3 TB/s, stores alone 4.3, the products alone 1.5 ms of a 4.9 ms pass, the sum of the
// parts = the pass), so the chain is kept short:
//   * the first product accumulates even and odd row groups separately (half the dependent DMMAs),
//   * lane p of a column's quad owns column group g = p: the square roots, norm stores and the
//     candidate bookkeeping of a tile run on all lanes at once instead of NG times on a quarter,
//   * the NEXT tile's loads are issued right after this tile's stores; the norms are finished
//     while they fly.
// ---------------------------------------------------------------------------------------------
template <int LG, int NG>
__global__ void __launch_bounds__(AM_THREADS)
qr_apply_exact_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n, int r, int i0, int L, int t,
                      const Panel* __restrict__ P, double* __restrict__ vn1, double* __restrict__ vn2,
                      int64_t s_total, Shard sh, Cand* __restrict__ cand, double* __restrict__ seg_w)
{
    constexpr int LP = LG * 8;                 // padded rows
    constexpr int SVT = LP + 2;                // row stride of sVt [refl][row]: == 2 (mod 8)
    constexpr int WT = 8 * NG;                 // columns per warp tile (divides OMB_TB)
    __shared__ double sV[LP * AM_SV];
    __shared__ double sVt[8 * SVT];
    __shared__ double sT[8 * AM_ST];
    __shared__ Cand s_c[AM_THREADS / 32];
    pdl_enter();
    for (int e = threadIdx.x; e < LP * 8; e += AM_THREADS) {
        const int k = e >> 3, a = e & 7;
        const double v = (k < L && a <= t) ? P->V[a][k] : 0.0;
        sV[k * AM_SV + a] = v;
        sVt[a * SVT + k] = v;
    }
    if (threadIdx.x < 64) {
        const int a = threadIdx.x >> 3, b = threadIdx.x & 7;
        sT[a * AM_ST + b] = (a <= b && b <= t) ? P->T[a][b] : 0.0;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = lane & 3, cq = lane >> 2;
    const bool last_row = (t + 1 == L);
    const bool lazy = P->lazy != 0;

    Cand best = cand_empty();
    const int64_t nwt = basis_tiles(n) * (OMB_TB / WT);
    const int64_t wstride = (int64_t)gridDim.x * (AM_THREADS / 32);
    int64_t wt = (int64_t)blockIdx.x * (AM_THREADS / 32) + warp;

    double c[LG][NG][2];
    double pv = -1.0;                           // norm of this lane's column (group g = p) of the tile in c
    // all loads of a warp tile, issued back to back (and the lane's norm, so that the end of the tile
    // does not add a dependent round trip)
    auto load_tile = [&](int64_t w) {
        const int64_t j0 = w * WT;
        const int64_t tbase = (j0 >> 7) * ((int64_t)r * OMB_TB) + (int64_t)i0 * OMB_TB + (j0 & (OMB_TB - 1));
        const double* in = src + tbase + (2 * p) * OMB_TB + cq;
        const int64_t j = j0 + 8 * p + cq;
        pv = (p < NG && j < n) ? vn1[j] : -1.0;
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = 8 * G + 2 * p + e;
#pragma unroll
                for (int g = 0; g < NG; ++g)
                    c[G][g][e] = (row < L) ? ldg_stream(in + (8 * G + e) * OMB_TB + 8 * g) : 0.0;
            }
    };
    bool have = wt < nwt && wt * WT < n;
    if (have) load_tile(wt);
    while (have) {
        const int64_t j0 = wt * WT;
        // Z^T = C^T V, even and odd row groups on separate accumulators
        double z[NG][2], zo[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) z[g][0] = z[g][1] = zo[g][0] = zo[g][1] = 0.0;
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double bv = sV[(8 * G + 2 * p + e) * AM_SV + cq];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (G & 1) dmma884(zo[g][0], zo[g][1], c[G][g][e], bv);
                    else dmma884(z[g][0], z[g][1], c[G][g][e], bv);
                }
            }
#pragma unroll
        for (int g = 0; g < NG; ++g) { z[g][0] += zo[g][0]; z[g][1] += zo[g][1]; }
        // Z'^T = Z^T T
        double zp[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) zp[g][0] = zp[g][1] = 0.0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double bv = sT[(2 * p + e) * AM_ST + cq];
#pragma unroll
            for (int g = 0; g < NG; ++g) dmma884(zp[g][0], zp[g][1], z[g][e], bv);
        }
        // C^T -= Z'^T V^T
#pragma unroll
        for (int g = 0; g < NG; ++g) { zp[g][0] = -zp[g][0]; zp[g][1] = -zp[g][1]; }
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double bv = sVt[(2 * p + e) * SVT + 8 * G + cq];
#pragma unroll
                for (int g = 0; g < NG; ++g) dmma884(c[G][g][0], c[G][g][1], zp[g][e], bv);
            }
        // rows below the block go back to HBM; their squares give the exact trailing norms: each of a
        // column's 4 lanes sums its rows (two accumulators), the quad combines, lane p keeps column group p
        const int64_t tbase = (j0 >> 7) * ((int64_t)r * OMB_TB) + (int64_t)i0 * OMB_TB + (j0 & (OMB_TB - 1));
        double* out = dst + tbase + (2 * p) * OMB_TB + cq;
        double sq[NG], sq1[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) sq[g] = sq1[g] = 0.0;
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = 8 * G + 2 * p + e;
#pragma unroll
                for (int g = 0; g < NG; ++g)
                    if (row > t && row < L) {
                        stg_stream(out + (8 * G + e) * OMB_TB + 8 * g, c[G][g][e]);
                        if (e) sq1[g] = fma(c[G][g][e], c[G][g][e], sq1[g]);
                        else sq[g] = fma(c[G][g][e], c[G][g][e], sq[g]);
                    }
            }
        double mysq = 0.0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            double v = sq[g] + sq1[g];
            v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
            v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
            if (p == g) mysq = v;
        }
        const int64_t j = j0 + 8 * p + cq;
        const double pvc = pv;
        // the next tile's loads fly while this tile's norms are finished
        wt += wstride;
        have = wt < nwt && wt * WT < n;
        if (have) load_tile(wt);
        double v1 = pvc;                         // < 0: a pivot or no column (p >= NG, j >= n)
        if (v1 > 0.0) {
            v1 = last_row ? 0.0 : sqrt(mysq);
            vn2[j] = v1;
        }
        if (v1 >= 0.0) {
            vn1[j] = v1;
            cand_push_lazy(best, v1, j, sh, P, s_total);
        }
        if (lazy) {
            // the segment's largest norm at the start of the next block (norms are >= 0 or -1: their
            // bit patterns order like signed integers)
            double wmax = v1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wmax = dmax(wmax, __shfl_xor_sync(0xFFFFFFFFu, wmax, o));
            if (lane == 0)
                atomicMax(reinterpret_cast<long long*>(seg_w + j0 / QR_SEG), __double_as_longlong(wmax));
        }
    }
    best = cand_block_reduce(best, s_c);
    if (threadIdx.x == 0) {
        cand[blockIdx.x] = best;
        if (blockIdx.x == 0) const_cast<Panel*>(P)->ncand = (int)gridDim.x;
    }
}

// ---------------------------------------------------------------------------------------------
// The same block-closing pass with its loads on the TMA engine (6 <= ceil(L/8) <= 13: shorter blocks stay on the register-staged kernel, whose 32-candidate warp tiles amortise the per-tile work better).
// ncu on qr_apply_exact_kernel: 40 % of the warps' samples wait for the tile's loads, and the LSU data
// pipe moves ONE 32-byte sector per wavefront for its fragment-shaped accesses (4 row segments of 64
// bytes per instruction) -- 848 M wavefronts for 25 GB, the pipe 63 % busy.  Here a warp tile (16
// candidates x L rows = rows of 128 bytes, 1 KB apart in HBM) arrives as ONE tensor copy:
//   * every warp runs its own pipeline of TWO landing stages -- no CTA-wide synchronisation: lane 0
//     issues the tensor load of the warp's tile k + 2 (mbarrier completion) as soon as tile k has been
//     pulled into the registers, so two tiles per warp (26 KB) are in flight while one is worked on;
//   * the stages use the 128-byte swizzle of the tensor map (16-byte chunk index XOR row mod 8): a
//     fragment read -- 4 rows x 4 candidates per half-warp -- hits 8 distinct chunks, i.e. one
//     conflict-free wavefront per half-warp (unswizzled rows of 128 bytes would be a 4-way conflict);
//   * the results leave from the registers (streaming stores, fire and forget).
// 8 warps x 2 stages x 13 KB = 208 KB of shared memory at L = 100: one CTA per SM.
// ---------------------------------------------------------------------------------------------
constexpr int AT_WARPS = 8;
constexpr int AT_THREADS = AT_WARPS * 32;
constexpr int AT_WT = 16;                      // candidates per warp tile (two DMMA column groups)

// byte offset of (row, byte column) in a stage of 128-byte rows under the 128-byte swizzle
__device__ __forceinline__ int at_swz(int row, int colb) { return row * 128 + (colb ^ ((row & 7) << 4)); }

template <int LG>
__global__ void __launch_bounds__(AT_THREADS, 1)
qr_apply_tma_kernel(const __grid_constant__ CUtensorMap map_in, double* __restrict__ dst, int64_t n, int r,
                    int i0, int L, int t, const Panel* __restrict__ P, double* __restrict__ vn1, double* __restrict__ vn2,
                    int64_t s_total, Shard sh, Cand* __restrict__ cand, double* __restrict__ seg_w)
{
    constexpr int NG = 2;
    constexpr int LP = LG * 8;                 // padded rows
    constexpr int SVT = LP + 2;                // row stride of sVt [refl][row]: == 2 (mod 8)
    constexpr int STAGE = LP * 128;            // bytes per stage (a multiple of 1 KB: the swizzle atom)
    __shared__ double sV[LP * AM_SV];
    __shared__ double sVt[8 * SVT];
    __shared__ double sT[8 * AM_ST];
    __shared__ Cand s_c[AT_WARPS];
    __shared__ uint64_t s_bar[AT_WARPS][2];
    extern __shared__ __align__(1024) unsigned char s_stages[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* my_st = s_stages + (size_t)warp * (2 * STAGE);
    if (lane == 0) {
        mbar_init(&s_bar[warp][0], 1);
        mbar_init(&s_bar[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_enter();
    for (int e = threadIdx.x; e < LP * 8; e += AT_THREADS) {
        const int k = e >> 3, a = e & 7;
        const double v = (k < L && a <= t) ? P->V[a][k] : 0.0;
        sV[k * AM_SV + a] = v;
        sVt[a * SVT + k] = v;
    }
    if (threadIdx.x < 64) {
        const int a = threadIdx.x >> 3, b = threadIdx.x & 7;
        sT[a * AM_ST + b] = (a <= b && b <= t) ? P->T[a][b] : 0.0;
    }
    fence_proxy_async();                       // the barrier inits are visible to the TMA unit
    __syncthreads();

    const int p = lane & 3, cq = lane >> 2;
    const bool last_row = (t + 1 == L);
    const bool lazy = P->lazy != 0;
    const uint32_t tx_bytes = (uint32_t)L * 128u;

    Cand best = cand_empty();
    const int64_t nwt = basis_tiles(n) * (OMB_TB / AT_WT);
    const int64_t wstride = (int64_t)gridDim.x * AT_WARPS;
    int64_t wt = (int64_t)blockIdx.x * AT_WARPS + warp;
    auto valid = [&](int64_t w) { return w < nwt && w * AT_WT < n; };
    // tensor load of warp tile w into stage s (lane 0)
    auto issue_tile = [&](int64_t w, int s) {
        if (lane == 0) {
            mbar_expect_tx(&s_bar[warp][s], tx_bytes);
            tma_load_2d(my_st + s * STAGE, &map_in, (int)((w & 7) * AT_WT), (int)((w >> 3) * r + i0), &s_bar[warp][s]);
        }
    };
    if (valid(wt)) issue_tile(wt, 0);
    if (valid(wt + wstride)) issue_tile(wt + wstride, 1);
    uint32_t phases = 0u;                       // bit s: parity of stage s's barrier
    int s = 0;
    while (valid(wt)) {
        const int64_t j0 = wt * AT_WT;
        const int64_t j = j0 + 8 * p + cq;
        // this lane's column of the tile (group g = p): its norm, fetched before the wait
        const double pv = (p < NG && j < n) ? vn1[j] : -1.0;
        mbar_wait(&s_bar[warp][s], (phases >> s) & 1u);
        phases ^= 1u << s;
        const unsigned char* in_st = my_st + s * STAGE;
        double c[LG][NG][2];
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = 8 * G + 2 * p + e;
#pragma unroll
                for (int g = 0; g < NG; ++g)
                    c[G][g][e] = (row < L) ? *reinterpret_cast<const double*>(in_st + at_swz(row, (8 * g + cq) * 8)) : 0.0;
            }
        // Z^T = C^T V, even and odd row groups on separate accumulators
        double z[NG][2], zo[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) z[g][0] = z[g][1] = zo[g][0] = zo[g][1] = 0.0;
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double bv = sV[(8 * G + 2 * p + e) * AM_SV + cq];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (G & 1) dmma884(zo[g][0], zo[g][1], c[G][g][e], bv);
                    else dmma884(z[g][0], z[g][1], c[G][g][e], bv);
                }
            }
#pragma unroll
        for (int g = 0; g < NG; ++g) { z[g][0] += zo[g][0]; z[g][1] += zo[g][1]; }
        // every staged element has been consumed by a DMMA of its lane: the stage takes the tile after next
        __syncwarp();
        if (valid(wt + 2 * wstride)) issue_tile(wt + 2 * wstride, s);
        // Z'^T = Z^T T
        double zp[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) zp[g][0] = zp[g][1] = 0.0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double bv = sT[(2 * p + e) * AM_ST + cq];
#pragma unroll
            for (int g = 0; g < NG; ++g) dmma884(zp[g][0], zp[g][1], z[g][e], bv);
        }
        // C^T -= Z'^T V^T
#pragma unroll
        for (int g = 0; g < NG; ++g) { zp[g][0] = -zp[g][0]; zp[g][1] = -zp[g][1]; }
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double bv = sVt[(2 * p + e) * SVT + 8 * G + cq];
#pragma unroll
                for (int g = 0; g < NG; ++g) dmma884(c[G][g][0], c[G][g][1], zp[g][e], bv);
            }
        // rows below the block go back to HBM; their squares give the exact trailing norms
        const int64_t tbase = (j0 >> 7) * ((int64_t)r * OMB_TB) + (int64_t)i0 * OMB_TB + (j0 & (OMB_TB - 1));
        double* out = dst + tbase + (2 * p) * OMB_TB + cq;
        double sq[NG], sq1[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) sq[g] = sq1[g] = 0.0;
#pragma unroll
        for (int G = 0; G < LG; ++G)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = 8 * G + 2 * p + e;
#pragma unroll
                for (int g = 0; g < NG; ++g)
                    if (row > t && row < L) {
                        stg_stream(out + (8 * G + e) * OMB_TB + 8 * g, c[G][g][e]);
                        if (e) sq1[g] = fma(c[G][g][e], c[G][g][e], sq1[g]);
                        else sq[g] = fma(c[G][g][e], c[G][g][e], sq[g]);
                    }
            }
        double mysq = 0.0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            double v = sq[g] + sq1[g];
            v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
            v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
            if (p == g) mysq = v;
        }
        double v1 = pv;                          // < 0: a pivot or no column (p >= NG, j >= n)
        if (v1 > 0.0) {
            v1 = last_row ? 0.0 : sqrt(mysq);
            vn2[j] = v1;
        }
        if (v1 >= 0.0) {
            vn1[j] = v1;
            cand_push_lazy(best, v1, j, sh, P, s_total);
        }
        if (lazy) {
            double wmax = v1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wmax = dmax(wmax, __shfl_xor_sync(0xFFFFFFFFu, wmax, o));
            if (lane == 0)
                atomicMax(reinterpret_cast<long long*>(seg_w + j0 / QR_SEG), __double_as_longlong(wmax));
        }
        wt += wstride;
        s ^= 1;
    }
    best = cand_block_reduce(best, s_c);
    if (threadIdx.x == 0) {
        cand[blockIdx.x] = best;
        if (blockIdx.x == 0) const_cast<Panel*>(P)->ncand = (int)gridDim.x;
    }
}

typedef void (*ApplyTmaFn)(const CUtensorMap, double*, int64_t, int, int, int, int, const Panel*, double*, double*,
                           int64_t, Shard, Cand*, double*);
static ApplyTmaFn pick_apply_tma(int L)
{
    static int lg_min = -1;                    // ($OMB_QR_APPLY_TMA_MIN: shortest block, in groups of 8 rows, that takes this kernel)
    if (lg_min < 0) { const char* e = getenv("OMB_QR_APPLY_TMA_MIN"); lg_min = e ? atoi(e) : 6; }
    if ((L + 7) / 8 < lg_min) return nullptr;
    switch ((L + 7) / 8) {
#define OMB_AT_CASE(LGV) case LGV: return qr_apply_tma_kernel<LGV>;
        OMB_AT_CASE(1) OMB_AT_CASE(2) OMB_AT_CASE(3) OMB_AT_CASE(4) OMB_AT_CASE(5) OMB_AT_CASE(6) OMB_AT_CASE(7)
        OMB_AT_CASE(8) OMB_AT_CASE(9) OMB_AT_CASE(10) OMB_AT_CASE(11) OMB_AT_CASE(12) OMB_AT_CASE(13)
#undef OMB_AT_CASE
        default: return nullptr;
    }
}

// 2-D view of a tiled basis (tiles * r rows of 128 candidates) with a box of 16 candidates x box_rows rows
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int basis_tensor_map(CUtensorMap* m, const double* base, int64_t ntiles, int r, int box_rows)
{
    static TmapEncodeFn enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) return -1;
        enc = (TmapEncodeFn)fp;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)OMB_TB, (cuuint64_t)(ntiles * r)};
    const cuuint64_t strides[1] = {(cuuint64_t)OMB_TB * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)AT_WT, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}

typedef void (*ApplyMmaFn)(const double*, double*, int64_t, int, int, int, int, const Panel*, double*, double*,
                           int64_t, Shard, Cand*, double*);

// tensor-path apply kernel for L rows (L <= 256); *ng = column groups per warp.  exact: the block-closing
// pass of a blocked schedule; !exact: block == 1 with more than QR_LREG trailing rows
static ApplyMmaFn pick_apply_mma(int L, bool exact, int* ng)
{
    const int lg = (L + 7) / 8;
    if (!exact) {
        *ng = lg <= 16 ? 2 : 1;
        if (lg <= 13) return qr_apply_mma_kernel<13, 2>;
        if (lg <= 16) return qr_apply_mma_kernel<16, 2>;
        if (lg <= 24) return qr_apply_mma_kernel<24, 1>;
        return qr_apply_mma_kernel<32, 1>;
    }
    switch (lg) {
#define OMB_AM_CASE(LGV, NGV) case LGV: *ng = NGV; return qr_apply_exact_kernel<LGV, NGV>;
        OMB_AM_CASE(1, 4) OMB_AM_CASE(2, 4) OMB_AM_CASE(3, 4) OMB_AM_CASE(4, 4) OMB_AM_CASE(5, 4) OMB_AM_CASE(6, 4)
        OMB_AM_CASE(7, 4) OMB_AM_CASE(8, 4) OMB_AM_CASE(9, 2) OMB_AM_CASE(10, 2) OMB_AM_CASE(11, 2)
        OMB_AM_CASE(12, 2) OMB_AM_CASE(13, 2) OMB_AM_CASE(14, 2) OMB_AM_CASE(15, 2) OMB_AM_CASE(16, 2)
#undef OMB_AM_CASE
        default: break;
    }
    if (lg <= 24) { *ng = 1; return qr_apply_exact_kernel<24, 1>; }
    *ng = 1;
    return qr_apply_exact_kernel<32, 1>;
}

// ---------------------------------------------------------------------------------------------
// panel kernel (1 CTA): pivot selection + reflector for global step i (block-local index t)
// ---------------------------------------------------------------------------------------------
constexpr int PN_THREADS = 256;

// ---- peer-to-peer exchange over NVLink (symmetric memory): every rank owns a buffer laid out as
//   LL slots [2 parity][world][2 * QR_REC] u64 | (reserved [2][world]) | barrier flags [world] i64 | error i64
// and holds the device addresses of all ranks' buffers.  A record travels in the low-latency format
// of the collective libraries: every 8-byte word carries 4 bytes of payload and a 4-byte step tag,
// so the receiver needs neither a fence nor a separate flag -- a word whose tag matches IS the data
// (8-byte stores are single NVLink transactions).  The sender stores the 2 * (8 + L) words of its
// record into slot [step parity][its rank] of EVERY peer's buffer; the receiver polls its OWN copy.
// Tags only grow (epoch * 4096 + step + 1), the buffer starts zeroed.
struct P2P {
    double* const* peers;     // device array of `world` buffer addresses (NULL: host-gathered path)
    double* mine;             // this rank's buffer
    int64_t epoch;
};
__device__ __forceinline__ unsigned long long* p2p_ll(double* buf, int world, int parity, int src)
{
    return reinterpret_cast<unsigned long long*>(buf) + ((int64_t)parity * world + src) * (2 * QR_REC);
}
__device__ __forceinline__ int64_t* p2p_flags(double* buf, int world)
{
    return reinterpret_cast<int64_t*>(buf + (int64_t)4 * world * QR_REC);
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// publish word k of a record to every peer (payload d, tag)
__device__ __forceinline__ void ll_publish(const P2P& pp, int world, int parity, int rank, int k, double d, unsigned tag)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
    const unsigned long long w0 = (bits & 0xFFFFFFFFull) | ((unsigned long long)tag << 32);
    const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
    for (int g = 0; g < world; ++g) {
        unsigned long long* dst = p2p_ll(pp.peers[g], world, parity, rank) + 2 * k;
        st_relaxed_sys(dst, w0);
        st_relaxed_sys(dst + 1, w1);
    }
}
// wait for word k of rank src's record in MY buffer; gives up after ~10 s and raises the error word
__device__ __forceinline__ double ll_receive(const P2P& pp, int world, int parity, int src, int k, unsigned tag, int64_t* err)
{
    const unsigned long long* p = p2p_ll(pp.mine, world, parity, src) + 2 * k;
    unsigned long long w0, w1, t0 = 0;
    unsigned spins = 0;
    for (;;) {
        w0 = ld_relaxed_sys(p);
        w1 = ld_relaxed_sys(p + 1);
        if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag) break;
        if ((++spins & 0x3FFu) == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 10000000000ull) { *err = 1; break; }
        }
    }
    return __longlong_as_double((long long)((w0 & 0xFFFFFFFFull) | (w1 << 32)));
}
__device__ __forceinline__ void st_release_sys(int64_t* p, int64_t v)
{
    asm volatile("st.release.sys.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int64_t ld_acquire_sys(const int64_t* p)
{
    int64_t v;
    asm volatile("ld.acquire.sys.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p >= want; gives up after ~10 s and raises the buffer's error word
__device__ __forceinline__ void p2p_wait(const int64_t* p, int64_t want, int64_t* err)
{
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_sys(p) < want) {
        if ((++spins & 0x3FFu) == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 10000000000ull) { *err = 1; break; }
        }
    }
}

// all ranks have finished the previous placement before anyone reuses the record slots
__global__ void qr_p2p_barrier_kernel(P2P pp, int rank, int world)
{
    int64_t* myflags = p2p_flags(pp.mine, world);
    const int g = threadIdx.x;
    if (g < world) {
        int64_t* peer_flags = p2p_flags(pp.peers[g], world);
        st_release_sys(peer_flags + 2 * world + rank, pp.epoch);
    }
    __syncthreads();
    if (g < world) p2p_wait(myflags + 2 * world + g, pp.epoch, myflags + 3 * world);
}

// multi-rank step A of the host-gathered path (1 CTA): local argmax + the winner's trailing column ->
// one record in `rec` (the P2P path does this inside the panel kernel)
__global__ void __launch_bounds__(PN_THREADS)
qr_local_kernel(const Panel* __restrict__ P, const Cand* __restrict__ cand, const double* __restrict__ src, int r,
                int i0, int L, int step, Shard sh, double* __restrict__ rec, P2P pp)
{
    __shared__ Cand s_c[PN_THREADS / 32];
    __shared__ double s_rec[QR_REC];
    __shared__ int64_t s_p;
    Cand c = cand_empty();
    const int ncand = P->ncand;
    for (int e = threadIdx.x; e < ncand; e += PN_THREADS) cand_merge(c, cand[e]);
    c = cand_block_reduce(c, s_c);
    if (threadIdx.x == 0) {
        s_p = c.idx;
        s_rec[0] = c.best;
        s_rec[1] = c.second;
        s_rec[2] = __longlong_as_double(c.key);
        s_rec[3] = __longlong_as_double(c.idx >= 0 ? glob_index(c.idx, sh) : -1);
        s_rec[4] = __longlong_as_double(c.idx);
        s_rec[5] = s_rec[6] = s_rec[7] = 0.0;
    }
    __syncthreads();
    const int64_t p = s_p;
    for (int k = threadIdx.x; k < L; k += PN_THREADS)
        s_rec[8 + k] = (p >= 0) ? src[basis_index(i0 + k, p, r)] : 0.0;
    __syncthreads();
    (void)step; (void)pp;
    for (int k = threadIdx.x; k < 8 + L; k += PN_THREADS) rec[k] = s_rec[k];
}

// The panel is pure latency (one CTA between two grid-wide passes), so it is organised around
// dependent-chain length, not work: every load a thread needs is issued before the first use, the
// candidate records are merged by a shuffle tree, and after the pivot column has arrived a single
// warp does the reflector algebra -- the t <= 8 dot products of a step run side by side on the 8
// quads of the warp (4 lanes x L/4 elements each), the tiny triangular products on one lane per row.
template <bool MULTI>
__global__ void __launch_bounds__(PN_THREADS)
qr_panel_kernel(Panel* __restrict__ P, const Cand* __restrict__ cand, int ncand, const double* __restrict__ src,
                int r, int i0, int L, int i, int t, int seq_norm, int phase, int64_t s_total, int64_t index_base, Shard sh,
                const double* recs, P2P pp, double* __restrict__ vn1, int64_t* __restrict__ piv,
                double* __restrict__ rdiag, double* __restrict__ gap)
{
    __shared__ Cand s_c[PN_THREADS / 32];
    __shared__ double s_V[QR_BMAX][QR_RMAX];   // earlier reflectors of this block (rows < t), then v_t
    __shared__ double s_T[QR_BMAX][QR_BMAX + 1];
    __shared__ double s_x[QR_RMAX];            // pivot column tail over rows i0..
    __shared__ double s_z[QR_BMAX], s_g[QR_BMAX];
    __shared__ int64_t s_win[4];               // winner: global index, LAPACK key, local column, gap bits
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_enter();
    // lazy down-dates: phase 0 is the regular panel of step i, phase 1 its repetition after a catch-up
    // pass (launched after every in-block panel, a no-op unless phase 0 raised `retry`)
    const bool lazy = P->lazy != 0;
    if (phase == 1 && (!lazy || P->retry == 0)) return;
    const double theta = (phase == 1) ? P->cstar : P->theta;   // bound on every skipped candidate's norm
    double rest = __longlong_as_double(P->rest_bits);          // largest block-start norm still skipped (this rank)
    const int xseq = P->xseq;

    // 0. block state -> shared memory (independent of the pivot: overlaps the argmax reduction)
    for (int a = warp; a < t; a += PN_THREADS / 32)
        for (int k = lane; k < L; k += 32) s_V[a][k] = P->V[a][k];
    if (threadIdx.x < QR_BMAX * QR_BMAX) {
        const int a = threadIdx.x >> 3, b = threadIdx.x & 7;
        s_T[a][b] = (a <= b && b < t) ? P->T[a][b] : 0.0;      // upper triangular; the rest of P->T is never written
    }

    Cand c = cand_empty();       // winner: idx = GLOBAL row index, key = LAPACK position
    int64_t p_local = -1;        // the winner's local column, -1 if another rank owns it
    const bool p2p = MULTI && pp.peers != nullptr;
    if (!MULTI || p2p) {
        // 1. argmax over this rank's pass-kernel records: all loads first, then the merges
        constexpr int NR = 5;
        const int nc = MULTI ? P->ncand : ncand;
        Cand rc[NR];
#pragma unroll
        for (int u = 0; u < NR; ++u) {
            const int e = threadIdx.x + u * PN_THREADS;
            rc[u] = cand_empty();
            if (e < nc) {
                const double2 lo = *reinterpret_cast<const double2*>(&cand[e].best);
                const longlong2 hi = *reinterpret_cast<const longlong2*>(&cand[e].idx);
                rc[u].best = lo.x; rc[u].second = lo.y; rc[u].idx = hi.x; rc[u].key = hi.y;
            }
        }
#pragma unroll
        for (int u = 0; u < NR; ++u) cand_merge(c, rc[u]);
        for (int e = threadIdx.x + NR * PN_THREADS; e < nc; e += PN_THREADS) cand_merge(c, cand[e]);
        if (phase == 1) {
            const int nc2 = P->ncand2;               // the catch-up pass's records
            for (int e = threadIdx.x; e < nc2; e += PN_THREADS) cand_merge(c, cand[QR_NCAND / 2 + e]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { Cand b = cand_shfl_xor(c, o); cand_merge(c, b); }
        if (lane == 0) s_c[warp] = c;
        __syncthreads();
        c = (lane < PN_THREADS / 32) ? s_c[lane] : cand_empty();
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { Cand b = cand_shfl_xor(c, o); cand_merge(c, b); }
        c.best = __shfl_sync(0xFFFFFFFFu, c.best, 0); c.second = __shfl_sync(0xFFFFFFFFu, c.second, 0);
        c.idx = __shfl_sync(0xFFFFFFFFu, c.idx, 0); c.key = __shfl_sync(0xFFFFFFFFu, c.key, 0);
        p_local = c.idx;
        // 2. pivot column tail at block start (every warp knows the winner: no second barrier)
        const double* col = src + basis_index(i0, p_local < 0 ? 0 : p_local, r);
        for (int k = threadIdx.x; k < L; k += PN_THREADS) s_x[k] = p_local >= 0 ? col[(int64_t)k * OMB_TB] : 0.0;
    }
    if (p2p) {
        // 3. exchange over NVLink: publish (header, column tail) to every peer in the tagged low-latency
        //    format, collect the world's headers, pick the winner in rank order (identical decision on
        //    every rank), then collect the winner's column tail
        __shared__ double s_hdr[64][8];
        // slot parity and tag follow the number of exchanges done (a step that needs a catch-up
        // exchanges twice): slot x & 1 is reused by exchange x + 2, which a rank can only reach after
        // every rank has published exchange x + 1, i.e. has finished reading exchange x
        const unsigned tag = (unsigned)(pp.epoch * 4096 + xseq + 1);
        const int parity = xseq & 1;
        int64_t* err = p2p_flags(pp.mine, sh.world) + 3 * sh.world;
        __syncthreads();                     // s_x complete
        for (int k = threadIdx.x; k < 8 + L; k += PN_THREADS) {
            double d = 0.0;
            if (k == 0) d = c.best;
            else if (k == 1) d = c.second;
            else if (k == 2) d = __longlong_as_double(c.key);
            else if (k == 3) d = __longlong_as_double(c.idx >= 0 ? glob_index(c.idx, sh) : -1);
            else if (k == 4) d = __longlong_as_double(c.idx);
            else if (k == 5) d = rest;
            else if (k >= 8) d = s_x[k - 8];
            ll_publish(pp, sh.world, parity, sh.rank, k, d, tag);
        }
        for (int e = threadIdx.x; e < sh.world * 8; e += PN_THREADS)
            s_hdr[e >> 3][e & 7] = ll_receive(pp, sh.world, parity, e >> 3, e & 7, tag, err);
        __syncthreads();
        if (threadIdx.x == 0) P->xseq = xseq + 1;
        c = cand_empty();
        p_local = -1;
        int wr = -1;
        for (int g = 0; g < sh.world; ++g) {
            Cand b;
            b.best = s_hdr[g][0]; b.second = s_hdr[g][1];
            b.key = __double_as_longlong(s_hdr[g][2]); b.idx = __double_as_longlong(s_hdr[g][3]);
            rest = dmax(rest, s_hdr[g][5]);
            if (b.idx < 0) continue;
            const bool wins = cand_better(b.best, b.key, c.best, c.key);
            cand_merge(c, b);
            if (wins) { wr = g; p_local = (g == sh.rank) ? __double_as_longlong(s_hdr[g][4]) : -1; }
        }
        if (wr < 0) wr = 0;
        for (int k = threadIdx.x; k < L; k += PN_THREADS) s_x[k] = ll_receive(pp, sh.world, parity, wr, 8 + k, tag, err);
    } else if (MULTI) {
        // host-gathered records: winner among the ranks in fixed order
        int wr = -1;
        for (int g = 0; g < sh.world; ++g) {
            const double* rcp = recs + (int64_t)g * QR_REC;
            Cand b;
            b.best = __ldcg(rcp + 0); b.second = __ldcg(rcp + 1);
            b.key = __double_as_longlong(__ldcg(rcp + 2)); b.idx = __double_as_longlong(__ldcg(rcp + 3));
            if (b.idx < 0) continue;
            const bool wins = cand_better(b.best, b.key, c.best, c.key);
            cand_merge(c, b);
            if (wins) { wr = g; p_local = (g == sh.rank) ? __double_as_longlong(__ldcg(rcp + 4)) : -1; }
        }
        const double* rcw = recs + (int64_t)(wr < 0 ? 0 : wr) * QR_REC + 8;
        for (int k = threadIdx.x; k < L; k += PN_THREADS) s_x[k] = __ldcg(rcw + k);
    }
    if (lazy) {
        // every thread holds the same (global) winner.  Inside a block the candidates of the skipped
        // segments are all below theta: the winner is certified iff its norm reaches theta
        const bool certified = (t == 0) || phase == 1 || c.best >= theta;
        if (threadIdx.x == 0) {
            P->retry = certified ? 0 : 1;
            P->rest_bits = __double_as_longlong(-1.0);           // the next pass (or catch-up) collects it anew
            if (!certified) { P->cstar = c.best; P->nretry += 1; }
            else if (t == 0) P->theta = P->alpha * c.best;       // block start: every norm is exact
            else if (phase == 1) P->theta = theta;               // what is still skipped is below cstar
        }
        if (!certified) return;                  // nothing of step i has been touched: catch-up, then phase 1
        if (t > 0 && c.second < rest) c.second = rest;           // the runner-up may sit in a skipped segment
    }
    if (threadIdx.x == 32) {
        s_win[0] = c.idx; s_win[1] = c.key; s_win[2] = p_local;
        s_win[3] = __double_as_longlong((c.second < 0.0 || c.best <= 0.0) ? 1.0 : (c.best - c.second) / c.best);
    }
    __syncthreads();                     // s_x, s_V, s_T complete
    if (warp == 1 && lane == 0) {
        // LAPACK's swap of positions i <-> pos(p) beside the reflector arithmetic of warp 0
        const int64_t p = s_win[0], pos_p = s_win[1], pl = s_win[2];
        piv[i] = p + index_base;
        gap[i] = __longlong_as_double(s_win[3]);
        if (pos_p != i) {
            const int64_t ci = P->col_at_pos[i];       // the column sitting at position i moves to pos(p)
            P->posmap[ci] = pos_p;
            if (pos_p < s_total) P->col_at_pos[pos_p] = ci;
        }
        P->col_at_pos[i] = p;
        if (p < s_total) P->posmap[p] = i;
        if (pl >= 0) vn1[pl] = -1.0;     // never a candidate again
    }
    if (warp != 0) return;

    const int qd = lane >> 2, ql = lane & 3;          // quad qd handles reflector qd in the dot products
    constexpr int NK = QR_RMAX / 32;
    // dots[qd] = V[qd] . y  for the 8 reflectors at once (rows beyond t hold zeros in s_T, so the
    // results of quads >= t are never used)
    auto quad_dots = [&](const double* y) {
        double a0 = 0.0, a1 = 0.0;
        if (qd < t) {
            const double* vq = s_V[qd];
            int k = ql;
            for (; k + 4 < L; k += 8) { a0 = fma(vq[k], y[k], a0); a1 = fma(vq[k + 4], y[k + 4], a1); }
            if (k < L) a0 = fma(vq[k], y[k], a0);
        }
        double sacc = a0 + a1;
        sacc += __shfl_xor_sync(0xFFFFFFFFu, sacc, 1);
        sacc += __shfl_xor_sync(0xFFFFFFFFu, sacc, 2);
        return sacc;
    };

    if (t > 0) {
        // x <- Q^T x = x - V (T^T (V^T x))   (compact WY of the block's earlier reflectors)
        const double zq = quad_dots(s_x);
        if (ql == 0) s_z[qd] = qd < t ? zq : 0.0;
        __syncwarp();
        if (lane < QR_BMAX) {
            double zz = 0.0;
#pragma unroll
            for (int b = 0; b < QR_BMAX; ++b) zz = fma(s_T[b][lane], s_z[b], zz);      // (T^T z)_lane; T is upper triangular
            s_g[lane] = lane < t ? zz : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < NK; ++u) {
            const int k = lane + 32 * u;
            if (k < L) {
                double xv = s_x[k];
                for (int a = 0; a < t; ++a) xv = fma(-s_V[a][k], s_g[a], xv);
                s_x[k] = xv;
            }
        }
        __syncwarp();
    }

    // 3. dlarfg on x[t:]  (alpha = x[t], tail x[t+1:])
    double xn2;
    if (seq_norm) {
        // dnrm2's order (block == 1: bit-identical to dlaqp2): one lane sums the tail sequentially
        double sacc = 0.0;
        if (lane == 0) for (int k = t + 1; k < L; ++k) sacc = fma(s_x[k], s_x[k], sacc);
        xn2 = __shfl_sync(0xFFFFFFFFu, sacc, 0);
    } else {
        double sacc = 0.0;
        for (int k = t + 1 + lane; k < L; k += 32) sacc = fma(s_x[k], s_x[k], sacc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xFFFFFFFFu, sacc, o);
        xn2 = sacc;
    }
    const double alpha = s_x[t];
    double beta = alpha, tau = 0.0, scal = 0.0;
    {
        const double xnorm = sqrt(xn2);
        if (L - t > 1 && xnorm != 0.0) {
            // dlapy2(alpha, xnorm)
            const double xa = fabs(alpha), ya = xnorm;
            const double w = xa > ya ? xa : ya, zmin = xa > ya ? ya : xa;
            double h = w;
            if (zmin != 0.0) { const double qq = zmin / w; h = w * sqrt(1.0 + qq * qq); }
            beta = -copysign(h, alpha);
            tau = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
    }
    if (lane == 0) { rdiag[i] = beta; P->tau[t] = tau; }
    for (int k = lane; k < L; k += 32) {
        double vv = 0.0;
        if (k == t) vv = 1.0;
        else if (k > t) vv = (tau != 0.0) ? s_x[k] * scal : 0.0;
        P->V[t][k] = vv;
        s_V[t][k] = vv;
    }
    __syncwarp();

    // 4. compact-WY column t of T, and q = Q e_t = e_t - V T (V^T e_t)
    const double zv = quad_dots(s_V[t]);               // V[:, qd]^T v_t
    if (ql == 0) s_z[qd] = qd < t ? zv : 0.0;
    __syncwarp();
    if (lane < QR_BMAX) {
        // T[0:t, t] = -tau * T[0:t, 0:t] * (V^T v_t);  T[t][t] = tau
        double val = 0.0;
        if (lane < t) {
            double sacc = 0.0;
            for (int b = lane; b < t; ++b) sacc = fma(s_T[lane][b], s_z[b], sacc);
            val = -tau * sacc;
        } else if (lane == t) {
            val = tau;
        }
        if (lane <= t) { s_T[lane][t] = val; P->T[lane][t] = val; }
    }
    __syncwarp();
    if (lane < QR_BMAX) {
        // g = T * (V^T e_t),  (V^T e_t)_b = V[b][t]
        double gv = 0.0;
        if (lane <= t) for (int b = lane; b <= t; ++b) gv = fma(s_T[lane][b], s_V[b][t], gv);
        s_g[lane] = gv;
    }
    __syncwarp();
    for (int k = lane; k < L; k += 32) {
        double qv = (k == t) ? 1.0 : 0.0;
        for (int a = 0; a <= t; ++a) qv = fma(-s_V[a][k], s_g[a], qv);
        P->qs[t][k] = qv;
    }
}

struct QrWs {
    double* vn1;
    double* vn2;
    Cand* cand;
    Panel* panel;
    double* vn_tmp;
    double* seg_max;           // [2][nseg]: largest norm of every 64-candidate segment at a block start (ping-pong)
    unsigned char* seg_skip;   // [nseg]: the segment sits the current block out
    int64_t nseg;
};

static int64_t qr_ws_layout(int64_t n, QrWs* w, char* base)
{
    int64_t off = 0;
    const int64_t npad = basis_tiles(n) * OMB_TB;
    auto take = [&](int64_t bytes) { int64_t o = off; off += round_up(bytes, 256); return base ? base + o : (char*)nullptr; };
    char* a = take((int64_t)sizeof(double) * npad);
    char* b = take((int64_t)sizeof(double) * npad);
    char* c = take((int64_t)sizeof(Cand) * QR_NCAND);
    char* d = take((int64_t)sizeof(Panel));
    char* e = take((int64_t)sizeof(double) * npad);
    const int64_t nseg = npad / QR_SEG;
    char* f = take((int64_t)sizeof(double) * 2 * nseg);
    char* g = take(nseg);
    if (w) {
        w->vn1 = (double*)a; w->vn2 = (double*)b; w->cand = (Cand*)c; w->panel = (Panel*)d; w->vn_tmp = (double*)e;
        w->seg_max = (double*)f; w->seg_skip = (unsigned char*)g; w->nseg = nseg;
    }
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

namespace omb {

static int64_t gemv_grid(int64_t n)
{
    int64_t g = ceil_div(basis_tiles(n) * (OMB_TB / 2), GV_THREADS);
    static int per_sm = 0;                     // a whole number of waves of resident CTAs: no straggler CTAs
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qr_gemv_kernel<GV_PASS>, GV_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 3;
        per_sm *= 2;                           // two whole waves (measured: 2.77 ms vs 2.82 ms for one or three)
    }
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (g > cap) g = cap;
    if (g > QR_NCAND / 2) g = QR_NCAND / 2;    // the upper half of the record array belongs to the catch-up pass
    return g;
}

// lazy down-dates: theta = alpha * (pivot norm at the block start).  Setting: < 0 automatic (by problem
// size), 0 off, (0, 1) fixed.  OMB_QR_LAZY overrides the default at load time, omb_qrcp_set_lazy() at run time.
static double g_lazy_setting = -2.0;           // -2: not read yet
static double lazy_setting()
{
    if (g_lazy_setting == -2.0) {
        const char* e = getenv("OMB_QR_LAZY");
        g_lazy_setting = e ? atof(e) : -1.0;
        if (!(g_lazy_setting < 1.0)) g_lazy_setting = -1.0;
        if (g_lazy_setting < 0.0) g_lazy_setting = -1.0;
    }
    return g_lazy_setting;
}
// A tighter bound skips more segments but fails its certification more often; a catch-up round costs
// two extra kernels (~25 us) whatever the size, a looser bound costs reads in proportion to n * r.
// Measured optima: 0.90 at 1.65M x 40, 0.92-0.94 at 2M x 100, 0.94-0.95 at 16.2M x 100.
static double lazy_alpha(int64_t n, int r)
{
    const double a = lazy_setting();
    if (a >= 0.0) return a;
    return (double)n * r < 1.0e8 ? 0.90 : 0.94;
}
// the segments' ping-pong buffers: block k reads the one its predecessor's apply pass filled
static double* seg_read(const QrWs& w, int i0, int block) { return w.seg_max + (((i0 / block) + 1) & 1) * w.nseg; }
static double* seg_write(const QrWs& w, int i0, int block) { return w.seg_max + ((i0 / block) & 1) * w.nseg; }

// norms -> vn1/vn2, position maps, and the step-0 argmax records.  Returns the record count (< 0: error code)
static int qr_start(const double* d_Ut, int64_t n, int r, int64_t s, const double* d_vn, const QrWs& w, Shard sh,
                    double alpha, cudaStream_t st, int* ncand)
{
    const int sms = sm_count();
    int rc;
    const double* vn = d_vn;
    int64_t g = ceil_div(n, 256);
    if (g > (int64_t)sms * 8) g = (int64_t)sms * 8;
    if (!vn) {
        qr_norms_kernel<<<(unsigned)g, 256, 0, st>>>(d_Ut, n, r, w.vn_tmp);
        if ((rc = check_launch("qr_norms_kernel"))) return rc;
        vn = w.vn_tmp;
    }
    qr_init_kernel<<<(unsigned)g, 256, 0, st>>>(vn, n, w.vn1, w.vn2, w.panel, alpha);
    if ((rc = check_launch("qr_init_kernel"))) return rc;
    // step-0 argmax: a read-only pass over zero rows leaves the norms untouched (and the segment
    // maxima where block 0 reads them)
    const int64_t gv = gemv_grid(n);
    qr_gemv_kernel<GV_START><<<(unsigned)gv, GV_THREADS, 0, st>>>(d_Ut, n, r, 0, 0, 0, w.panel, w.vn1, w.vn2, s, sh,
                                                        w.cand, (const double*)nullptr, w.seg_max + w.nseg, w.seg_skip);
    if ((rc = check_launch("qr_gemv_kernel"))) return rc;
    *ncand = (int)gv;
    return 0;
}

// the pass that follows the panel of step i (block-local step t of the block starting at i0)
static int qr_pass(const double* src, double* d_work, int64_t n, int r, int64_t s, int block, int i0, int t,
                   const QrWs& w, Shard sh, cudaStream_t st, int* ncand)
{
    const int sms = sm_count();
    const int L = r - i0;
    const int64_t ntiles = basis_tiles(n);
    int rc;
    if (t == block - 1) {
        int64_t g;
        int lmax = 0, ng = 0;
        Apply1Fn f1 = (block == 1) ? pick_apply1(L, &lmax) : nullptr;
        if (f1) {
            g = ntiles;
            if (g > (int64_t)sms * ar_min_blocks(lmax)) g = (int64_t)sms * ar_min_blocks(lmax);   // one wave
            if (g > QR_NCAND) g = QR_NCAND;
            launch_pdl(f1, dim3((unsigned)g), dim3(AR_THREADS), 0, st, src, d_work, n, r, i0, L, w.panel, w.vn1, w.vn2, s, sh, w.cand);
            if ((rc = check_launch("qr_apply1_kernel"))) return rc;
        } else {
            // (block == 1 with more than QR_LREG trailing rows also lands here: same algorithm,
            //  tensor-path rounding instead of the oracle's fma order)
            // the block-closing pass on the TMA engine ($OMB_QR_APPLY_TMA=0: the register-staged kernel)
            static int use_tma = -1;
            if (use_tma < 0) { const char* e = getenv("OMB_QR_APPLY_TMA"); use_tma = e ? atoi(e) : 1; }
            ApplyTmaFn ft = (block > 1 && use_tma && (((uintptr_t)src) & 127) == 0) ? pick_apply_tma(L) : nullptr;
            CUtensorMap map_in;
            if (ft && basis_tensor_map(&map_in, src, ntiles, r, L)) ft = nullptr;
            if (ft) {
                const int lp = ((L + 7) / 8) * 8;
                const size_t stage = (size_t)AT_WARPS * 2 * lp * 128;
                OMB_CUDA(cudaFuncSetAttribute(ft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
                g = sms;                       // one CTA of 8 independent warp pipelines per SM
                const int64_t need = ceil_div(ntiles * (OMB_TB / AT_WT), AT_WARPS);
                if (g > need) g = need;
                launch_pdl(ft, dim3((unsigned)g), dim3(AT_THREADS), stage, st, map_in, d_work, n, r, i0, L, t, (const Panel*)w.panel,
                           w.vn1, w.vn2, s, sh, w.cand, seg_write(w, i0, block));
                if ((rc = check_launch("qr_apply_tma_kernel"))) return rc;
                *ncand = (int)g;
                return 0;
            }
            ApplyMmaFn fm = pick_apply_mma(L, block > 1, &ng);
            g = ceil_div(ntiles * (OMB_TB / (8 * ng)), AM_THREADS / 32);
            int per_sm = 0;                    // resident CTAs per SM of this instantiation: one full wave
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fm, AM_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 3;
            if (g > (int64_t)sms * per_sm) g = (int64_t)sms * per_sm;
            if (g > QR_NCAND) g = QR_NCAND;
            launch_pdl(fm, dim3((unsigned)g), dim3(AM_THREADS), 0, st, src, d_work, n, r, i0, L, t, w.panel, w.vn1, w.vn2, s, sh, w.cand,
                       seg_write(w, i0, block));
            if ((rc = check_launch("qr_apply_mma_kernel"))) return rc;
        }
        *ncand = (int)g;
    } else {
        const int64_t gv = gemv_grid(n);
        launch_pdl(qr_gemv_kernel<GV_PASS>, dim3((unsigned)gv), dim3(GV_THREADS), 0, st, src, n, r, i0, L, t,
                   w.panel, w.vn1, w.vn2, s, sh, w.cand, (const double*)seg_read(w, i0, block), seg_write(w, i0, block),
                   w.seg_skip);
        if ((rc = check_launch("qr_gemv_kernel"))) return rc;
        *ncand = (int)gv;
    }
    return 0;
}

// lazy down-dates: the conditional catch-up pass after the panel of in-block step t >= 1 (the panel
// is then launched a second time with phase = 1); both exit at once unless that panel raised `retry`
static int qr_catchup(const double* src, int64_t n, int r, int64_t s, int block, int i0, int t, const QrWs& w, Shard sh,
                      cudaStream_t st)
{
    int64_t g = gemv_grid(n);
    const int64_t small = (int64_t)sm_count() * 4;     // few segments join: a small grid keeps the no-op launch cheap
    if (g > small) g = small;
    launch_pdl(qr_gemv_kernel<GV_CATCHUP>, dim3((unsigned)g), dim3(GV_THREADS), 0, st, src, n, r, i0, r - i0, t,
               w.panel, w.vn1, w.vn2, s, sh, w.cand, (const double*)seg_read(w, i0, block), seg_write(w, i0, block),
               w.seg_skip);
    return check_launch("qr_gemv_kernel");
}

static int qr_check(const void* d_Ut, const void* d_work, const void* d_ws, int64_t n, int64_t r, int64_t s, int block)
{
    OMB_CHECK_ARG(d_Ut && d_work && d_ws, "null pointer");
    OMB_CHECK_ARG(n > 0 && r > 0 && s > 0, "non-positive size");
    OMB_CHECK_ARG(r <= QR_RMAX, "r exceeds the supported number of modes (256)");
    OMB_CHECK_ARG(s <= r, "s must be <= r");
    OMB_CHECK_ARG(block >= 1 && block <= QR_BMAX, "block must be in [1, 8]");
    OMB_CHECK_ARG((((uintptr_t)d_Ut | (uintptr_t)d_work) & 15) == 0, "basis pointers must be 16-byte aligned");
    return 0;
}

}  // namespace omb

extern "C" int omb_qrcp(const double* d_Ut, int64_t n, int64_t r, int64_t s, const double* d_vn, double* d_work,
                        void* d_ws, int block, int64_t index_base, int64_t* d_piv, double* d_rdiag, double* d_gap,
                        void* stream)
{
    int rc = qr_check(d_Ut, d_work, d_ws, n, r, s, block);
    if (rc) return rc;
    OMB_CHECK_ARG(d_piv && d_rdiag && d_gap, "null pointer");
    OMB_CHECK_ARG(s <= n, "s must be <= n");
    cudaStream_t st = (cudaStream_t)stream;
    QrWs w;
    qr_ws_layout(n, &w, (char*)d_ws);
    const int ri = (int)r;
    Shard sh;
    sh.n_c_loc = n; sh.n_c = n; sh.cell0 = 0; sh.rank = 0; sh.world = 1;
    int ncand = 0;
    const double alpha = block > 1 ? lazy_alpha(n, ri) : 0.0;
    if ((rc = qr_start(d_Ut, n, ri, s, d_vn, w, sh, alpha, st, &ncand))) return rc;
    const double* src = d_Ut;   // trailing matrix as of the block start, rows i0..r-1
    int i0 = 0;
    for (int i = 0; i < (int)s; ++i) {
        const int t = i - i0;
        for (int phase = 0; phase < ((alpha > 0.0 && t > 0) ? 2 : 1); ++phase) {
            if (phase == 1 && (rc = qr_catchup(src, n, ri, s, block, i0, t, w, sh, st))) return rc;
            launch_pdl(qr_panel_kernel<false>, dim3(1), dim3(PN_THREADS), 0, st, w.panel, (const Cand*)w.cand, ncand, src, ri,
                       i0, ri - i0, i, t, block == 1 ? 1 : 0, phase, s, index_base, sh, (const double*)nullptr,
                       P2P{nullptr, nullptr, 0}, w.vn1, d_piv, d_rdiag, d_gap);
            if ((rc = check_launch("qr_panel_kernel"))) return rc;
        }
        if (i == (int)s - 1) break;           // no further pivot needed: skip the last pass
        if ((rc = qr_pass(src, d_work, n, ri, s, block, i0, t, w, sh, st, &ncand))) return rc;
        if (t == block - 1) { src = d_work; i0 = i + 1; }
    }
    return 0;
}

// ---- multi-rank stepping interface (one process per GPU; the caller all-gathers the records) ----
static Shard make_shard(int64_t n_c_loc, int64_t n_c, int64_t cell0, int rank, int world)
{
    Shard sh;
    sh.n_c_loc = n_c_loc; sh.n_c = n_c; sh.cell0 = cell0; sh.rank = rank; sh.world = world;
    return sh;
}

extern "C" int64_t omb_qrcp_record_doubles(void) { return QR_REC; }

extern "C" int omb_qrcp_mr_start(const double* d_Ut, int64_t n, int64_t r, int64_t s, const double* d_vn, void* d_ws,
                                 int64_t n_c_loc, int64_t n_c, int64_t cell0, int rank, int world, void* stream)
{
    OMB_CHECK_ARG(d_Ut && d_ws, "null pointer");
    OMB_CHECK_ARG(n > 0 && r > 0 && r <= QR_RMAX && s > 0 && s <= r, "bad size");
    OMB_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && n_c_loc > 0 && n % n_c_loc == 0, "bad shard");
    QrWs w;
    qr_ws_layout(n, &w, (char*)d_ws);
    int ncand = 0;
    // the host-gathered exchange steps from the host: no lazy down-dates (their retry is decided on the device)
    return qr_start(d_Ut, n, (int)r, s, d_vn, w, make_shard(n_c_loc, n_c, cell0, rank, world), 0.0, (cudaStream_t)stream,
                    &ncand);
}

// step i, part A: this rank's best candidate and its trailing column -> d_rec (QR_REC doubles)
extern "C" int omb_qrcp_mr_local(const double* d_Ut, const double* d_work, int64_t n, int64_t r, void* d_ws, int block,
                                 int64_t i, int64_t n_c_loc, int64_t n_c, int64_t cell0, int rank, int world,
                                 double* d_rec, void* stream)
{
    OMB_CHECK_ARG(d_Ut && d_work && d_ws && d_rec, "null pointer");
    OMB_CHECK_ARG(block >= 1 && block <= QR_BMAX && i >= 0 && i < r, "bad step");
    QrWs w;
    qr_ws_layout(n, &w, (char*)d_ws);
    const int i0 = (int)(i / block) * block;
    const double* src = i0 == 0 ? d_Ut : d_work;
    qr_local_kernel<<<1, PN_THREADS, 0, (cudaStream_t)stream>>>(w.panel, w.cand, src, (int)r, i0, (int)r - i0, (int)i,
                                                               make_shard(n_c_loc, n_c, cell0, rank, world), d_rec,
                                                               P2P{nullptr, nullptr, 0});
    return check_launch("qr_local_kernel");
}

// step i, part B: the winner among the `world` gathered records, reflector, then this rank's pass
extern "C" int omb_qrcp_mr_step(const double* d_Ut, double* d_work, int64_t n, int64_t r, int64_t s, void* d_ws,
                                int block, int64_t i, int64_t n_c_loc, int64_t n_c, int64_t cell0, int rank, int world,
                                const double* d_recs, int64_t* d_piv, double* d_rdiag, double* d_gap, void* stream)
{
    int rc = qr_check(d_Ut, d_work, d_ws, n, r, s, block);
    if (rc) return rc;
    OMB_CHECK_ARG(d_recs && d_piv && d_rdiag && d_gap && i >= 0 && i < s, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    QrWs w;
    qr_ws_layout(n, &w, (char*)d_ws);
    const Shard sh = make_shard(n_c_loc, n_c, cell0, rank, world);
    const int ri = (int)r;
    const int i0 = (int)(i / block) * block, t = (int)i - i0;
    const double* src = i0 == 0 ? d_Ut : d_work;
    qr_panel_kernel<true><<<1, PN_THREADS, 0, st>>>(w.panel, w.cand, 0, src, ri, i0, ri - i0, (int)i, t, 0, 0, s, 0, sh,
                                                     d_recs, P2P{nullptr, nullptr, 0}, w.vn1, d_piv, d_rdiag, d_gap);
    if ((rc = check_launch("qr_panel_kernel"))) return rc;
    if (i == s - 1) return 0;
    int ncand = 0;
    return qr_pass(src, d_work, n, ri, s, block, i0, t, w, sh, st, &ncand);
}

// ---- multi-rank placement with the per-step exchange done by the kernels themselves over NVLink
// peer memory (no NCCL call, no host round trip inside the loop).  d_peers: device array of `world`
// symmetric-buffer addresses (entry `rank` == d_mine), each omb_qrcp_p2p_buffer_doubles(world) doubles,
// zero-filled once at allocation.  epoch: strictly increasing per call, identical on every rank.
extern "C" int64_t omb_qrcp_p2p_buffer_doubles(int world) { return (int64_t)4 * world * QR_REC + 3 * world + 8; }
// index (in doubles) of the buffer's error word: non-zero after a peer failed to answer within 10 s
extern "C" int64_t omb_qrcp_p2p_error_index(int world) { return (int64_t)4 * world * QR_REC + 3 * world; }

extern "C" int omb_qrcp_p2p(const double* d_Ut, int64_t n, int64_t r, int64_t s, const double* d_vn, double* d_work,
                            void* d_ws, int block, int64_t n_c_loc, int64_t n_c, int64_t cell0, int rank, int world,
                            const void* d_peers, double* d_mine, int64_t epoch, int64_t* d_piv, double* d_rdiag,
                            double* d_gap, void* stream)
{
    int rc = qr_check(d_Ut, d_work, d_ws, n, r, s, block);
    if (rc) return rc;
    OMB_CHECK_ARG(d_peers && d_mine && d_piv && d_rdiag && d_gap, "null pointer");
    OMB_CHECK_ARG(world >= 2 && world <= 64 && rank >= 0 && rank < world && epoch > 0 && s < 2048, "bad p2p argument");
    cudaStream_t st = (cudaStream_t)stream;
    QrWs w;
    qr_ws_layout(n, &w, (char*)d_ws);
    const Shard sh = make_shard(n_c_loc, n_c, cell0, rank, world);
    const P2P pp{(double* const*)d_peers, d_mine, epoch};
    const int ri = (int)r;
    qr_p2p_barrier_kernel<<<1, 64, 0, st>>>(pp, rank, world);
    if ((rc = check_launch("qr_p2p_barrier_kernel"))) return rc;
    int ncand = 0;
    const double alpha = block > 1 ? lazy_alpha(n, ri) : 0.0;
    if ((rc = qr_start(d_Ut, n, ri, s, d_vn, w, sh, alpha, st, &ncand))) return rc;
    const double* src = d_Ut;
    int i0 = 0;
    for (int i = 0; i < (int)s; ++i) {
        const int t = i - i0;
        // (the certification uses the GLOBAL winner, so every rank takes the same retry decisions)
        for (int phase = 0; phase < ((alpha > 0.0 && t > 0) ? 2 : 1); ++phase) {
            if (phase == 1 && (rc = qr_catchup(src, n, ri, s, block, i0, t, w, sh, st))) return rc;
            launch_pdl(qr_panel_kernel<true>, dim3(1), dim3(PN_THREADS), 0, st, w.panel, (const Cand*)w.cand, 0, src, ri, i0,
                       ri - i0, i, t, 0, phase, s, (int64_t)0, sh, (const double*)nullptr, pp, w.vn1, d_piv, d_rdiag, d_gap);
            if ((rc = check_launch("qr_panel_kernel"))) return rc;
        }
        if (i == (int)s - 1) break;
        if ((rc = qr_pass(src, d_work, n, ri, s, block, i0, t, w, sh, st, &ncand))) return rc;
        if (t == block - 1) { src = d_work; i0 = i + 1; }
    }
    return 0;
}

// ---- lazy down-dates: switch and statistics ----
// alpha in (0, 1): segments whose largest norm at a block start is below alpha * (pivot norm) sit the
// block's read-only passes out (exact: see the file header); 0 switches the scheme off; negative =
// automatic (by problem size; the default, or $OMB_QR_LAZY).  Returns the previous setting (-1 =
// automatic).  Process-wide.
extern "C" double omb_qrcp_set_lazy(double alpha)
{
    const double prev = lazy_setting();
    g_lazy_setting = alpha < 0.0 ? -1.0 : ((alpha > 0.0 && alpha < 1.0) ? alpha : 0.0);
    return prev;
}

// what the read-only passes of the last placement on this workspace actually visited:
// out[0] = (segment, row) visits (64 candidates x 8 bytes each), out[1] = segment visits,
// out[2] = catch-up rounds, out[3] = 1 if the lazy scheme was on, out[4] = its alpha in parts per
// million, out[5] = 0 (reserved).  Synchronises the stream.
extern "C" int omb_qrcp_stats(const void* d_ws, int64_t n, int64_t* out, void* stream)
{
    OMB_CHECK_ARG(d_ws && out && n > 0, "bad argument");
    QrWs w;
    qr_ws_layout(n, &w, (char*)const_cast<void*>(d_ws));
    Panel* hp = (Panel*)malloc(sizeof(Panel));
    OMB_CHECK_ARG(hp != nullptr, "out of host memory");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(hp, w.panel, sizeof(Panel), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { free(hp); set_error("omb_qrcp_stats: %s", cudaGetErrorString(e)); return (int)e; }
    out[0] = (int64_t)hp->seg_rows; out[1] = (int64_t)hp->seg_visits; out[2] = hp->nretry; out[3] = hp->lazy;
    out[4] = (int64_t)(hp->alpha * 1.0e6 + 0.5); out[5] = 0;
    free(hp);
    return 0;
}
