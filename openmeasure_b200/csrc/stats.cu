// stats.cu -- K1: centring/scaling statistics, bit-exact with numpy's pairwise add.reduce.
//
// Replaces np.average(x, axis=1), np.std(x), np.max(x), np.min(x) over the n_cells-row feature
// blocks (reference sparse_sensing.py:110-161).  numpy reduces a contiguous FP64 range with a
// fixed tree: ranges > 128 elements split at n2 = (n/2) & ~7, leaves (<= 128 elements) are summed
// with 8 interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail.
// The kernels evaluate exactly that tree: a quad of lanes owns one leaf (lane q = accumulators 2q and
// 2q+1, one 128-bit load per octet, all loads of a leaf issued before its first add), a warp owns
// the 8 leaves of a depth-(D-3) node and combines them by shuffles, and a small second kernel sweeps
// the levels above.  Nodes that do not exist at a given depth are tracked with a flag, not with
// +0.0, so even the sign of zero follows numpy.  (Row means of m <= 64 snapshots come from the Gram
// kernel's fragments instead, gram.cu; the 8-lane row kernel below serves m > 1024.)
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

// ---------------------------------------------------------------------------------------------
// leaf and subtree evaluation
// ---------------------------------------------------------------------------------------------
struct MapId {
    __device__ __forceinline__ double operator()(double x) const { return x; }
};
struct MapSqDev {
    double mu;
    __device__ __forceinline__ double operator()(double x) const
    {
        double d = x - mu;
        return d * d;
    }
};

// Sum of a leaf (n <= 128) by an 8-lane group; every lane of the group returns the result.
// `gmask` is the shuffle mask of the group's 8 lanes.  Optionally tracks min/max of raw values.
template <class Map, bool MINMAX>
__device__ __forceinline__ double leaf_sum(const double* __restrict__ a, int64_t n, int l8, unsigned gmask,
                                           Map f, double& lo, double& hi)
{
    if (n < 8) {
        double res = -0.0;
        for (int64_t i = 0; i < n; ++i) {
            double x = a[i];
            if (MINMAX) { lo = fmin(lo, x); hi = fmax(hi, x); }
            res += f(x);
        }
        return res;
    }
    double x = a[l8];
    if (MINMAX) { lo = fmin(lo, x); hi = fmax(hi, x); }
    double r = f(x);
    const int64_t nfull = n - (n % 8);
    for (int64_t i = 8; i < nfull; i += 8) {
        x = a[i + l8];
        if (MINMAX) { lo = fmin(lo, x); hi = fmax(hi, x); }
        r += f(x);
    }
    r += __shfl_xor_sync(gmask, r, 1);
    r += __shfl_xor_sync(gmask, r, 2);
    r += __shfl_xor_sync(gmask, r, 4);
    for (int64_t i = nfull; i < n; ++i) {
        x = a[i];
        if (MINMAX) { lo = fmin(lo, x); hi = fmax(hi, x); }
        r += f(x);
    }
    return r;
}

// Full pairwise tree of a range by one 8-lane group (used for rows with m > 128).
template <class Map>
__device__ double tree_sum_group(const double* __restrict__ a, int64_t n, int l8, unsigned gmask, Map f)
{
    if (n <= 128) {
        double lo = 0, hi = 0;
        return leaf_sum<Map, false>(a, n, l8, gmask, f, lo, hi);
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    double left = tree_sum_group(a, n2, l8, gmask, f);
    double right = tree_sum_group(a + n2, n - n2, l8, gmask, f);
    return left + right;
}

// Walk `depth` levels down numpy's tree from (off, n) following the bits of `path` (MSB first).
// Returns false when that slot does not exist (an ancestor is already a leaf and this is not its
// left-most descendant).
__device__ __forceinline__ bool descend(int64_t& off, int64_t& n, uint32_t path, int depth)
{
    for (int d = depth - 1; d >= 0; --d) {
        if (n <= 128) return (path & ((2u << d) - 1u)) == 0u;
        const int64_t n2 = (int64_t)(((uint64_t)n >> 1) & ~(uint64_t)7);      // n/2 rounded down to 8
        if ((path >> d) & 1u) { off += n2; n -= n2; }
        else n = n2;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// row means: one 8-lane group per row
// ---------------------------------------------------------------------------------------------
constexpr int RM_THREADS = 256;

__global__ void __launch_bounds__(RM_THREADS)
row_means_kernel(const double* __restrict__ X, int64_t rows, int64_t m, double* __restrict__ cnt)
{
    const int l8 = threadIdx.x & 7;
    const unsigned gmask = 0xFFu << (threadIdx.x & 24);
    const int64_t g0 = ((int64_t)blockIdx.x * RM_THREADS + threadIdx.x) >> 3;
    const int64_t gstride = ((int64_t)gridDim.x * RM_THREADS) >> 3;
    const double dm = (double)m;
    for (int64_t row = g0; row < rows; row += gstride) {
        double s = tree_sum_group(X + row * m, m, l8, gmask, MapId());
        if (l8 == 0) cnt[row] = s / dm;
    }
}

// ---------------------------------------------------------------------------------------------
// block tree sum.  Tree depth D = LT + LC: a CTA evaluates one depth-LT node (a subtree of 2^LC
// slots) and writes its value to the level-LT array; a second kernel sweeps the top LT levels.
// ---------------------------------------------------------------------------------------------
constexpr int BS_THREADS = 128;               // 4 independent warps: no CTA barrier in the sweep
constexpr int BS_CTAS_PER_SM = 5;
constexpr int BS_GX_MAX = 4096;               // CTAs per feature (bounds the min/max partial arrays)

// Tree depth D = LW + LQ: a WARP evaluates one depth-LW node (a subtree of 2^LQ <= 8 leaf slots, one
// per quad) and writes its value to the level-LW array; a second kernel sweeps the top LW levels.
struct BlockPlan { int D, LW, LQ; };

// Exact depth of numpy's tree over n0 elements: the distinct range sizes of each level are few, so
// the levels are walked as small sets until every range is a leaf (<= 128 elements).
static BlockPlan make_plan(int64_t n0)
{
    int64_t cur[256], nxt[256];
    int nc = 1, D = 0;
    cur[0] = n0;
    for (;;) {
        int nn = 0;
        bool split = false;
        for (int i = 0; i < nc; ++i) {
            const int64_t n = cur[i];
            if (n <= 128) continue;
            split = true;
            const int64_t n2 = (n / 2) & ~(int64_t)7;
            const int64_t kids[2] = {n2, n - n2};
            for (int k = 0; k < 2; ++k) {
                bool seen = false;
                for (int j = 0; j < nn; ++j) seen |= (nxt[j] == kids[k]);
                if (!seen && nn < 256) nxt[nn++] = kids[k];
            }
        }
        if (!split) break;
        ++D;
        nc = nn;
        for (int i = 0; i < nn; ++i) cur[i] = nxt[i];
    }
    BlockPlan p;
    p.D = D;
    p.LQ = D < 3 ? D : 3;
    p.LW = D - p.LQ;
    return p;
}

// Running min/max by compare-select (one DSETP + selects; fmin/fmax cost ~10 instructions each in
// FP64 because of their NaN rules).  A NaN never wins a comparison; the block SUM still turns NaN,
// and block_top_kernel restores numpy's NaN result for min/max from it.
__device__ __forceinline__ void track_minmax(double x, double& lo, double& hi)
{
    lo = (x < lo) ? x : lo;
    hi = (x > hi) ? x : hi;
}

// Sum of a leaf (n <= 128) by a QUAD: lane q owns numpy's accumulators 2q and 2q+1 and fetches them
// with one 128-bit load per octet.  All (<= 16) octet loads are issued before the first add, so a
// thread keeps up to 256 bytes in flight: the pass is bound by HBM, not by the add chain.
// Every lane of the quad returns the result.  `vec`: the leaf is 16-byte aligned.
template <class Map, bool MINMAX>
__device__ __forceinline__ double leaf_sum_quad(const double* __restrict__ a, int n, int q, unsigned qmask, bool vec,
                                                Map f, double& lo, double& hi, double2* keep = nullptr)
{
    if (n < 8) {
        double res = -0.0;
        for (int i = 0; i < n; ++i) {
            double x = a[i];
            if (MINMAX) track_minmax(x, lo, hi);
            res += f(x);
        }
        return res;
    }
    const int noct = n >> 3;
    double2 x[16];
    const double* p = a + 2 * q;
    if (vec) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < noct) x[i] = ldg_stream2(p + 8 * i);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < noct) { x[i].x = ldg_stream(p + 8 * i); x[i].y = ldg_stream(p + 8 * i + 1); }
    }
    if (keep) {
#pragma unroll
        for (int i = 0; i < 16; ++i) keep[i] = x[i];
    }
    double r0 = 0.0, r1 = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (i < noct) {
            if (MINMAX) { track_minmax(x[i].x, lo, hi); track_minmax(x[i].y, lo, hi); }
            if (i == 0) { r0 = f(x[i].x); r1 = f(x[i].y); }
            else { r0 += f(x[i].x); r1 += f(x[i].y); }
        }
    }
    double r = r0 + r1;                       // (r0+r1), (r2+r3), (r4+r5), (r6+r7)
    r += __shfl_xor_sync(qmask, r, 1);
    r += __shfl_xor_sync(qmask, r, 2);
    for (int i = noct * 8; i < n; ++i) {
        double xx = a[i];
        if (MINMAX) track_minmax(xx, lo, hi);
        r += f(xx);
    }
    return r;
}

// one up-sweep level by shuffles: slot `sl` (a multiple of 2*half) absorbs slot sl+half if present
__device__ __forceinline__ double upsweep_level(double v, int sl, int half, int lane_stride, unsigned present_bits)
{
    const double pv = __shfl_xor_sync(0xFFFFFFFFu, v, half * lane_stride);
    if ((sl & (2 * half - 1)) == 0 && ((present_bits >> (sl + half)) & 1u)) v = v + pv;
    return v;
}

// (offset, length) of every depth-LW node of numpy's tree over block_elems elements (length -1: the node does not
// exist, its range ended in a leaf higher up).  The walk from the root is ~20 levels of dependent 64-bit integer
// work; done per node inside the sweep it sat between two batches of loads and kept a third of the warps' time
// without memory requests in flight (3.8 - 4.5 TB/s).  One table per call, shared by all features.
__global__ void __launch_bounds__(256)
node_table_kernel(int64_t block_elems, int LW, int64_t* __restrict__ table)
{
    const int64_t nwn = (int64_t)1 << LW;
    for (int64_t wn = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wn < nwn; wn += (int64_t)gridDim.x * blockDim.x) {
        int64_t off = 0, n = block_elems;
        const bool present = descend(off, n, (uint32_t)wn, LW);
        table[2 * wn] = off;
        table[2 * wn + 1] = present ? n : -1;
    }
}

template <int MODE>   // 0: sum/min/max   1: sum of squared deviations about stats[f*4]/count
__global__ void __launch_bounds__(BS_THREADS, BS_CTAS_PER_SM)
block_tree_kernel(const double* __restrict__ X, int64_t block_elems, int LW, int LQ,
                  const double* __restrict__ stats, double mean_count, double* __restrict__ top_val,
                  unsigned char* __restrict__ top_flag, double* __restrict__ cta_min,
                  double* __restrict__ cta_max, const int64_t* __restrict__ table)
{
    const int f = blockIdx.y;
    const int64_t nwn = (int64_t)1 << LW;
    const double* base = X + (int64_t)f * block_elems;
    const bool vec = (reinterpret_cast<uintptr_t>(base) & 15) == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane & 3, sl = lane >> 2;       // one quad per leaf slot of the warp's node
    const unsigned qmask = 0xFu << (lane & 28);
    const int nslots = 1 << LQ;
    double mu = 0.0;
    if (MODE == 1) mu = stats[f * 4 + 0] / mean_count;
    double lo = __longlong_as_double(0x7FF0000000000000LL), hi = -lo;

    const int64_t nwarps = (int64_t)gridDim.x * (BS_THREADS / 32);
    int64_t wn = (int64_t)blockIdx.x * (BS_THREADS / 32) + warp;
    longlong2 nxt = make_longlong2(0, -1);
    if (wn < nwn) nxt = *reinterpret_cast<const longlong2*>(table + 2 * wn);
    for (; wn < nwn; wn += nwarps) {
        int64_t off = nxt.x, n = nxt.y;
        const bool present = n >= 0;
        if (wn + nwarps < nwn) nxt = *reinterpret_cast<const longlong2*>(table + 2 * (wn + nwarps));   // next node's range
        double v = 0.0;
        bool here = false;
        if (present && sl < nslots) {
            here = descend(off, n, (uint32_t)sl, LQ);
            if (here) {
                if (MODE == 0) v = leaf_sum_quad<MapId, true>(base + off, (int)n, q, qmask, vec, MapId(), lo, hi);
                else { MapSqDev mp; mp.mu = mu; v = leaf_sum_quad<MapSqDev, false>(base + off, (int)n, q, qmask, vec, mp, lo, hi); }
            }
        }
        // up-sweep inside the warp: a node's value ends in its left-most slot
        const unsigned hb = __ballot_sync(0xFFFFFFFFu, here);
        unsigned pres = 0;                         // bit s: slot s of this warp's node is present
#pragma unroll
        for (int sidx = 0; sidx < 8; ++sidx) pres |= ((hb >> (4 * sidx)) & 1u) << sidx;
        v = upsweep_level(v, sl, 1, 4, pres);
        v = upsweep_level(v, sl, 2, 4, pres);
        v = upsweep_level(v, sl, 4, 4, pres);
        if (lane == 0) {
            top_val[(int64_t)f * nwn + wn] = present ? v : 0.0;
            top_flag[(int64_t)f * nwn + wn] = present ? 1 : 0;
        }
    }
    if (MODE == 0) {
        __shared__ double s_lo[BS_THREADS / 32], s_hi[BS_THREADS / 32];
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
        }
        if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < BS_THREADS / 32; ++k) { lo = fmin(lo, s_lo[k]); hi = fmax(hi, s_hi[k]); }
            cta_min[(int64_t)f * gridDim.x + blockIdx.x] = lo;
            cta_max[(int64_t)f * gridDim.x + blockIdx.x] = hi;
        }
    }
}

// Row means for rows of up to 1024 snapshots: the block-tree machinery applied per row.  A row's
// numpy tree has S = 2^D <= 8 leaf slots; a warp evaluates 8 / S rows at once, one quad per slot
// (128-bit loads, up to 256 bytes in flight per lane), and combines a row's slots by shuffles.
template <bool COPY>
__global__ void __launch_bounds__(BS_THREADS, COPY ? 4 : BS_CTAS_PER_SM)
row_means_quad_kernel(const double* __restrict__ X, int64_t rows, int m, int D, double* __restrict__ cnt,
                      double* __restrict__ X0c)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane & 3, sl = lane >> 2;
    const unsigned qmask = 0xFu << (lane & 28);
    const int S = 1 << D, rows_per_warp = 8 >> D;
    const bool xal = (reinterpret_cast<uintptr_t>(X) & 15) == 0;
    const double dm = (double)m;
    const int64_t nwarps = (int64_t)gridDim.x * (BS_THREADS / 32);
    const int64_t ngroups = ceil_div(rows, (int64_t)rows_per_warp);
    for (int64_t grp = (int64_t)blockIdx.x * (BS_THREADS / 32) + warp; grp < ngroups; grp += nwarps) {
        const int64_t row = grp * rows_per_warp + (sl >> D);
        int64_t off = 0, n = m;
        bool here = row < rows;
        if (here) here = descend(off, n, (uint32_t)(sl & (S - 1)), D);
        double v = 0.0, lo = 0.0, hi = 0.0;
        double2 keep[16];
        if (here) {
            const double* a = X + row * m + off;
            const bool vec = xal && (((row * m) & 1) == 0);
            v = leaf_sum_quad<MapId, false>(a, (int)n, q, qmask, vec, MapId(), lo, hi, COPY ? keep : nullptr);
        }
        const unsigned hb = __ballot_sync(0xFFFFFFFFu, here);
        unsigned pres = 0;
#pragma unroll
        for (int sidx = 0; sidx < 8; ++sidx) pres |= ((hb >> (4 * sidx)) & 1u) << sidx;
        if (S > 1) v = upsweep_level(v, sl, 1, 4, pres);
        if (S > 2) v = upsweep_level(v, sl, 2, 4, pres);
        if (S > 4) v = upsweep_level(v, sl, 4, 4, pres);
        const double mean = v / dm;
        if (q == 0 && (sl & (S - 1)) == 0 && row < rows) cnt[row] = mean;
        if (COPY) {
            // centred copy X0c = X - mean for the many-snapshot contraction kernels (m even, X 16-byte aligned),
            // straight from the registers the leaf sums were formed from: HBM sees one read of X and one write of
            // X0c.  DADD shares the FP64 pipe with DMMA -- centring inside the tensor-core kernels, once per tile an
            // element takes part in, cost them a quarter of the pipe.
            const double mr = __shfl_sync(0xFFFFFFFFu, mean, 4 * ((sl >> D) << D));     // the mean of this lane's row
            if (here) {
                double* dst = X0c + row * m + off;
                const int noct = (n >= 8) ? (int)(n >> 3) : 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (i < noct) {
                        double2 x = keep[i];
                        x.x -= mr; x.y -= mr;
                        stg_stream2(dst + 8 * i + 2 * q, x);
                    }
                }
                if (q == 0) {
                    const double* a = X + row * m + off;
                    for (int i = noct * 8; i < (int)n; ++i) dst[i] = a[i] - mr;       // ragged tail of the leaf
                }
            }
        }
    }
}

// X0c[i][j] = X[i][j] - cnt[i]  (centring values given: block means, or row means of wide rows)
__global__ void __launch_bounds__(256)
center_given_kernel(const double* __restrict__ X, int64_t rows, int64_t m, const double* __restrict__ cnt,
                    double* __restrict__ X0c)
{
    const int64_t half = m >> 1, total = rows * half;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / half;
        const double c = cnt[row];
        double2 x = ldg_stream2(X + 2 * e);
        x.x -= c; x.y -= c;
        stg_stream2(X0c + 2 * e, x);
    }
}

// Many CTAs: every aligned run of 1024 level-LT node values is swept (10 tree levels, same adjacent-pair order) in
// shared memory and leaves its value in the run's first slot; the one-CTA-per-feature sweep below then starts at
// stride 1024 (it used to walk all 2^19 values of a config-3 feature block through global memory: 0.57 ms).
constexpr int BM_RUN = 1024;
__global__ void __launch_bounds__(BM_RUN / 2)
block_mid_kernel(int LT, double* __restrict__ top_val, const unsigned char* __restrict__ top_flag)
{
    __shared__ double sv[BM_RUN];
    __shared__ unsigned char sf[BM_RUN];
    const int64_t ntop = (int64_t)1 << LT;
    const int64_t base = (int64_t)blockIdx.y * ntop + (int64_t)blockIdx.x * BM_RUN;
    for (int e = threadIdx.x; e < BM_RUN; e += blockDim.x) { sv[e] = top_val[base + e]; sf[e] = top_flag[base + e]; }
    __syncthreads();
    for (int half = 1; half < BM_RUN; half <<= 1) {
        for (int t = threadIdx.x; t * 2 * half + half < BM_RUN; t += blockDim.x) {
            const int idx = t * 2 * half;
            if (sf[idx + half]) sv[idx] = sv[idx] + sv[idx + half];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) top_val[base] = sv[0];
}

// One CTA per feature: sweep the top LT levels in place (from stride half0), reduce min/max, write stats.
template <int MODE>
__global__ void __launch_bounds__(1024)
block_top_kernel(int LT, int64_t half0, double* __restrict__ top_val, const unsigned char* __restrict__ top_flag,
                 const double* __restrict__ top_min, const double* __restrict__ top_max, int nmm,
                 double* __restrict__ stats)
{
    const int f = blockIdx.x;
    const int64_t ntop = (int64_t)1 << LT;
    double* val = top_val + (int64_t)f * ntop;
    const unsigned char* flag = top_flag + (int64_t)f * ntop;
    for (int64_t half = half0; half < ntop; half <<= 1) {
        for (int64_t t = threadIdx.x; t * 2 * half + half < ntop; t += blockDim.x) {
            int64_t idx = t * 2 * half;
            if (flag[idx + half]) val[idx] = val[idx] + val[idx + half];
        }
        __syncthreads();
    }
    if (MODE == 0) {
        __shared__ double s_lo[32], s_hi[32];
        double lo = __longlong_as_double(0x7FF0000000000000LL), hi = -lo;
        for (int t = threadIdx.x; t < nmm; t += blockDim.x) {
            lo = fmin(lo, top_min[(int64_t)f * nmm + t]);
            hi = fmax(hi, top_max[(int64_t)f * nmm + t]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
            const double inf = __longlong_as_double(0x7FF0000000000000LL);
            // a NaN element makes the sum NaN (inf - inf does too, but then min/max are -inf/+inf)
            if (val[0] != val[0] && !(lo == -inf && hi == inf)) lo = hi = val[0];
            stats[f * 4 + 0] = val[0];
            stats[f * 4 + 1] = lo;
            stats[f * 4 + 2] = hi;
        }
    } else {
        if (threadIdx.x == 0) stats[f * 4 + 3] = val[0];
    }
}

// ---------------------------------------------------------------------------------------------
// scale factors from block statistics (one thread per feature), optional scalar centring
// ---------------------------------------------------------------------------------------------
__global__ void finalize_scale_kernel(const double* __restrict__ stats, int F, double count, int scale_type,
                                      double* __restrict__ scl)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double sum = stats[f * 4 + 0], lo = stats[f * 4 + 1], hi = stats[f * 4 + 2], q = stats[f * 4 + 3];
    const double mean = sum / count;
    const double var = q / count;
    const double sd = sqrt(var);
    double s = 1.0;
    switch (scale_type) {
        case OMB_SCALE_STD: s = sd; break;
        case OMB_SCALE_NONE: s = 1.0; break;
        case OMB_SCALE_PARETO: s = sqrt(sd); break;
        case OMB_SCALE_VAST: s = (sd * sd) / mean; break;
        case OMB_SCALE_RANGE: s = hi - lo; break;
        case OMB_SCALE_LEVEL: s = mean; break;
        case OMB_SCALE_MAX: s = hi; break;
        case OMB_SCALE_VARIANCE: s = var; break;
        case OMB_SCALE_POISSON: s = sqrt(mean); break;
        case OMB_SCALE_L2NORM: s = sqrt(fma(count * mean, mean, q)); break;  // sum x^2 = q + N mean^2
        default: break;
    }
    scl[f] = s;
}

__global__ void fill_cnt_kernel(const double* __restrict__ stats, double count, int64_t n_c_loc,
                                double* __restrict__ cnt)
{
    const int f = blockIdx.y;
    const double mean = stats[f * 4 + 0] / count;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_c_loc;
         i += (int64_t)gridDim.x * blockDim.x)
        cnt[(int64_t)f * n_c_loc + i] = mean;
}

// X0 = (X - cnt) / scl  (IEEE subtraction and division, same bits as numpy broadcasting)
__global__ void __launch_bounds__(256)
scale_rows_kernel(const double* __restrict__ X, int64_t n_c, int64_t m, const double* __restrict__ cnt,
                  const double* __restrict__ scl, double* __restrict__ X0)
{
    const int f = blockIdx.y;
    const double s = scl ? scl[f] : 1.0;
    const int64_t total = n_c * m;
    const double* xb = X + (int64_t)f * total;
    double* ob = X0 + (int64_t)f * total;
    const double* cb = cnt ? cnt + (int64_t)f * n_c : nullptr;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / m;
        const double c = cb ? cb[row] : 0.0;
        ob[e] = (xb[e] - c) / s;
    }
}

__global__ void __launch_bounds__(256)
unscale_kernel(const double* __restrict__ x0, const double* __restrict__ cnt, const double* __restrict__ scl,
               int64_t n_c, int64_t n, double* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x0[i];
        if (scl) v = scl[i / n_c] * v;
        if (cnt) v = v + cnt[i];
        out[i] = v;
    }
}

}  // namespace omb

using namespace omb;

extern "C" int omb_unscale(const double* d_x0, const double* d_cnt, const double* d_scl, int64_t n_c, int64_t n,
                           double* d_out, void* stream)
{
    OMB_CHECK_ARG(d_x0 && d_out, "null pointer");
    OMB_CHECK_ARG(n > 0 && n_c > 0, "non-positive size");
    int64_t g = ceil_div(n, 256);
    if (g > (int64_t)sm_count() * 8) g = (int64_t)sm_count() * 8;
    unscale_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(d_x0, d_cnt, d_scl, n_c, n, d_out);
    return check_launch("unscale_kernel");
}

extern "C" int omb_row_means(const double* d_X, int64_t rows, int64_t m, double* d_cnt, void* stream)
{
    OMB_CHECK_ARG(d_X && d_cnt, "null pointer");
    OMB_CHECK_ARG(rows > 0 && m > 0, "non-positive size");
    const BlockPlan rp = make_plan(m);
    if (m >= 32 && rp.D <= 3) {
        int64_t g = ceil_div(ceil_div(rows, (int64_t)(8 >> rp.D)), BS_THREADS / 32);
        const int64_t cap2 = (int64_t)sm_count() * BS_CTAS_PER_SM;
        if (g > cap2) g = cap2;
        row_means_quad_kernel<false><<<(unsigned)g, BS_THREADS, 0, (cudaStream_t)stream>>>(d_X, rows, (int)m, rp.D, d_cnt, nullptr);
        return check_launch("row_means_quad_kernel");
    }
    int64_t groups_per_cta = RM_THREADS / 8;
    int64_t grid = ceil_div(rows, groups_per_cta);
    int64_t cap = (int64_t)sm_count() * 32;
    if (grid > cap) grid = cap;
    row_means_kernel<<<(unsigned)grid, RM_THREADS, 0, (cudaStream_t)stream>>>(d_X, rows, m, d_cnt);
    return check_launch("row_means_kernel");
}

// X0c[i][j] = X[i][j] - cnt[i] for j < m, 0 for m <= j < ld: the centred copy with an EVEN row pitch for an odd
// snapshot count (the tensor-core kernels need 16-byte aligned rows; a zero snapshot changes neither the Gram of the
// first m columns nor the back-projection)
__global__ void __launch_bounds__(256)
center_pad_kernel(const double* __restrict__ X, int64_t rows, int64_t m, int64_t ld, const double* __restrict__ cnt,
                  double* __restrict__ X0c)
{
    const int64_t total = rows * ld;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / ld, j = e - row * ld;
        X0c[e] = j < m ? X[row * m + j] - cnt[row] : 0.0;
    }
}

// dst[tile][q][128] = src[tile][q][128] for q < r_dst (r_dst <= r_src): drops the zero padding mode(s) of a basis
__global__ void __launch_bounds__(256)
copy_modes_kernel(const double* __restrict__ src, int r_src, double* __restrict__ dst, int r_dst, int64_t ntiles)
{
    const int64_t per = (int64_t)r_dst * OMB_TB / 2, total = ntiles * per;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tile = e / per, k = e - tile * per;
        stg_stream2(dst + tile * ((int64_t)r_dst * OMB_TB) + 2 * k, ldg_stream2(src + tile * ((int64_t)r_src * OMB_TB) + 2 * k));
    }
}

static int center_given(const double* d_X, int64_t rows, int64_t m, const double* d_cnt, double* d_X0c, cudaStream_t st)
{
    int64_t g = ceil_div(rows * (m >> 1), 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    center_given_kernel<<<(unsigned)g, 256, 0, st>>>(d_X, rows, m, d_cnt, d_X0c);
    return check_launch("center_given_kernel");
}

extern "C" int omb_center_rows(const double* d_X, int64_t rows, int64_t m, int compute_means, double* d_cnt,
                               double* d_X0c, void* stream)
{
    OMB_CHECK_ARG(d_X && d_cnt && d_X0c, "null pointer");
    OMB_CHECK_ARG(rows > 0 && m > 0 && (m & 1) == 0, "m must be positive and even");
    OMB_CHECK_ARG(((reinterpret_cast<uintptr_t>(d_X) | reinterpret_cast<uintptr_t>(d_X0c)) & 15) == 0, "X, X0c must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (!compute_means) return center_given(d_X, rows, m, d_cnt, d_X0c, st);
    const BlockPlan rp = make_plan(m);
    if (m >= 32 && rp.D <= 3) {
        int64_t g = ceil_div(ceil_div(rows, (int64_t)(8 >> rp.D)), BS_THREADS / 32);
        const int64_t cap2 = (int64_t)sm_count() * 4;
        if (g > cap2) g = cap2;
        row_means_quad_kernel<true><<<(unsigned)g, BS_THREADS, 0, st>>>(d_X, rows, (int)m, rp.D, d_cnt, d_X0c);
        return check_launch("row_means_quad_kernel");
    }
    int rc = omb_row_means(d_X, rows, m, d_cnt, stream);
    if (rc) return rc;
    return center_given(d_X, rows, m, d_cnt, d_X0c, st);
}

extern "C" int64_t omb_block_stats_ws_bytes(int64_t F, int64_t block_elems)
{
    if (F <= 0 || block_elems <= 0) return 0;
    BlockPlan p = make_plan(block_elems);
    int64_t ntop = (int64_t)1 << p.LW;
    // level-LW values + flags, per-CTA min/max partials, per feature; the node table (offset, length per node)
    return F * ntop * (int64_t)sizeof(double) + round_up(F * ntop, 256) + 2 * F * BS_GX_MAX * (int64_t)sizeof(double) + 256 +
           2 * ntop * (int64_t)sizeof(int64_t);
}

extern "C" int omb_block_stats(const double* d_X, int64_t F, int64_t block_elems, int mode,
                               int64_t mean_count, double* d_out, void* d_ws, void* stream)
{
    OMB_CHECK_ARG(d_X && d_out && d_ws, "null pointer");
    OMB_CHECK_ARG(F > 0 && block_elems > 0, "non-positive size");
    OMB_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 or 1");
    BlockPlan p = make_plan(block_elems);
    OMB_CHECK_ARG(p.LW <= 30, "block too large");
    const int64_t ntop = (int64_t)1 << p.LW;
    double* top_val = (double*)d_ws;
    double* top_min = top_val + F * ntop;
    double* top_max = top_min + F * BS_GX_MAX;
    unsigned char* top_flag = (unsigned char*)(top_max + F * BS_GX_MAX);
    int64_t* table = (int64_t*)((((uintptr_t)(top_flag + round_up(F * ntop, 256))) + 15) & ~(uintptr_t)15);
    {
        int64_t gt = ceil_div(ntop, 256);
        if (gt > 2048) gt = 2048;
        node_table_kernel<<<(unsigned)gt, 256, 0, (cudaStream_t)stream>>>(block_elems, p.LW, table);
        int rc0 = check_launch("node_table_kernel");
        if (rc0) return rc0;
    }
    // persistent warps: ~BS_CTAS_PER_SM CTAs per SM over all features, each warp strides over nodes
    int64_t gx = ceil_div((int64_t)sm_count() * BS_CTAS_PER_SM, F);
    const int64_t need = ceil_div(ntop, BS_THREADS / 32);
    if (gx > need) gx = need;
    if (gx > BS_GX_MAX) gx = BS_GX_MAX;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)F);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0)
        block_tree_kernel<0><<<grid, BS_THREADS, 0, st>>>(d_X, block_elems, p.LW, p.LQ, d_out, (double)mean_count, top_val,
                                                           top_flag, top_min, top_max, table);
    else
        block_tree_kernel<1><<<grid, BS_THREADS, 0, st>>>(d_X, block_elems, p.LW, p.LQ, d_out, (double)mean_count, top_val,
                                                           top_flag, top_min, top_max, table);
    int rc = check_launch("block_tree_kernel");
    if (rc) return rc;
    int64_t half0 = 1;
    if (ntop >= 4 * BM_RUN) {
        dim3 mg((unsigned)(ntop / BM_RUN), (unsigned)F);
        block_mid_kernel<<<mg, BM_RUN / 2, 0, st>>>(p.LW, top_val, top_flag);
        if ((rc = check_launch("block_mid_kernel"))) return rc;
        half0 = BM_RUN;
    }
    if (mode == 0)
        block_top_kernel<0><<<(unsigned)F, 1024, 0, st>>>(p.LW, half0, top_val, top_flag, top_min, top_max, (int)gx, d_out);
    else
        block_top_kernel<1><<<(unsigned)F, 1024, 0, st>>>(p.LW, half0, top_val, top_flag, top_min, top_max, (int)gx, d_out);
    return check_launch("block_top_kernel");
}

extern "C" int omb_finalize_scale(const double* d_stats, int64_t F, int64_t count, int scale_type,
                                  double* d_scl, int fill_cnt, double* d_cnt, int64_t n_c_loc, void* stream)
{
    OMB_CHECK_ARG(d_stats && d_scl, "null pointer");
    OMB_CHECK_ARG(F > 0 && count > 0, "non-positive size");
    OMB_CHECK_ARG(scale_type >= 0 && scale_type <= OMB_SCALE_L2NORM, "unknown scale_type");
    cudaStream_t st = (cudaStream_t)stream;
    finalize_scale_kernel<<<(unsigned)ceil_div(F, 64), 64, 0, st>>>(d_stats, (int)F, (double)count, scale_type, d_scl);
    int rc = check_launch("finalize_scale_kernel");
    if (rc) return rc;
    if (fill_cnt) {
        OMB_CHECK_ARG(d_cnt && n_c_loc > 0, "fill_cnt needs d_cnt and n_c_loc");
        int64_t gx = ceil_div(n_c_loc, 256);
        if (gx > 4096) gx = 4096;
        dim3 grid((unsigned)gx, (unsigned)F);
        fill_cnt_kernel<<<grid, 256, 0, st>>>(d_stats, (double)count, n_c_loc, d_cnt);
        rc = check_launch("fill_cnt_kernel");
    }
    return rc;
}

extern "C" int omb_scale_rows(const double* d_X, int64_t F, int64_t n_c, int64_t m, const double* d_cnt,
                              const double* d_scl, double* d_X0, void* stream)
{
    OMB_CHECK_ARG(d_X && d_X0, "null pointer");
    OMB_CHECK_ARG(F > 0 && n_c > 0 && m > 0, "non-positive size");
    int64_t gx = ceil_div(n_c * m, 256);
    int64_t cap = (int64_t)sm_count() * 16;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, (unsigned)F);
    scale_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, n_c, m, d_cnt, d_scl, d_X0);
    return check_launch("scale_rows_kernel");
}

extern "C" int omb_center_rows_padded(const double* d_X, int64_t rows, int64_t m, int64_t ld_out, int compute_means,
                                      double* d_cnt, double* d_X0c, void* stream)
{
    OMB_CHECK_ARG(d_X && d_cnt && d_X0c, "null pointer");
    OMB_CHECK_ARG(rows > 0 && m > 0 && ld_out >= m, "bad size");
    if (ld_out == m && (m & 1) == 0 && ((reinterpret_cast<uintptr_t>(d_X) | reinterpret_cast<uintptr_t>(d_X0c)) & 15) == 0)
        return omb_center_rows(d_X, rows, m, compute_means, d_cnt, d_X0c, stream);
    if (compute_means) {
        int rc = omb_row_means(d_X, rows, m, d_cnt, stream);
        if (rc) return rc;
    }
    int64_t g = ceil_div(rows * ld_out, 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    center_pad_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(d_X, rows, m, ld_out, d_cnt, d_X0c);
    return check_launch("center_pad_kernel");
}

extern "C" int omb_copy_modes(const double* d_src, int64_t r_src, double* d_dst, int64_t r_dst, int64_t n, void* stream)
{
    OMB_CHECK_ARG(d_src && d_dst, "null pointer");
    OMB_CHECK_ARG(n > 0 && r_dst > 0 && r_dst <= r_src, "bad size");
    const int64_t ntiles = basis_tiles(n);
    int64_t g = ceil_div(ntiles * r_dst * (OMB_TB / 2), 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    copy_modes_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(d_src, (int)r_src, d_dst, (int)r_dst, ntiles);
    return check_launch("copy_modes_kernel");
}
