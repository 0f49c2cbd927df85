// p2p.cu -- small all-gathers over NVLink peer memory (multi-GPU plumbing of the row-sharded path).
//
// The cross-rank objects of the path are tiny (F*4 block statistics, the m x m Gram, the s x (r+1)
// sensor rows); an NCCL collective costs tens of microseconds of launch and
// protocol latency for each of them.  Here one small kernel stores a rank's payload straight into
// every peer's symmetric buffer in the tagged low-latency format (every 8-byte word = 4 bytes of
// payload + a 4-byte sequence tag: a word whose tag matches IS the data -- no fence, no flag) and
// polls its own buffer for the peers' words.  Layout of a rank's buffer (u64 words):
//     [2 parity][world][2 * capacity]   followed by one error word
// Successive calls alternate parity; a slot is only rewritten two calls later, which a peer can
// reach only after this rank has published (i.e. finished reading) the call in between.
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

__device__ __forceinline__ void p2p_st(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long p2p_ld(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(512)
p2p_allgather_kernel(const double* __restrict__ src, int64_t n, double* __restrict__ out,
                     unsigned long long* const* __restrict__ peers, unsigned long long* __restrict__ mine,
                     int64_t capacity, unsigned tag, int parity, int rank, int world)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t slot = ((int64_t)parity * world + rank) * 2 * capacity;
    // publish: my words into slot [parity][rank] of every peer (myself included)
    for (int64_t k = tid; k < n; k += nth) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(src[k]);
        const unsigned long long w0 = (bits & 0xFFFFFFFFull) | ((unsigned long long)tag << 32);
        const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
        for (int g = 0; g < world; ++g) {
            unsigned long long* dst = peers[g] + slot + 2 * k;
            p2p_st(dst, w0);
            p2p_st(dst + 1, w1);
        }
    }
    // collect: every rank's words from my own buffer
    unsigned long long* err = mine + (int64_t)2 * world * 2 * capacity;
    for (int64_t e = tid; e < n * world; e += nth) {
        const int g = (int)(e / n);
        const int64_t k = e - (int64_t)g * n;
        const unsigned long long* p = mine + ((int64_t)parity * world + g) * 2 * capacity + 2 * k;
        unsigned long long w0, w1, t0 = 0;
        unsigned spins = 0;
        for (;;) {
            w0 = p2p_ld(p);
            w1 = p2p_ld(p + 1);
            if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag) break;
            if ((++spins & 0x3FFu) == 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t0 == 0) t0 = t1;
                else if (t1 - t0 > 10000000000ull) { *err = 1; break; }       // a peer never answered (10 s)
            }
        }
        out[e] = __longlong_as_double((long long)((w0 & 0xFFFFFFFFull) | (w1 << 32)));
    }
}

// All-reduce variant: the same publish step, then thread k combines element k of every rank IN RANK
// ORDER while it collects (identical bits on every rank), so a statistics exchange is this one kernel
// instead of an all-gather followed by a handful of framework element-wise kernels.
//   op 0: out[k] = ((x_0[k] + x_1[k]) + x_2[k]) + ...
//   op 1: block statistics, first pass: groups of 4 = {sum, min, max, -}: sum / min / max / rank 0's value
//   op 2: block statistics, second pass: element 3 of every group summed, the others rank 0's value
__global__ void __launch_bounds__(512)
p2p_allreduce_kernel(const double* __restrict__ src, int64_t n, double* __restrict__ out,
                     unsigned long long* const* __restrict__ peers, unsigned long long* __restrict__ mine,
                     int64_t capacity, unsigned tag, int parity, int rank, int world, int op)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t slot = ((int64_t)parity * world + rank) * 2 * capacity;
    for (int64_t k = tid; k < n; k += nth) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(src[k]);
        const unsigned long long w0 = (bits & 0xFFFFFFFFull) | ((unsigned long long)tag << 32);
        const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
        for (int g = 0; g < world; ++g) {
            unsigned long long* dst = peers[g] + slot + 2 * k;
            p2p_st(dst, w0);
            p2p_st(dst + 1, w1);
        }
    }
    unsigned long long* err = mine + (int64_t)2 * world * 2 * capacity;
    for (int64_t k = tid; k < n; k += nth) {
        const int col = (int)(k & 3);
        double acc = 0.0;
        for (int g = 0; g < world; ++g) {
            const unsigned long long* p = mine + ((int64_t)parity * world + g) * 2 * capacity + 2 * k;
            unsigned long long w0, w1, t0 = 0;
            unsigned spins = 0;
            for (;;) {
                w0 = p2p_ld(p);
                w1 = p2p_ld(p + 1);
                if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag) break;
                if ((++spins & 0x3FFu) == 0) {
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t0 == 0) t0 = t1;
                    else if (t1 - t0 > 10000000000ull) { *err = 1; break; }   // a peer never answered (10 s)
                }
            }
            const double v = __longlong_as_double((long long)((w0 & 0xFFFFFFFFull) | (w1 << 32)));
            if (g == 0) { acc = v; continue; }
            if (op == 0) acc = acc + v;
            else if (op == 1) {
                if (col == 0) acc = acc + v;
                else if (col == 1) acc = (v < acc || v != v) ? v : acc;       // min, NaN propagates like torch.min
                else if (col == 2) acc = (v > acc || v != v) ? v : acc;
            } else if (col == 3) acc = acc + v;
        }
        out[k] = acc;
    }
}

}  // namespace omb

using namespace omb;

extern "C" int64_t omb_p2p_allgather_buffer_doubles(int world, int64_t capacity)
{
    if (world < 1 || capacity < 1) return 0;
    return (int64_t)2 * world * 2 * capacity + 8;
}
extern "C" int64_t omb_p2p_allgather_error_index(int world, int64_t capacity) { return (int64_t)2 * world * 2 * capacity; }

extern "C" int omb_p2p_allgather(const double* d_src, int64_t n, double* d_out, const void* d_peers, double* d_mine,
                                 int64_t capacity, int64_t seq, int rank, int world, void* stream)
{
    OMB_CHECK_ARG(d_src && d_out && d_peers && d_mine, "null pointer");
    OMB_CHECK_ARG(world >= 2 && rank >= 0 && rank < world, "bad rank/world");
    OMB_CHECK_ARG(n > 0 && n <= capacity, "payload exceeds the buffer capacity");
    OMB_CHECK_ARG(seq > 0, "seq must be positive and strictly increasing");
    const unsigned tag = (unsigned)((seq % 0xFFFFFFFEll) + 1);                   // never 0 (the buffer starts zeroed)
    int64_t g = ceil_div(n * world, 512);
    if (g > 16) g = 16;
    p2p_allgather_kernel<<<(unsigned)g, 512, 0, (cudaStream_t)stream>>>(
        d_src, n, d_out, (unsigned long long* const*)d_peers, (unsigned long long*)d_mine, capacity, tag, (int)(seq & 1),
        rank, world);
    return check_launch("p2p_allgather_kernel");
}

extern "C" int omb_p2p_allreduce(const double* d_src, int64_t n, double* d_out, const void* d_peers, double* d_mine,
                                 int64_t capacity, int64_t seq, int rank, int world, int op, void* stream)
{
    OMB_CHECK_ARG(d_src && d_out && d_peers && d_mine, "null pointer");
    OMB_CHECK_ARG(world >= 2 && rank >= 0 && rank < world, "bad rank/world");
    OMB_CHECK_ARG(n > 0 && n <= capacity, "payload exceeds the buffer capacity");
    OMB_CHECK_ARG(seq > 0, "seq must be positive and strictly increasing");
    OMB_CHECK_ARG(op >= 0 && op <= 2 && (op == 0 || (n & 3) == 0), "bad op");
    const unsigned tag = (unsigned)((seq % 0xFFFFFFFEll) + 1);
    int64_t g = ceil_div(n, 512);
    if (g > 16) g = 16;
    p2p_allreduce_kernel<<<(unsigned)g, 512, 0, (cudaStream_t)stream>>>(
        d_src, n, d_out, (unsigned long long* const*)d_peers, (unsigned long long*)d_mine, capacity, tag, (int)(seq & 1),
        rank, world, op);
    return check_launch("p2p_allreduce_kernel");
}
