// common.cuh -- shared helpers for libomb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace omb {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define OMB_CHECK_ARG(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            omb::set_error("%s: invalid argument: %s", __func__, msg); \
            return -1;                                             \
        }                                                          \
    } while (0)

#define OMB_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            omb::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
            return (int)e__;                                                        \
        }                                                                           \
    } while (0)

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int sm_count();

// Basis layout ("tiled mode-major"): the n candidate rows are cut into tiles of OMB_TB; a tile
// stores its r modes back to back, Ut[tile][q][OMB_TB].  Every pass over the trailing rows
// [i0, r) of a tile is then ONE contiguous (r - i0) * OMB_TB * 8-byte burst in HBM.
constexpr int OMB_TB = 128;
__host__ __device__ static inline int64_t basis_tiles(int64_t n) { return (n + OMB_TB - 1) / OMB_TB; }
__host__ __device__ static inline int64_t basis_index(int64_t q, int64_t j, int64_t r)
{
    return (j >> 7) * (r << 7) + (q << 7) + (j & (OMB_TB - 1));
}

// streaming (evict-first) 128-bit and 64-bit global accesses for data touched once per pass
__device__ __forceinline__ double2 ldg_stream2(const double* p)
{
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg_stream(const double* p)
{
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream2(double* p, double2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(double* p, double v)
{
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// ---- TMA (bulk async copy) helpers: contiguous shared <-> global bursts.  SASS: UBLKCP ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// make this thread's generic-proxy shared-memory writes visible to the async (TMA) proxy
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk load completing on an mbarrier
__device__ __forceinline__ void tma_load_bulk(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// ---- TMA tensor copies (2-D tiled tensor maps, cuTensorMapEncodeTiled on the host).  SASS: UTMALDG / UTMASTG ----
// global -> shared: box of the map at coordinates (c0 = innermost, c1), completing on an mbarrier
__device__ __forceinline__ void tma_load_2d(void* sdst, const void* tmap, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(sdst)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// shared -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const void* tmap, int c0, int c1, const void* ssrc)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(smem_u32(ssrc)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the sources of all committed bulk stores have been read (their shared memory may be overwritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... and have been written to global memory
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// FP64 tensor-core MMA: D(8x8) += A(8x4) * B(4x8).  SASS: DMMA.8x8x4.
//   a  = A[lane/4][lane%4],  b = B[lane%4][lane/4],  c0/c1 = C[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace omb
