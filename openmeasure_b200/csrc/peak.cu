// peak.cu -- measured FP64 tensor-pipe roofline of THIS GPU, taken in the same process as the bench line.
//
// bench.py divides the FP64-bound stages (Gram, back-projection, reconstruct) by this number instead
// of a spec-sheet figure: MEASURED_PEAKS.json only holds the HBM copy bandwidth and the BF16 GEMM
// rate.  The kernel is the densest DMMA.8x8x4 stream the SM accepts: 8 warps per CTA, 2 CTAs per SM,
// 8 independent accumulator chains per warp, operands in registers (no memory traffic at all), so a
// real kernel can only approach it.  `ms_target` sets the duration: ~5 ms gives the burst figure
// (boost clocks), several hundred ms the sustained one under the power cap.
#include "common.cuh"
#include "../../include/omb200.h"

namespace omb {

__global__ void __launch_bounds__(256, 2) fp64_peak_kernel(double* out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = (double)threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace omb

using namespace omb;

extern "C" int omb_fp64_peak(double ms_target, double* d_scratch, int64_t scratch_doubles, double* h_tflops,
                             double* h_ms, void* stream)
{
    OMB_CHECK_ARG(d_scratch && h_tflops, "null pointer");
    const int blocks = sm_count() * 2, threads = 256;
    OMB_CHECK_ARG(scratch_doubles >= (int64_t)blocks * threads, "scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    OMB_CUDA(cudaEventCreate(&e0));
    OMB_CUDA(cudaEventCreate(&e1));
    auto run = [&](int iters, float* ms) -> int {
        OMB_CUDA(cudaEventRecord(e0, st));
        fp64_peak_kernel<<<blocks, threads, 0, st>>>(d_scratch, iters, 1.0000001, 1e-9);
        int rc = check_launch("fp64_peak_kernel");
        if (rc) return rc;
        OMB_CUDA(cudaEventRecord(e1, st));
        OMB_CUDA(cudaEventSynchronize(e1));
        OMB_CUDA(cudaEventElapsedTime(ms, e0, e1));
        return 0;
    };
    float ms = 0.f;
    int rc = run(2048, &ms);                       // warm-up + calibration
    if (rc) return rc;
    rc = run(2048, &ms);
    if (rc) return rc;
    double want = ms_target > 0.5 ? ms_target : 0.5;
    int64_t iters = (int64_t)(2048.0 * want / (ms > 1e-3 ? ms : 1e-3));
    if (iters < 2048) iters = 2048;
    if (iters > (1 << 30)) iters = 1 << 30;
    rc = run((int)iters, &ms);
    if (rc) return rc;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flop = 2.0 * 256.0 * 8.0 * (double)iters * (threads / 32) * blocks;
    *h_tflops = flop / (ms * 1e-3) / 1e12;
    if (h_ms) *h_ms = ms;
    return 0;
}
