// fp64 throughput microbenchmark: plain DFMA against the fp64 tensor-core instruction (DMMA, mma.sync m8n8k4).
// (include/overiva_b200.h: oiva_fp64_peak.)  It answers two questions with numbers from the machine the library runs
// on: (1) what fp64 rate the FMA-bound covariance shapes (many channels: M = 16, K = 4) can be held against, and
// (2) whether re-shaping the Hermitian rank-1 accumulation into 8x8x4 matrix products could pay: the real-embedded
// Gram form needs ~1.7x the multiply-adds of the Hermitian-aware DFMA form (cov.cuh), so DMMA only wins if it
// sustains well over 1.7x the DFMA rate.  SURVEY.md section 7.1(12) asked for this evaluation.
#include "common.cuh"

namespace oiva {

constexpr int PEAK_CHAINS = 8;  // independent accumulators per thread (DFMA) / accumulator fragments per warp (DMMA)

__global__ void __launch_bounds__(256) k_peak_dfma(double* __restrict__ out, int iters, double x, double y) {
    double a[PEAK_CHAINS];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) a[i] = (double)(threadIdx.x + i);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < PEAK_CHAINS; ++i) a[i] = fma(a[i], y, x);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s += a[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keeps the chains alive, never true in practice
}

__global__ void __launch_bounds__(256) k_peak_dmma(double* __restrict__ out, int iters, double x, double y) {
    double c0[PEAK_CHAINS], c1[PEAK_CHAINS];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) {
        c0[i] = (double)(threadIdx.x + i);
        c1[i] = (double)(threadIdx.x - i);
    }
    const double a = x + 1e-9 * threadIdx.x, b = y;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < PEAK_CHAINS; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i])
                             : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace oiva

// kind 0: DFMA, kind 1: DMMA (mma.sync.aligned.m8n8k4.f64).  Runs `reps` timed launches of a grid that fills every SM
// (8 resident 256-thread CTAs per SM) on `stream`, SYNCHRONISES, and returns the best launch in TFLOP/s.
extern "C" int oiva_fp64_peak(int kind, int iters, int reps, double* tflops, void* stream) {
    using namespace oiva;
    OIVA_REQUIRE(tflops && (kind == 0 || kind == 1) && iters > 0 && reps > 0, "oiva_fp64_peak: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    OIVA_CUDA_CHECK(cudaGetDevice(&dev));
    OIVA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * 8, threads = 256;
    double* sink = nullptr;
    OIVA_CUDA_CHECK(cudaMalloc(&sink, (size_t)grid * threads * sizeof(double)));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    OIVA_CUDA_CHECK(cudaEventCreate(&e0));
    OIVA_CUDA_CHECK(cudaEventCreate(&e1));
    double best_ms = 1e30;
    for (int r = 0; r < reps + 1; ++r) {  // launch 0 is the warm-up
        cudaEventRecord(e0, st);
        if (kind == 0) k_peak_dfma<<<grid, threads, 0, st>>>(sink, iters, 1e-9, 0.999999);
        else k_peak_dmma<<<grid, threads, 0, st>>>(sink, iters, 1e-9, 0.999999);
        cudaEventRecord(e1, st);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaFree(sink);
            oiva_set_error("oiva_fp64_peak: %s", cudaGetErrorString(e));
            return OIVA_ERR_CUDA;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    // DFMA: 2 flops per lane-FMA; DMMA m8n8k4: 8*8*4 multiply-adds = 512 flops per warp instruction
    const double per_thread = kind == 0 ? 2.0 * PEAK_CHAINS * 4.0 : 512.0 / 32.0 * PEAK_CHAINS * 4.0;
    *tflops = per_thread * iters * (double)grid * threads / (best_ms * 1e-3) / 1e12;
    return OIVA_OK;
}
