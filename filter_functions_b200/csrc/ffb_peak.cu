// FP64 roofline denominators measured on the device at hand: a register-resident DFMA stream and a
// DMMA.8x8x4 stream (the same kernels as tools/microbench/fp64_peaks.cu, best of 5 launches).
// MEASURED_PEAKS.json only holds HBM and bf16 numbers, so bench.py asks the library for these.
#include "ffb_common.cuh"

namespace {

constexpr int PEAK_ITERS = 2048;

__global__ void __launch_bounds__(512) peak_dfma_kernel(double* out, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(512) peak_dmma_kernel(double* out, double a, double b) {
  double c0[16], c1[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    c0[i] = threadIdx.x + i;
    c1[i] = i;
  }
  for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i])
                   : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

}  // namespace

int ffbi_fp64_peak(ffb_ctx* ctx, double* dfma, double* dmma) {
  DevBuf out;
  FFB_TRY(out.alloc(ctx, 64));
  cudaEvent_t e0, e1;
  FFB_CUDA(ctx, cudaEventCreate(&e0));
  FFB_CUDA(ctx, cudaEventCreate(&e1));
  const int grid = ctx->sm_count * 4, threads = 512;
  double best[2] = {0.0, 0.0};
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 7; ++rep) {
      FFB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
      if (which == 0)
        peak_dfma_kernel<<<grid, threads, 0, ctx->stream>>>(out.as<double>(), 1.0000001, 1e-9);
      else
        peak_dmma_kernel<<<grid, threads, 0, ctx->stream>>>(out.as<double>(), 1.0000001, 1e-9);
      FFB_LAUNCHED(ctx);
      FFB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
      FFB_CUDA(ctx, cudaEventSynchronize(e1));
      float ms = 0.f;
      FFB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
      const double flops = which == 0 ? 2.0 * 16 * PEAK_ITERS * (double)grid * threads
                                      : 2.0 * 256 * 16 * PEAK_ITERS * (double)grid * (threads / 32);
      if (rep >= 2) best[which] = std::max(best[which], flops / (ms * 1e-3) * 1e-12);
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (dfma) *dfma = best[0];
  if (dmma) *dmma = best[1];
  return FFB_OK;
}
