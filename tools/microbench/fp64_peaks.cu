// FP64 micro-benchmarks for B200 (sm_100a): the roofline denominators MEASURED_PEAKS.json lacks.
//
//   dfma      : dependent-chain-free DFMA stream              -> vector FP64 peak
//   dmma      : mma.sync.m8n8k4.f64 stream                    -> FP64 tensor peak
//   mixed r   : 1 DMMA per r DFMA in the same warp            -> do the two pipes overlap?
//   sincos_*  : CUDA sincos() on small / large arguments, and the 3-constant Cody-Waite variant
//               the control-matrix kernel uses
//   rcp       : 1.0/x (IEEE) and the MUFU.RCP64H + 2 Newton steps used in the kernel
//
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peaks fp64_peaks.cu
// Every number is printed as one JSON line; timings are CUDA events around `reps` launches.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// The inner loop of the control-matrix kernel without the operand generator: A fragments come from
// shared memory (one LDS.64 per row tile), two accumulator tiles (Re/Im) per row tile and NT frequency
// tiles per warp.  Does the LDS-fed DMMA stream reach the register-resident peak?
template <int MT, int NT>
__global__ void k_dmma_lds(double* out, double b0, double b1, int units) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * MT * 32; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  __syncthreads();
  double cre[NT][MT][2], cim[NT][MT][2];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int m = 0; m < MT; ++m) { cre[n][m][0] = cre[n][m][1] = 0.0; cim[n][m][0] = cim[n][m][1] = 0.0; }
  for (int u = 0; u < units; ++u) {
    const double* up = sm + (u & 1) * MT * 32;
    const double br = b0 + u, bi = b1 - u;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const double a = up[m * 32 + lane];
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        dmma884(cre[n][m][0], cre[n][m][1], a, br + n);
        dmma884(cim[n][m][0], cim[n][m][1], a, bi + n);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int m = 0; m < MT; ++m) s += cre[n][m][0] + cre[n][m][1] + cim[n][m][0] + cim[n][m][1];
  if (s == 123.456) out[0] = s;
}

// R DFMA per DMMA, both streams independent.
template <int R>
__global__ void k_mixed(double* out, double a, double b) {
  double c0[8], c1[8], acc[8 * R];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
#pragma unroll
  for (int i = 0; i < 8 * R; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma884(c0[i], c1[i], a, b);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[i * R + r] = fma(acc[i * R + r], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < 8 * R; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__global__ void k_sincos(double* out, double x0, double dx) {
  double x = x0 + dx * (threadIdx.x + blockIdx.x * blockDim.x);
  double s = 0;
  for (int it = 0; it < ITERS / 16; ++it) {
    double sn, cs;
    sincos(x, &sn, &cs);
    s += sn * cs;
    x += dx;
  }
  if (s == 123.456) out[0] = s;
}

// 3-constant Cody-Waite reduction (fma keeps the products exact enough for |x| < 2^30) followed by
// the usual minimax kernels on [-pi/4, pi/4]. Absolute error ~1e-16, which is what a unit-modulus
// phase factor needs.
__device__ __forceinline__ void sincos_cw(double x, double* sn, double* cs) {
  const double two_over_pi = 6.36619772367581382433e-01;
  const double p1 = 1.57079632679489655800e+00, p2 = 6.12323399573676603587e-17,
               p3 = -1.49738490485916983294e-33;
  double k = rint(x * two_over_pi);
  double r = fma(-k, p1, x);
  r = fma(-k, p2, r);
  r = fma(-k, p3, r);
  int q = (int)(long long)k;
  double r2 = r * r;
  double ps = fma(r2, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(ps, r2, 2.75573137070700676789e-06);
  ps = fma(ps, r2, -1.98412698298579493134e-04);
  ps = fma(ps, r2, 8.33333333332248946124e-03);
  ps = fma(ps, r2, -1.66666666666666324348e-01);
  double s = fma(r * r2, ps, r);
  double pc = fma(r2, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(pc, r2, -2.75573143513906633035e-07);
  pc = fma(pc, r2, 2.48015872894767294178e-05);
  pc = fma(pc, r2, -1.38888888888741095749e-03);
  pc = fma(pc, r2, 4.16666666666666019037e-02);
  double c = fma(r2 * r2, pc, fma(r2, -0.5, 1.0));
  double ss = (q & 1) ? c : s;
  double cc = (q & 1) ? s : c;
  *sn = (q & 2) ? -ss : ss;
  *cs = ((q + 1) & 2) ? -cc : cc;
}

__global__ void k_sincos_cw(double* out, double x0, double dx) {
  double x = x0 + dx * (threadIdx.x + blockIdx.x * blockDim.x);
  double s = 0;
  for (int it = 0; it < ITERS / 16; ++it) {
    double sn, cs;
    sincos_cw(x, &sn, &cs);
    s += sn * cs;
    x += dx;
  }
  if (s == 123.456) out[0] = s;
}

__global__ void k_div(double* out, double x0, double dx) {
  double x = x0 + dx * (threadIdx.x + blockIdx.x * blockDim.x);
  double s = 0;
  for (int it = 0; it < ITERS / 4; ++it) {
    s += 1.0 / x;
    x += dx;
  }
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ double rcp_nr(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

__global__ void k_rcp_nr(double* out, double x0, double dx) {
  double x = x0 + dx * (threadIdx.x + blockIdx.x * blockDim.x);
  double s = 0;
  for (int it = 0; it < ITERS / 4; ++it) {
    s += rcp_nr(x);
    x += dx;
  }
  if (s == 123.456) out[0] = s;
}

// accuracy probe of sincos_cw against sincos() on a spread of arguments
__global__ void k_sincos_err(double* maxerr, double x0, double dx, int n) {
  int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i >= n) return;
  double x = x0 + dx * i;
  double s0, c0, s1, c1;
  sincos(x, &s0, &c0);
  sincos_cw(x, &s1, &c1);
  double e = fmax(fabs(s0 - s1), fabs(c0 - c1));
  unsigned long long* p = (unsigned long long*)maxerr;
  atomicMax(p, (unsigned long long)__double_as_longlong(e));
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out;
  CK(cudaMalloc(&out, 64));
  CK(cudaMemset(out, 0, 64));
  printf("{\"bench\":\"device\",\"name\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, sms, p.clockRate);

  const int reps = 20;
  for (int threads : {128, 256, 512}) {
    for (int bps : {1, 2, 4}) {
      const int grid = sms * bps;
      double ms = time_ms([&] { k_dfma<<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps);
      double fl = 2.0 * 16 * ITERS * (double)grid * threads;
      printf("{\"bench\":\"dfma\",\"threads\":%d,\"blocks_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n",
             threads, bps, ms, fl / ms * 1e-9);
    }
  }
  for (int threads : {128, 256, 512}) {
    for (int bps : {1, 2, 4}) {
      const int grid = sms * bps;
      double ms = time_ms([&] { k_dmma<8><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps);
      double fl = 2.0 * 256 * 8 * ITERS * (double)grid * (threads / 32);
      printf("{\"bench\":\"dmma8\",\"threads\":%d,\"blocks_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n",
             threads, bps, ms, fl / ms * 1e-9);
      ms = time_ms([&] { k_dmma<24><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps);
      fl = 2.0 * 256 * 24 * ITERS * (double)grid * (threads / 32);
      printf("{\"bench\":\"dmma24\",\"threads\":%d,\"blocks_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n",
             threads, bps, ms, fl / ms * 1e-9);
    }
  }
  {
    const int units = 4096;
    auto run = [&](auto kern, int mt, int nt, int threads, int bps, const char* name) {
      const int grid = sms * bps;
      const size_t smem = 2 * mt * 32 * sizeof(double);
      double ms = time_ms([&] { kern<<<grid, threads, smem>>>(out, 1.0000001, 1e-9, units); }, reps);
      double fl = 2.0 * 256 * 2 * mt * nt * units * (double)grid * (threads / 32);
      printf("{\"bench\":\"%s\",\"MT\":%d,\"NT\":%d,\"threads\":%d,\"blocks_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n",
             name, mt, nt, threads, bps, ms, fl / ms * 1e-9);
    };
    run(k_dmma_lds<12, 1>, 12, 1, 128, 1, "dmma_lds");
    run(k_dmma_lds<12, 1>, 12, 1, 128, 2, "dmma_lds");
    run(k_dmma_lds<12, 1>, 12, 1, 256, 1, "dmma_lds");
    run(k_dmma_lds<6, 2>, 6, 2, 128, 2, "dmma_lds");
    run(k_dmma_lds<6, 2>, 6, 2, 256, 1, "dmma_lds");
    run(k_dmma_lds<2, 1>, 2, 1, 256, 3, "dmma_lds");
    run(k_dmma_lds<2, 4>, 2, 4, 256, 3, "dmma_lds");
  }
  {
    const int threads = 256, grid = sms * 4;
    double ms = time_ms([&] { k_mixed<1><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps);
    double fl_mma = 2.0 * 256 * 8 * ITERS * (double)grid * (threads / 32);
    double fl_fma = 2.0 * 8 * 1 * ITERS * (double)grid * threads;
    printf("{\"bench\":\"mixed\",\"dfma_per_dmma\":1,\"ms\":%.4f,\"tflops_dmma\":%.3f,\"tflops_dfma\":%.3f}\n",
           ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
    ms = time_ms([&] { k_mixed<4><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps);
    fl_fma = 2.0 * 8 * 4 * ITERS * (double)grid * threads;
    printf("{\"bench\":\"mixed\",\"dfma_per_dmma\":4,\"ms\":%.4f,\"tflops_dmma\":%.3f,\"tflops_dfma\":%.3f}\n",
           ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
    ms = time_ms([&] { k_mixed<8><<<grid, threads>>>(out, 1.0000001, 1e-9); }, reps);
    fl_fma = 2.0 * 8 * 8 * ITERS * (double)grid * threads;
    printf("{\"bench\":\"mixed\",\"dfma_per_dmma\":8,\"ms\":%.4f,\"tflops_dmma\":%.3f,\"tflops_dfma\":%.3f}\n",
           ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
  }
  {
    const int threads = 256, grid = sms * 8;
    const double n = (double)grid * threads * (ITERS / 16);
    double ms = time_ms([&] { k_sincos<<<grid, threads>>>(out, 0.1, 1e-7); }, reps);
    printf("{\"bench\":\"sincos_small\",\"ms\":%.4f,\"gevals_per_s\":%.2f}\n", ms, n / ms * 1e-6);
    ms = time_ms([&] { k_sincos<<<grid, threads>>>(out, 5.0e4, 1e-3); }, reps);
    printf("{\"bench\":\"sincos_5e4\",\"ms\":%.4f,\"gevals_per_s\":%.2f}\n", ms, n / ms * 1e-6);
    ms = time_ms([&] { k_sincos<<<grid, threads>>>(out, 5.0e5, 1e-3); }, reps);
    printf("{\"bench\":\"sincos_5e5_slowpath\",\"ms\":%.4f,\"gevals_per_s\":%.2f}\n", ms, n / ms * 1e-6);
    ms = time_ms([&] { k_sincos_cw<<<grid, threads>>>(out, 5.0e5, 1e-3); }, reps);
    printf("{\"bench\":\"sincos_cw_5e5\",\"ms\":%.4f,\"gevals_per_s\":%.2f}\n", ms, n / ms * 1e-6);
    const double nd = (double)grid * threads * (ITERS / 4);
    ms = time_ms([&] { k_div<<<grid, threads>>>(out, 1.5, 1e-7); }, reps);
    printf("{\"bench\":\"div_ieee\",\"ms\":%.4f,\"gevals_per_s\":%.2f}\n", ms, nd / ms * 1e-6);
    ms = time_ms([&] { k_rcp_nr<<<grid, threads>>>(out, 1.5, 1e-7); }, reps);
    printf("{\"bench\":\"rcp_nr2\",\"ms\":%.4f,\"gevals_per_s\":%.2f}\n", ms, nd / ms * 1e-6);
  }
  {
    double* me;
    CK(cudaMalloc(&me, 8));
    for (double x0 : {0.0, 1.0e3, 1.0e5, 1.0e6, 1.0e7, -3.0e6}) {
      CK(cudaMemset(me, 0, 8));
      const int n = 1 << 22;
      k_sincos_err<<<n / 256, 256>>>(me, x0, 0.37, n);
      double h;
      CK(cudaMemcpy(&h, me, 8, cudaMemcpyDeviceToHost));
      printf("{\"bench\":\"sincos_cw_maxabs_err\",\"x0\":%.1e,\"err\":%.3e}\n", x0, h);
    }
  }
  return 0;
}
