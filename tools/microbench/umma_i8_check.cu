// Bring-up check for the int8 tcgen05 path (sm_100a): ONE CTA computes D[M=128, N] += A[128, K] * B[N, K]^T
// with tcgen05.mma.kind::i8 from shared-memory operands in the canonical K-major no-swizzle layout,
// accumulators in TMEM, read back with tcgen05.ld, and compares with a CPU loop.  Exercises: descriptor
// encoding (LBO / SBO), instruction descriptor (u8 / s8 formats, M, N), accumulate flag over several
// K = 32 steps, tcgen05.commit -> mbarrier, TMEM alloc / dealloc.
//
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o umma_i8_check umma_i8_check.cu && ./umma_i8_check
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                          \
  do {                                                                                    \
    cudaError_t e = (x);                                                                  \
    if (e != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);      \
      exit(1);                                                                            \
    }                                                                                     \
  } while (0)

constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// K-major, no swizzle: 16-byte unit (row r, k-chunk c) at  base + c * LBO + (r / 8) * SBO + (r % 8) * 16
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, layout type 0 (no swizzle)
}

__device__ __forceinline__ uint32_t make_idesc(int n, bool a_signed, bool b_signed) {
  uint32_t d = 0;
  d |= 2u << 4;                       // accumulator format S32
  d |= (a_signed ? 1u : 0u) << 7;     // A format: 0 = u8, 1 = s8
  d |= (b_signed ? 1u : 0u) << 10;    // B format
  // bits 15 / 16: A, B K-major = 0
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

template <int N>
__global__ void __launch_bounds__(128) umma_check_kernel(const int8_t* __restrict__ A,   // [KS][128][32]
                                                         const int8_t* __restrict__ B,   // [KS][N][32]
                                                         int32_t* __restrict__ D,        // [128][N]
                                                         int ksteps, int a_signed, int b_signed) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t mbar;
  uint8_t* sA = smem;                       // per k-step: 2 chunks x 16 row groups x 128 B = 4096 B
  uint8_t* sB = smem + (size_t)ksteps * 4096;  // per k-step: 2 chunks x (N/8) row groups x 128 B
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int B_STEP = 2 * (N / 8) * 128;

  // operands -> canonical layout: [k-step][chunk c][row group][row in group][16 bytes]
  for (int i = tid; i < ksteps * M * 2; i += 128) {
    const int ks = i / (M * 2), r = (i / 2) % M, c = i % 2;
    const uint4 v = *reinterpret_cast<const uint4*>(A + ((size_t)ks * M + r) * 32 + c * 16);
    *reinterpret_cast<uint4*>(sA + (size_t)ks * 4096 + c * 2048 + (r / 8) * 128 + (r % 8) * 16) = v;
  }
  for (int i = tid; i < ksteps * N * 2; i += 128) {
    const int ks = i / (N * 2), r = (i / 2) % N, c = i % 2;
    const uint4 v = *reinterpret_cast<const uint4*>(B + ((size_t)ks * N + r) * 32 + c * 16);
    *reinterpret_cast<uint4*>(sB + (size_t)ks * B_STEP + c * (N / 8) * 128 + (r / 8) * 128 + (r % 8) * 16) = v;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;

  if (tid == 0) {
    const uint32_t idesc = make_idesc(N, a_signed != 0, b_signed != 0);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t da = make_desc(smem_u32(sA + (size_t)ks * 4096), 2048, 128);
      const uint64_t db = make_desc(smem_u32(sB + (size_t)ks * B_STEP), (N / 8) * 128, 128);
      const uint32_t accumulate = ks > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
  }
  // wait for the MMAs (phase 0)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // TMEM -> registers: warp w reads lanes 32 w .. 32 w + 31 (rows of D), 32 columns at a time
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[(size_t)row * N + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

template <int N>
int run(int ksteps, int a_signed, int b_signed) {
  std::vector<int8_t> A((size_t)ksteps * M * 32), B((size_t)ksteps * N * 32);
  uint32_t s = 12345u + N + 7 * ksteps + a_signed * 3 + b_signed;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int)((s >> 16) & 0xFF); };
  for (auto& x : A) x = (int8_t)rnd();
  for (auto& x : B) x = (int8_t)rnd();
  std::vector<int32_t> ref((size_t)M * N, 0), got((size_t)M * N, -1);
  for (int ks = 0; ks < ksteps; ++ks)
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        int32_t acc = 0;
        for (int k = 0; k < 32; ++k) {
          const int a = a_signed ? (int)A[((size_t)ks * M + m) * 32 + k] : (int)(uint8_t)A[((size_t)ks * M + m) * 32 + k];
          const int b = b_signed ? (int)B[((size_t)ks * N + n) * 32 + k] : (int)(uint8_t)B[((size_t)ks * N + n) * 32 + k];
          acc += a * b;
        }
        ref[(size_t)m * N + n] += acc;
      }
  int8_t *dA, *dB;
  int32_t* dD;
  CHECK(cudaMalloc(&dA, A.size()));
  CHECK(cudaMalloc(&dB, B.size()));
  CHECK(cudaMalloc(&dD, got.size() * 4));
  CHECK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CHECK(cudaMemset(dD, 0xFF, got.size() * 4));
  const size_t smem = (size_t)ksteps * (4096 + 2 * (N / 8) * 128);
  CHECK(cudaFuncSetAttribute(umma_check_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_check_kernel<N><<<1, 128, smem>>>(dA, dB, dD, ksteps, a_signed, b_signed);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t i = 0; i < got.size(); ++i) bad += got[i] != ref[i];
  printf("N=%3d ksteps=%d a_%s b_%s : %zu / %zu mismatches (D[0][0] = %d, ref %d; D[127][N-1] = %d, ref %d)\n",
         N, ksteps, a_signed ? "s8" : "u8", b_signed ? "s8" : "u8", bad, got.size(), got[0], ref[0],
         got.back(), ref.back());
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return bad != 0;
}

int main() {
  int fails = 0;
  fails += run<96>(1, 1, 1);
  fails += run<96>(4, 1, 1);
  fails += run<96>(3, 0, 1);
  fails += run<96>(3, 1, 0);
  fails += run<96>(2, 0, 0);
  fails += run<128>(2, 1, 1);
  fails += run<64>(2, 1, 0);
  printf(fails ? "UMMA_I8_CHECK FAILED (%d)\n" : "UMMA_I8_CHECK OK\n", fails);
  return fails;
}
