// K4 -- control matrix of a pulse sequence from the cached control matrices of its constituents,
// plus the two small helpers the sequencing layer needs on every call (Liouville representation of
// the cumulative propagators, phase factors).
//
// Replaces numeric.calculate_control_matrix_from_atomic (numeric.py:621-704):
//     B(w) = sum_g phi_{g-1}(w) B^{(g)}(w) Q^{(g-1)},   phi_{-1} = 1, Q^{(-1)} = identity
// (the reference loops over g and does a (n_omega, n_basis) x (n_basis, n_basis) matmul per noise
// operator with omega moved to the second-to-last axis, numeric.py:679, :691-701),
// superoperator.liouville_representation (superoperator.py:51-84) and util.cexp (util.py:136-162).
//
// Mapping: omega stays the fastest axis (coalesced 512 B per warp and row); a thread owns one
// frequency, one noise operator and LT output basis columns, with the LT x n_basis slice of Q staged in
// shared memory per constituent pulse.
#include "ffb_common.cuh"

namespace {

template <int LT, bool QC>
__global__ void __launch_bounds__(256)
from_atomic_kernel(int P, int n_nops, int n_basis, int n_omega, const double2* __restrict__ phases,
                   const double2* __restrict__ Bat, const double* __restrict__ Q, int correlations,
                   double2* __restrict__ out) {
  extern __shared__ double qs[];  // [n_basis][LT] (x2 if complex)
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  const int l0 = blockIdx.z * LT;
  const bool active = w < n_omega;
  const int QW = QC ? 2 : 1;
  double2 acc[LT];
#pragma unroll
  for (int i = 0; i < LT; ++i) acc[i] = make_double2(0.0, 0.0);
  const size_t pulse_stride = (size_t)n_nops * n_basis * n_omega;

  for (int g = 0; g < P; ++g) {
    const double2* Bg = Bat + (size_t)g * pulse_stride + (size_t)j * n_basis * n_omega;
    if (g == 0) {
      if (active) {
#pragma unroll
        for (int i = 0; i < LT; ++i) {
          if (l0 + i < n_basis) acc[i] = Bg[(size_t)(l0 + i) * n_omega + w];
        }
      }
    } else {
      __syncthreads();
      const double* Qg = Q + (size_t)(g - 1) * n_basis * n_basis * QW;
      for (int e = threadIdx.x; e < n_basis * LT; e += blockDim.x) {
        const int k = e / LT, i = e % LT;
        if (QC) {
          qs[2 * e] = l0 + i < n_basis ? Qg[2 * ((size_t)k * n_basis + l0 + i)] : 0.0;
          qs[2 * e + 1] = l0 + i < n_basis ? Qg[2 * ((size_t)k * n_basis + l0 + i) + 1] : 0.0;
        } else {
          qs[e] = l0 + i < n_basis ? Qg[(size_t)k * n_basis + l0 + i] : 0.0;
        }
      }
      __syncthreads();
      if (active) {
        if (correlations) {
#pragma unroll
          for (int i = 0; i < LT; ++i) acc[i] = make_double2(0.0, 0.0);
        }
        const double2 ph = phases[(size_t)(g - 1) * n_omega + w];
        for (int k = 0; k < n_basis; ++k) {
          const double2 b = Bg[(size_t)k * n_omega + w];
          const double xr = ph.x * b.x - ph.y * b.y;
          const double xi = ph.x * b.y + ph.y * b.x;
#pragma unroll
          for (int i = 0; i < LT; ++i) {
            if (QC) {
              const double qr = qs[2 * (k * LT + i)], qi = qs[2 * (k * LT + i) + 1];
              acc[i].x += xr * qr - xi * qi;
              acc[i].y += xr * qi + xi * qr;
            } else {
              const double qv = qs[k * LT + i];
              acc[i].x += xr * qv;
              acc[i].y += xi * qv;
            }
          }
        }
      }
    }
    if (correlations && active) {
      double2* dst = out + (size_t)g * pulse_stride + (size_t)j * n_basis * n_omega;
#pragma unroll
      for (int i = 0; i < LT; ++i) {
        if (l0 + i < n_basis) dst[(size_t)(l0 + i) * n_omega + w] = acc[i];
      }
    }
  }
  if (!correlations && active) {
    double2* dst = out + (size_t)j * n_basis * n_omega;
#pragma unroll
    for (int i = 0; i < LT; ++i) {
      if (l0 + i < n_basis) dst[(size_t)(l0 + i) * n_omega + w] = acc[i];
    }
  }
}

// out[n,i,j] = tr(C_i U_n C_j U_n^dagger); one block per (n, j)
__global__ void __launch_bounds__(128)
liouville_kernel(int d, int n_basis, const double* __restrict__ U, const double* __restrict__ basis,
                 double* __restrict__ out) {
  extern __shared__ double sm[];
  const int n = blockIdx.y, j = blockIdx.x;
  const int dd = d * d;
  double* Us = sm;           // d*d complex
  double* T = Us + 2 * dd;   // C_j U^+
  double* W = T + 2 * dd;    // U C_j U^+
  const double* Un = U + (size_t)n * 2 * dd;
  const double* Cj = basis + (size_t)j * 2 * dd;
  for (int e = threadIdx.x; e < 2 * dd; e += blockDim.x) Us[e] = Un[e];
  __syncthreads();
  for (int e = threadIdx.x; e < dd; e += blockDim.x) {
    const int a = e / d, b = e % d;
    cplx acc = {0.0, 0.0};
    for (int c = 0; c < d; ++c) {  // T[a][b] = sum_c C_j[a][c] conj(U[b][c])
      const cplx cj = {Cj[2 * (a * d + c)], Cj[2 * (a * d + c) + 1]};
      const cplx u = {Us[2 * (b * d + c)], Us[2 * (b * d + c) + 1]};
      acc = cadd(acc, cmulc(cj, u));
    }
    T[2 * e] = acc.re;
    T[2 * e + 1] = acc.im;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < dd; e += blockDim.x) {
    const int a = e / d, b = e % d;
    cplx acc = {0.0, 0.0};
    for (int c = 0; c < d; ++c) {
      const cplx u = {Us[2 * (a * d + c)], Us[2 * (a * d + c) + 1]};
      const cplx t = {T[2 * (c * d + b)], T[2 * (c * d + b) + 1]};
      acc = cadd(acc, cmul(u, t));
    }
    W[2 * e] = acc.re;
    W[2 * e + 1] = acc.im;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_basis; i += blockDim.x) {
    const double* Ci = basis + (size_t)i * 2 * dd;
    cplx acc = {0.0, 0.0};
    for (int a = 0; a < d; ++a) {
      for (int b = 0; b < d; ++b) {  // tr(C_i W) = sum_ab C_i[a][b] W[b][a]
        const cplx ci = {Ci[2 * (a * d + b)], Ci[2 * (a * d + b) + 1]};
        const cplx wv = {W[2 * (b * d + a)], W[2 * (b * d + a) + 1]};
        acc = cadd(acc, cmul(ci, wv));
      }
    }
    double* dst = out + 2 * (((size_t)n * n_basis + i) * n_basis + j);
    dst[0] = acc.re;
    dst[1] = acc.im;
  }
}

__global__ void cexp_kernel(int n, const double* __restrict__ x, double scale,
                            double2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sn, cs;
  sincos(x[i] * scale, &sn, &cs);
  out[i] = make_double2(cs, sn);
}

template <int LT>
int launch_from_atomic(ffb_ctx* ctx, int P, int n_nops, int n_basis, int n_omega,
                       const double* phases, const double* B_atomic, const double* Q,
                       int q_is_complex, int correlations, double* out) {
  dim3 grid(ceil_div(n_omega, 256), n_nops, ceil_div(n_basis, LT));
  const size_t smem = (size_t)n_basis * LT * sizeof(double) * (q_is_complex ? 2 : 1);
  if (q_is_complex) {
    auto kern = from_atomic_kernel<LT, true>;
    FFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, ctx->stream>>>(P, n_nops, n_basis, n_omega,
                                           reinterpret_cast<const double2*>(phases),
                                           reinterpret_cast<const double2*>(B_atomic), Q,
                                           correlations, reinterpret_cast<double2*>(out));
  } else {
    auto kern = from_atomic_kernel<LT, false>;
    FFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, ctx->stream>>>(P, n_nops, n_basis, n_omega,
                                           reinterpret_cast<const double2*>(phases),
                                           reinterpret_cast<const double2*>(B_atomic), Q,
                                           correlations, reinterpret_cast<double2*>(out));
  }
  FFB_LAUNCHED(ctx);
  return FFB_OK;
}

}  // namespace

int ffbi_from_atomic(ffb_ctx* ctx, int P, int n_nops, int n_basis, int n_omega,
                     const double* phases, const double* B_atomic, const double* Q,
                     int q_is_complex, int correlations, double* out) {
  FFB_REQUIRE(ctx, P >= 1 && n_nops >= 1 && n_basis >= 1 && n_omega >= 1,
              "from_atomic: bad shape (P=%d, n_nops=%d, n_basis=%d, n_omega=%d)", P, n_nops,
              n_basis, n_omega);
  FFB_REQUIRE(ctx, n_nops <= 65535, "from_atomic: too many noise operators (%d)", n_nops);
  if (n_basis <= 4)
    return launch_from_atomic<4>(ctx, P, n_nops, n_basis, n_omega, phases, B_atomic, Q,
                                 q_is_complex, correlations, out);
  if (n_basis <= 8)
    return launch_from_atomic<8>(ctx, P, n_nops, n_basis, n_omega, phases, B_atomic, Q,
                                 q_is_complex, correlations, out);
  return launch_from_atomic<16>(ctx, P, n_nops, n_basis, n_omega, phases, B_atomic, Q,
                                q_is_complex, correlations, out);
}

int ffbi_liouville(ffb_ctx* ctx, int n, int d, int n_basis, const double* U, const double* basis,
                   double* out) {
  FFB_REQUIRE(ctx, n >= 1 && d >= 1 && n_basis >= 1 && n <= 65535,
              "liouville: bad shape (n=%d, d=%d, n_basis=%d)", n, d, n_basis);
  dim3 grid(n_basis, n);
  const size_t smem = (size_t)6 * d * d * sizeof(double);
  FFB_CUDA(ctx, cudaFuncSetAttribute(liouville_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  liouville_kernel<<<grid, 128, smem, ctx->stream>>>(d, n_basis, U, basis, out);
  FFB_LAUNCHED(ctx);
  return FFB_OK;
}

int ffbi_cexp(ffb_ctx* ctx, int n, const double* x, double scale, double* out) {
  FFB_REQUIRE(ctx, n >= 1, "cexp: n=%d", n);
  cexp_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(n, x, scale,
                                                         reinterpret_cast<double2*>(out));
  FFB_LAUNCHED(ctx);
  return FFB_OK;
}
