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

// exp(i x) - 1 = -2 sin^2(x/2) + i sin(x) (util.cexpm1, util.py:165-182: no cancellation for small x)
__global__ void cexpm1_kernel(int n, const double* __restrict__ x, double2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s = sin(x[i] / 2);
  out[i] = make_double2(-2.0 * (s * s), sin(x[i]));
}

// ------------------------------------------------------------------------------------------------
// Batched concatenation: n_seq gate sequences drawn from a library of n_lib
// pulses with cached control matrices (randomized benchmarking, examples/randomized_benchmarking.py:
// 70-91 of the reference calls ff.concatenate once per sequence).  seq_scan_kernel forms the running
// products the reference builds with util.adot / util.mdot (pulse_sequence.py:1745, :1827) for every
// sequence at once; sequence_batch_kernel is from_atomic_kernel with the constituent's control matrix
// gathered from the library and the cumulative phase factors (np.cumprod, pulse_sequence.py:1824) kept
// as a running product in registers.  A negative library index is padding (no gate).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
seq_scan_kernel(int L, int n_basis, int d, const int* __restrict__ idx,
                const double* __restrict__ lib_liouville, const double2* __restrict__ lib_U,
                double* __restrict__ Qcum, double2* __restrict__ U_total) {
  extern __shared__ double sm[];
  const int s = blockIdx.x;
  const int nn = n_basis * n_basis, dd = d * d;
  double* R0 = sm;                 // running Liouville product (ping-pong)
  double* R1 = R0 + nn;
  double2* V0 = reinterpret_cast<double2*>(R1 + nn);  // running propagator (ping-pong)
  double2* V1 = V0 + dd;
  const int* seq = idx + (size_t)s * L;
  for (int e = threadIdx.x; e < nn; e += blockDim.x) R0[e] = (e / n_basis == e % n_basis) ? 1.0 : 0.0;
  for (int e = threadIdx.x; e < dd; e += blockDim.x)
    V0[e] = make_double2((e / d == e % d) ? 1.0 : 0.0, 0.0);
  __syncthreads();
  for (int g = 0; g < L; ++g) {
    const int c = seq[g];
    if (c >= 0) {
      const double* Lc = lib_liouville + (size_t)c * nn;
      for (int e = threadIdx.x; e < nn; e += blockDim.x) {  // R1 = L_c R0
        const int i = e / n_basis, j = e % n_basis;
        double acc = 0.0;
        for (int k = 0; k < n_basis; ++k) acc += Lc[i * n_basis + k] * R0[k * n_basis + j];
        R1[e] = acc;
      }
      const double2* Uc = lib_U + (size_t)c * dd;
      for (int e = threadIdx.x; e < dd; e += blockDim.x) {  // V1 = U_c V0
        const int i = e / d, j = e % d;
        double re = 0.0, im = 0.0;
        for (int k = 0; k < d; ++k) {
          const double2 a = Uc[i * d + k], b = V0[k * d + j];
          re += a.x * b.x - a.y * b.y;
          im += a.x * b.y + a.y * b.x;
        }
        V1[e] = make_double2(re, im);
      }
      __syncthreads();
      double* tr = R0; R0 = R1; R1 = tr;
      double2* tv = V0; V0 = V1; V1 = tv;
    }
    if (g + 1 < L) {
      double* dst = Qcum + ((size_t)s * (L - 1) + g) * nn;
      for (int e = threadIdx.x; e < nn; e += blockDim.x) dst[e] = R0[e];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < dd; e += blockDim.x) U_total[(size_t)s * dd + e] = V0[e];
}

template <int LT>
__global__ void __launch_bounds__(128)
sequence_batch_kernel(int L, int n_nops, int n_basis, int n_omega, const int* __restrict__ idx,
                      const double2* __restrict__ lib_B, const double2* __restrict__ lib_phase,
                      const double* __restrict__ Qcum, double2* __restrict__ out_B) {
  extern __shared__ double qs[];  // [n_basis][LT]
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y / n_nops, j = blockIdx.y % n_nops;
  const int l0 = blockIdx.z * LT;
  const bool active = w < n_omega;
  const int* seq = idx + (size_t)s * L;
  const size_t lib_stride = (size_t)n_nops * n_basis * n_omega;
  double2 acc[LT];
#pragma unroll
  for (int i = 0; i < LT; ++i) acc[i] = make_double2(0.0, 0.0);
  double2 ph = make_double2(1.0, 0.0);  // product of the total phases of the gates so far
  bool first = true;                    // no gate applied yet: Q is the identity
  for (int g = 0; g < L; ++g) {
    const int c = seq[g];
    if (c < 0) continue;  // uniform over the block
    const double2* Bg = lib_B + (size_t)c * lib_stride + (size_t)j * n_basis * n_omega + w;
    if (first) {
      if (active) {
#pragma unroll
        for (int i = 0; i < LT; ++i)
          if (l0 + i < n_basis) acc[i] = Bg[(size_t)(l0 + i) * n_omega];
      }
      first = false;
    } else {
      __syncthreads();
      const double* Qg = Qcum + ((size_t)s * (L - 1) + g - 1) * n_basis * n_basis;
      for (int e = threadIdx.x; e < n_basis * LT; e += blockDim.x) {
        const int k = e / LT, i = e % LT;
        qs[e] = l0 + i < n_basis ? Qg[(size_t)k * n_basis + l0 + i] : 0.0;
      }
      __syncthreads();
      if (active) {
        for (int k = 0; k < n_basis; ++k) {
          const double2 b = Bg[(size_t)k * n_omega];
          const double xr = ph.x * b.x - ph.y * b.y;
          const double xi = ph.x * b.y + ph.y * b.x;
#pragma unroll
          for (int i = 0; i < LT; ++i) {
            const double qv = qs[k * LT + i];
            acc[i].x += xr * qv;
            acc[i].y += xi * qv;
          }
        }
      }
    }
    if (active) {
      const double2 p2 = lib_phase[(size_t)c * n_omega + w];
      ph = make_double2(ph.x * p2.x - ph.y * p2.y, ph.x * p2.y + ph.y * p2.x);
    }
  }
  if (active) {
    double2* dst = out_B + ((size_t)s * n_nops + j) * n_basis * n_omega + w;
#pragma unroll
    for (int i = 0; i < LT; ++i)
      if (l0 + i < n_basis) dst[(size_t)(l0 + i) * n_omega] = acc[i];
  }
}

// F[s,a,b,w] = sum_k conj(B[s,a,k,w]) B[s,b,k,w]   (one sequence per blockIdx.y)
__global__ void __launch_bounds__(128)
ff_batched_kernel(int n_nops, int n_basis, int n_omega, const double2* __restrict__ B,
                  double2* __restrict__ F) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_omega) return;
  const int s = blockIdx.y;
  const int a = blockIdx.z / n_nops, b = blockIdx.z % n_nops;
  const double2* Ba = B + ((size_t)s * n_nops + a) * n_basis * n_omega + w;
  const double2* Bb = B + ((size_t)s * n_nops + b) * n_basis * n_omega + w;
  double re = 0.0, im = 0.0;
  for (int k = 0; k < n_basis; ++k) {
    const double2 x = Ba[(size_t)k * n_omega], y = Bb[(size_t)k * n_omega];
    re += x.x * y.x + x.y * y.y;
    im += __dsub_rn(__dmul_rn(x.x, y.y), __dmul_rn(x.y, y.x));   // exactly antisymmetric in (a, b), see ffb_filter.cu
  }
  F[(((size_t)s * n_nops + a) * n_nops + b) * n_omega + w] = make_double2(re, im);
}

// ------------------------------------------------------------------------------------------------
// Periodic repetition: numeric.calculate_control_matrix_periodic
// (numeric.py:884-954), B_tot(w) = B(w) sum_{g<G} (phi(w) L)^g.  The reference solves a d^2 x d^2 linear
// system per frequency and falls back to the explicit sum where I - phi L is ill-conditioned.  Here the
// geometric series is built by binary doubling ON THE ROW VECTORS u_m = B S_m, S_m = sum_{g<m} (phi L)^g:
//     u_{2m}  = u_m + phi^m (u_m L^m)        u_{m+1} = B + phi (u_m L)
// (S_m and L commute), i.e. ~2 log2(G) products of the frequency-dependent rows with FIXED matrices L^m
// -- no per-frequency inverse, no conditioning check, O(n_nops n_basis^2 log G) per frequency.
//     out[j,l,w] = A[j,l,w] + phi(w)^m sum_k X[j,k,w] Q[k,l]
// ------------------------------------------------------------------------------------------------
template <int LT, bool QC>
__global__ void __launch_bounds__(256)
periodic_step_kernel(int n_basis, int n_omega, int m, const double2* __restrict__ phases,
                     const double2* __restrict__ A, const double2* __restrict__ X,
                     const double* __restrict__ Q, double2* __restrict__ out) {
  extern __shared__ double qs[];  // [n_basis][LT] (x2 if complex)
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  const int l0 = blockIdx.z * LT;
  for (int e = threadIdx.x; e < n_basis * LT; e += blockDim.x) {
    const int k = e / LT, i = e % LT;
    const bool ok = l0 + i < n_basis;
    if (QC) {
      qs[2 * e] = ok ? Q[2 * ((size_t)k * n_basis + l0 + i)] : 0.0;
      qs[2 * e + 1] = ok ? Q[2 * ((size_t)k * n_basis + l0 + i) + 1] : 0.0;
    } else {
      qs[e] = ok ? Q[(size_t)k * n_basis + l0 + i] : 0.0;
    }
  }
  __syncthreads();
  if (w >= n_omega) return;
  // phi^m by binary powering of the GIVEN phase factor (as matrix_power does in the reference)
  double2 ph = make_double2(1.0, 0.0), base = phases[w];
  for (int e = m; e > 0; e >>= 1) {
    if (e & 1) ph = make_double2(ph.x * base.x - ph.y * base.y, ph.x * base.y + ph.y * base.x);
    base = make_double2(base.x * base.x - base.y * base.y, 2.0 * base.x * base.y);
  }
  const size_t row0 = (size_t)j * n_basis * n_omega + w;
  double2 acc[LT];
#pragma unroll
  for (int i = 0; i < LT; ++i) acc[i] = make_double2(0.0, 0.0);
  for (int k = 0; k < n_basis; ++k) {
    const double2 x = X[row0 + (size_t)k * n_omega];
#pragma unroll
    for (int i = 0; i < LT; ++i) {
      if (QC) {
        const double qr = qs[2 * (k * LT + i)], qi = qs[2 * (k * LT + i) + 1];
        acc[i].x += x.x * qr - x.y * qi;
        acc[i].y += x.x * qi + x.y * qr;
      } else {
        const double qv = qs[k * LT + i];
        acc[i].x += x.x * qv;
        acc[i].y += x.y * qv;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LT; ++i) {
    if (l0 + i < n_basis) {
      const size_t e = row0 + (size_t)(l0 + i) * n_omega;
      const double2 a = A[e];
      out[e] = make_double2(a.x + ph.x * acc[i].x - ph.y * acc[i].y,
                            a.y + ph.x * acc[i].y + ph.y * acc[i].x);
    }
  }
}

// C = A B for n x n matrices (real, or complex if QC); the fixed matrices L^m of the doubling scheme
template <bool QC>
__global__ void __launch_bounds__(256)
small_matmul_kernel(int n, const double* __restrict__ A, const double* __restrict__ B,
                    double* __restrict__ C) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * n) return;
  const int i = e / n, j = e % n;
  double re = 0.0, im = 0.0;
  for (int k = 0; k < n; ++k) {
    if (QC) {
      const double ar = A[2 * (i * n + k)], ai = A[2 * (i * n + k) + 1];
      const double br = B[2 * (k * n + j)], bi = B[2 * (k * n + j) + 1];
      re += ar * br - ai * bi;
      im += ar * bi + ai * br;
    } else {
      re += A[i * n + k] * B[k * n + j];
    }
  }
  if (QC) {
    C[2 * e] = re;
    C[2 * e + 1] = im;
  } else {
    C[e] = re;
  }
}

template <int LT>
int launch_from_atomic(ffb_ctx* ctx, int P, int n_nops, int n_rows, int n_basis, int n_omega,
                       const double* phases, const double* B_atomic, const double* Q,
                       int q_is_complex, int correlations, double* out) {
  dim3 grid(ceil_div(n_omega, 256), n_rows, ceil_div(n_basis, LT));
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

// Rows [j0, j0 + jn) of the noise-operator axis only (the strides stay those of the full arrays), so
// that a caller can pipeline the download of finished rows with the computation of the next ones.
int ffbi_from_atomic_rows(ffb_ctx* ctx, int P, int n_nops, int j0, int jn, int n_basis, int n_omega,
                          const double* phases, const double* B_atomic, const double* Q,
                          int q_is_complex, int correlations, double* out) {
  FFB_REQUIRE(ctx, P >= 1 && n_nops >= 1 && n_basis >= 1 && n_omega >= 1,
              "from_atomic: bad shape (P=%d, n_nops=%d, n_basis=%d, n_omega=%d)", P, n_nops,
              n_basis, n_omega);
  FFB_REQUIRE(ctx, n_nops <= 65535, "from_atomic: too many noise operators (%d)", n_nops);
  FFB_REQUIRE(ctx, j0 >= 0 && jn >= 1 && j0 + jn <= n_nops, "from_atomic: bad row range");
  const size_t off = (size_t)j0 * n_basis * n_omega * 2;
  B_atomic += off;
  out += off;
  // FP64-bound regime (n_basis / 4 flop per byte): DMMA kernel of ffb_atomic_dmma.cu
  if (!q_is_complex && n_basis >= 32)
    return ffbi_from_atomic_dmma(ctx, P, n_nops, jn, n_basis, n_omega, phases, B_atomic, Q,
                                 correlations, out);
  if (n_basis <= 4)
    return launch_from_atomic<4>(ctx, P, n_nops, jn, n_basis, n_omega, phases, B_atomic, Q,
                                 q_is_complex, correlations, out);
  if (n_basis <= 8)
    return launch_from_atomic<8>(ctx, P, n_nops, jn, n_basis, n_omega, phases, B_atomic, Q,
                                 q_is_complex, correlations, out);
  return launch_from_atomic<16>(ctx, P, n_nops, jn, n_basis, n_omega, phases, B_atomic, Q,
                                q_is_complex, correlations, out);
}

int ffbi_from_atomic(ffb_ctx* ctx, int P, int n_nops, int n_basis, int n_omega,
                     const double* phases, const double* B_atomic, const double* Q,
                     int q_is_complex, int correlations, double* out) {
  return ffbi_from_atomic_rows(ctx, P, n_nops, 0, n_nops, n_basis, n_omega, phases, B_atomic, Q,
                               q_is_complex, correlations, out);
}

int ffbi_liouville(ffb_ctx* ctx, int n, int d, int n_basis, const double* U, const double* basis,
                   double* out) {
  FFB_REQUIRE(ctx, n >= 1 && d >= 1 && n_basis >= 1 && n <= 65535,
              "liouville: bad shape (n=%d, d=%d, n_basis=%d)", n, d, n_basis);
  dim3 grid(n_basis, n);
  const size_t smem = (size_t)6 * d * d * sizeof(double);
  FFB_TRY(ffb_func_smem(ctx, liouville_kernel, smem));
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

int ffbi_cexpm1(ffb_ctx* ctx, int n, const double* x, double* out) {
  FFB_REQUIRE(ctx, n >= 1, "cexpm1: n=%d", n);
  cexpm1_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(n, x, reinterpret_cast<double2*>(out));
  FFB_LAUNCHED(ctx);
  return FFB_OK;
}

int ffbi_concatenate_many(ffb_ctx* ctx, int n_seq, int L, int d, int n_nops, int n_basis,
                          int n_omega, const int* idx, const double* lib_B, const double* lib_phase,
                          const double* lib_liouville, const double* lib_U, double* U_total,
                          double* out_B, double* out_F) {
  FFB_REQUIRE(ctx, n_seq >= 1 && L >= 1 && d >= 1 && n_nops >= 1 && n_basis >= 1 && n_omega >= 1,
              "concatenate_many: bad shape (n_seq=%d, L=%d, d=%d, n_nops=%d, n_basis=%d, n_omega=%d)",
              n_seq, L, d, n_nops, n_basis, n_omega);
  FFB_REQUIRE(ctx, n_basis <= 64, "concatenate_many: n_basis=%d > 64 is not supported "
              "(use concatenate per sequence)", n_basis);
  FFB_REQUIRE(ctx, (long long)n_seq * n_nops <= 65535, "concatenate_many: n_seq * n_nops = %lld "
              "exceeds 65535; split the batch", (long long)n_seq * n_nops);
  const int nn = n_basis * n_basis;
  DevBuf Qcum;
  FFB_TRY(Qcum.alloc(ctx, (size_t)n_seq * std::max(1, L - 1) * nn * 8));
  {
    const size_t smem = (size_t)2 * nn * 8 + (size_t)2 * d * d * 16;
    FFB_CUDA(ctx, cudaFuncSetAttribute(seq_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    seq_scan_kernel<<<n_seq, 128, smem, ctx->stream>>>(
        L, n_basis, d, idx, lib_liouville, reinterpret_cast<const double2*>(lib_U),
        Qcum.as<double>(), reinterpret_cast<double2*>(U_total));
    FFB_LAUNCHED(ctx);
  }
  {
    const int LT = n_basis <= 4 ? 4 : n_basis <= 8 ? 8 : 16;
    dim3 grid(ceil_div(n_omega, 128), n_seq * n_nops, ceil_div(n_basis, LT));
    const size_t smem = (size_t)n_basis * LT * 8;
    auto args = [&](auto kern) {
      kern<<<grid, 128, smem, ctx->stream>>>(L, n_nops, n_basis, n_omega, idx,
                                             reinterpret_cast<const double2*>(lib_B),
                                             reinterpret_cast<const double2*>(lib_phase),
                                             Qcum.as<double>(), reinterpret_cast<double2*>(out_B));
    };
    if (LT == 4) args(sequence_batch_kernel<4>);
    else if (LT == 8) args(sequence_batch_kernel<8>);
    else args(sequence_batch_kernel<16>);
    FFB_LAUNCHED(ctx);
  }
  if (out_F) {
    dim3 grid(ceil_div(n_omega, 128), n_seq, n_nops * n_nops);
    FFB_REQUIRE(ctx, n_nops * n_nops <= 65535 && n_seq <= 65535, "concatenate_many: grid too large");
    ff_batched_kernel<<<grid, 128, 0, ctx->stream>>>(n_nops, n_basis, n_omega,
                                                     reinterpret_cast<const double2*>(out_B),
                                                     reinterpret_cast<double2*>(out_F));
    FFB_LAUNCHED(ctx);
  }
  return FFB_OK;
}

namespace {

template <int LT>
int launch_periodic_step(ffb_ctx* ctx, int n_nops, int n_basis, int n_omega, int m,
                         const double* phases, const double* A, const double* X, const double* Q,
                         int q_is_complex, double* out) {
  dim3 grid(ceil_div(n_omega, 256), n_nops, ceil_div(n_basis, LT));
  const size_t smem = (size_t)n_basis * LT * sizeof(double) * (q_is_complex ? 2 : 1);
  auto launch = [&](auto kern) -> int {
    FFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, ctx->stream>>>(n_basis, n_omega, m,
                                           reinterpret_cast<const double2*>(phases),
                                           reinterpret_cast<const double2*>(A),
                                           reinterpret_cast<const double2*>(X), Q,
                                           reinterpret_cast<double2*>(out));
    FFB_LAUNCHED(ctx);
    return FFB_OK;
  };
  return q_is_complex ? launch(periodic_step_kernel<LT, true>)
                      : launch(periodic_step_kernel<LT, false>);
}

int periodic_step(ffb_ctx* ctx, int n_nops, int n_basis, int n_omega, int m, const double* phases,
                  const double* A, const double* X, const double* Q, int q_is_complex, double* out) {
  if (n_basis <= 4)
    return launch_periodic_step<4>(ctx, n_nops, n_basis, n_omega, m, phases, A, X, Q, q_is_complex, out);
  if (n_basis <= 8)
    return launch_periodic_step<8>(ctx, n_nops, n_basis, n_omega, m, phases, A, X, Q, q_is_complex, out);
  return launch_periodic_step<16>(ctx, n_nops, n_basis, n_omega, m, phases, A, X, Q, q_is_complex, out);
}

}  // namespace

int ffbi_control_matrix_periodic(ffb_ctx* ctx, int n_nops, int n_basis, int n_omega, int repeats,
                                 const double* phases, const double* B, const double* L,
                                 int l_is_complex, double* out) {
  FFB_REQUIRE(ctx, n_nops >= 1 && n_basis >= 1 && n_omega >= 1 && repeats >= 1 && n_nops <= 65535,
              "periodic control matrix: bad shape (n_nops=%d, n_basis=%d, n_omega=%d, repeats=%d)",
              n_nops, n_basis, n_omega, repeats);
  const size_t b_bytes = (size_t)n_nops * n_basis * n_omega * 16;
  const size_t l_bytes = (size_t)n_basis * n_basis * (l_is_complex ? 16 : 8);
  if (repeats == 1) {
    FFB_CUDA(ctx, cudaMemcpyAsync(out, B, b_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return FFB_OK;
  }
  DevBuf u0, u1, Lm0, Lm1;
  FFB_TRY(u0.alloc(ctx, b_bytes));
  FFB_TRY(u1.alloc(ctx, b_bytes));
  FFB_TRY(Lm0.alloc(ctx, l_bytes));
  FFB_TRY(Lm1.alloc(ctx, l_bytes));
  auto matmul = [&](const double* A_, const double* B_, double* C_) -> int {
    const unsigned blocks = (unsigned)ceil_div(n_basis * n_basis, 256);
    if (l_is_complex) small_matmul_kernel<true><<<blocks, 256, 0, ctx->stream>>>(n_basis, A_, B_, C_);
    else small_matmul_kernel<false><<<blocks, 256, 0, ctx->stream>>>(n_basis, A_, B_, C_);
    FFB_LAUNCHED(ctx);
    return FFB_OK;
  };
  // u = B S_m and Lm = L^m for m running through the binary prefixes of `repeats`
  const double* u = B;      // m = 1: S_1 = identity
  const double* Lm = L;
  double* u_next = u0.as<double>();
  double* L_next = Lm0.as<double>();
  auto flip_u = [&]() { u = u_next; u_next = (u_next == u0.as<double>()) ? u1.as<double>() : u0.as<double>(); };
  auto flip_L = [&]() { Lm = L_next; L_next = (L_next == Lm0.as<double>()) ? Lm1.as<double>() : Lm0.as<double>(); };
  int m = 1;
  int top = 30;
  while (!((repeats >> top) & 1)) --top;
  for (int bit = top - 1; bit >= 0; --bit) {
    // doubling: u_{2m} = u_m + phi^m (u_m L^m)
    FFB_TRY(periodic_step(ctx, n_nops, n_basis, n_omega, m, phases, u, u, Lm, l_is_complex, u_next));
    flip_u();
    FFB_TRY(matmul(Lm, Lm, L_next));
    flip_L();
    m *= 2;
    if ((repeats >> bit) & 1) {
      // increment: u_{m+1} = B + phi (u_m L)
      FFB_TRY(periodic_step(ctx, n_nops, n_basis, n_omega, 1, phases, B, u, L, l_is_complex, u_next));
      flip_u();
      FFB_TRY(matmul(Lm, L, L_next));
      flip_L();
      m += 1;
    }
  }
  FFB_CUDA(ctx, cudaMemcpyAsync(out, u, b_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return FFB_OK;
}
