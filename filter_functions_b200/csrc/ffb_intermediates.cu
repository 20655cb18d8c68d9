// K7 -- materialising variant of the control-matrix calculation: cache_intermediates=True of
// numeric.calculate_control_matrix_from_scratch (numeric.py:828-879 of the reference).
//
// The fused kernel of ffb_ctrlmat.cu exists to AVOID the (G, n_omega, d, d)-sized arrays; the
// reference's gradient and second-order code consume exactly those arrays (gradient.py:477-501,
// numeric.py:1598-1612), so this file writes them, with the reference's own rounding sequence:
//   eigvecs_propagated[g]      = Q_{g}^+ V_g                                  (numeric.py:93-95)
//   n_opers_transformed[j,g]   = s_j^{(g)} V_g^+ B_j V_g                      (numeric.py:98-123)
//   basis_transformed[g,k]     = U^+ C_k U, U = eigvecs_propagated[g]         (numeric.py:126-141)
//   phase_factors[g,w]         = exp(i w t_g)                                 (numeric.py:865)
//   first_order_integral[g,w,m,n] = (exp(i x dt) - 1) / (i x), x = w + (E_m - E_n); dt if x == 0
//                                                                               (numeric.py:144-167)
//   control_matrix_step[g,j,k,w]  = phase sum_mn Bbar_j[m,n] I[m,n] Cbar_k[n,m] (numeric.py:867)
//   control_matrix_step_cumulative[g-1] = sum_{g' < g} step[g']               (numeric.py:859-860)
// and the control matrix as the running sum of the steps, as the reference accumulates it.
// These are store-bound streaming kernels (omega fastest everywhere except first_order_integral, whose
// reference layout is (G, n_omega, d, d)).
#include <algorithm>

#include "ffb_common.cuh"

namespace {

// one block per segment: U = Q_g^+ V_g, then all operators through shared memory
__global__ void __launch_bounds__(128)
transform_raw_kernel(int G, int d, int n_nops, int n_basis, const double* __restrict__ eigvecs,
                     const double* __restrict__ propagators, const double* __restrict__ n_opers,
                     const double* __restrict__ n_coeffs, const double* __restrict__ basis,
                     double* __restrict__ eigvecs_propagated, double* __restrict__ n_opers_transformed,
                     double* __restrict__ basis_transformed) {
  extern __shared__ double sm[];
  const int g = blockIdx.x;
  const int dd = d * d;
  double* V = sm;
  double* U = V + 2 * dd;
  double* T = U + 2 * dd;
  const double* Vg = eigvecs + (size_t)g * 2 * dd;
  const double* Qg = propagators + (size_t)g * 2 * dd;
  for (int e = threadIdx.x; e < 2 * dd; e += blockDim.x) V[e] = Vg[e];
  __syncthreads();
  for (int e = threadIdx.x; e < dd; e += blockDim.x) {
    const int a = e / d, b = e % d;
    cplx acc = {0.0, 0.0};
    for (int c = 0; c < d; ++c) {  // U[a][b] = sum_c conj(Q[c][a]) V[c][b]
      const cplx qv = {Qg[2 * (c * d + a)], -Qg[2 * (c * d + a) + 1]};
      const cplx vv = {V[2 * (c * d + b)], V[2 * (c * d + b) + 1]};
      acc = cadd(acc, cmul(qv, vv));
    }
    U[2 * e] = acc.re;
    U[2 * e + 1] = acc.im;
    eigvecs_propagated[(size_t)g * 2 * dd + 2 * e] = acc.re;
    eigvecs_propagated[(size_t)g * 2 * dd + 2 * e + 1] = acc.im;
  }
  __syncthreads();
  for (int op = 0; op < n_nops + n_basis; ++op) {
    const bool is_noise = op < n_nops;
    const double* O = is_noise ? n_opers + (size_t)op * 2 * dd : basis + (size_t)(op - n_nops) * 2 * dd;
    const double* W = is_noise ? V : U;
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {  // T = O W
      const int a = e / d, b = e % d;
      cplx acc = {0.0, 0.0};
      for (int c = 0; c < d; ++c) {
        const cplx o = {O[2 * (a * d + c)], O[2 * (a * d + c) + 1]};
        const cplx w = {W[2 * (c * d + b)], W[2 * (c * d + b) + 1]};
        acc = cadd(acc, cmul(o, w));
      }
      T[2 * e] = acc.re;
      T[2 * e + 1] = acc.im;
    }
    __syncthreads();
    const double scale = is_noise ? n_coeffs[(size_t)op * G + g] : 1.0;
    double* dst = is_noise ? n_opers_transformed + ((size_t)op * G + g) * 2 * dd
                           : basis_transformed + ((size_t)g * n_basis + (op - n_nops)) * 2 * dd;
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {  // X = W^+ T
      const int a = e / d, b = e % d;
      cplx acc = {0.0, 0.0};
      for (int c = 0; c < d; ++c) {
        const cplx w = {W[2 * (c * d + a)], -W[2 * (c * d + a) + 1]};
        const cplx t = {T[2 * (c * d + b)], T[2 * (c * d + b) + 1]};
        acc = cadd(acc, cmul(w, t));
      }
      dst[2 * e] = acc.re * scale;
      dst[2 * e + 1] = acc.im * scale;
    }
    __syncthreads();
  }
}

// phase_factors[g,w] and first_order_integral[g,w,m,n]; grid (omega tiles, G)
__global__ void __launch_bounds__(128)
integral_kernel(int g0, int d, int n_omega, const double* __restrict__ eigvals,
                const double* __restrict__ omega, const double* __restrict__ dt,
                const double* __restrict__ t, double2* __restrict__ phase_factors,
                double2* __restrict__ integral) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = g0 + blockIdx.y;   // the segment axis is walked in slices of <= 65535 (grid limit)
  if (w >= n_omega) return;
  const double om = omega[w], dtg = dt[g];
  double sn, cs;
  sincos(__dmul_rn(om, t[g]), &sn, &cs);
  phase_factors[(size_t)g * n_omega + w] = make_double2(cs, sn);
  double2* dst = integral + ((size_t)g * n_omega + w) * d * d;
  const double* E = eigvals + (size_t)g * d;
  for (int m = 0; m < d; ++m) {
    for (int n = 0; n < d; ++n) {
      // x = w + (E_m - E_n), y = x dt, each individually rounded as in numeric.py:156-158
      const double x = __dadd_rn(om, __dsub_rn(E[m], E[n]));
      double2 val;
      if (x == 0.0) {
        val = make_double2(dtg, 0.0);
      } else {
        const double y = __dmul_rn(x, dtg);
        double s2, c2, s1, c1;
        sincos(0.5 * y, &s2, &c2);
        sincos(y, &s1, &c1);
        // exp(i y) - 1 = -2 sin^2(y/2) + i sin(y)   (util.cexpm1, util.py:165-182); divided by i x
        const double re = -2.0 * s2 * s2, im = s1;
        val = make_double2(im / x, -re / x);
      }
      dst[m * d + n] = val;
    }
  }
}

// control_matrix_step[g,j,k,w]; grid (omega tiles, n_nops * n_basis, G)
__global__ void __launch_bounds__(128)
step_kernel(int g0, int G, int d, int n_nops, int n_basis, int n_omega,
            const double2* __restrict__ n_opers_transformed,
            const double2* __restrict__ basis_transformed, const double2* __restrict__ phase_factors,
            const double2* __restrict__ integral, double2* __restrict__ step) {
  extern __shared__ double2 coef[];  // M[m][n] = Bbar_j[m][n] * Cbar_k[n][m]
  const int g = g0 + blockIdx.z;
  const int j = blockIdx.y / n_basis, k = blockIdx.y % n_basis;
  const int dd = d * d;
  const double2* Bb = n_opers_transformed + ((size_t)j * G + g) * dd;
  const double2* Cb = basis_transformed + ((size_t)g * n_basis + k) * dd;
  for (int e = threadIdx.x; e < dd; e += blockDim.x) {
    const int m = e / d, n = e % d;
    const double2 b = Bb[m * d + n], c = Cb[n * d + m];
    coef[e] = make_double2(b.x * c.x - b.y * c.y, b.x * c.y + b.y * c.x);
  }
  __syncthreads();
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_omega) return;
  const double2* I = integral + ((size_t)g * n_omega + w) * dd;
  double re = 0.0, im = 0.0;
  for (int e = 0; e < dd; ++e) {
    const double2 x = I[e], c = coef[e];
    re += c.x * x.x - c.y * x.y;
    im += c.x * x.y + c.y * x.x;
  }
  const double2 ph = phase_factors[(size_t)g * n_omega + w];
  step[(((size_t)g * n_nops + j) * n_basis + k) * n_omega + w] =
      make_double2(ph.x * re - ph.y * im, ph.x * im + ph.y * re);
}

// cumulative[g-1] = sum_{g'<g} step[g'], out = sum of all steps; one thread per (j,k,w)
__global__ void __launch_bounds__(256)
cumulate_kernel(int G, size_t per_step, const double2* __restrict__ step,
                double2* __restrict__ cumulative, double2* __restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per_step) return;
  double re = 0.0, im = 0.0;
  for (int g = 0; g < G; ++g) {
    if (g > 0) cumulative[(size_t)(g - 1) * per_step + e] = make_double2(re, im);
    const double2 s = step[(size_t)g * per_step + e];
    re += s.x;
    im += s.y;
  }
  out[e] = make_double2(re, im);
}

}  // namespace

int ffbi_control_matrix_intermediates(ffb_ctx* ctx, int G, int d, int n_nops, int n_basis,
                                      int n_omega, const double* eigvals, const double* eigvecs,
                                      const double* propagators, const double* omega,
                                      const double* basis, const double* n_opers,
                                      const double* n_coeffs, const double* dt, const double* t,
                                      double* out, double* n_opers_transformed,
                                      double* eigvecs_propagated, double* basis_transformed,
                                      double* phase_factors, double* first_order_integral,
                                      double* step, double* cumulative) {
  FFB_CHECK_DIM(ctx, d);
  FFB_REQUIRE(ctx, G >= 1 && d >= 1 && d <= 32 && n_nops >= 1 && n_basis >= 1 && n_omega >= 1,
              "control matrix intermediates: bad shape");
  FFB_REQUIRE(ctx, (long long)n_nops * n_basis <= 65535,
              "control matrix intermediates: n_nops*n_basis=%d exceeds 65535", n_nops * n_basis);
  const int dd = d * d;
  {
    const size_t smem = (size_t)6 * dd * sizeof(double);
    FFB_CUDA(ctx, cudaFuncSetAttribute(transform_raw_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    transform_raw_kernel<<<G, 128, smem, ctx->stream>>>(G, d, n_nops, n_basis, eigvecs, propagators,
                                                        n_opers, n_coeffs, basis, eigvecs_propagated,
                                                        n_opers_transformed, basis_transformed);
    FFB_LAUNCHED(ctx);
  }
  for (int g0 = 0; g0 < G; g0 += 65535) {
    const int gn = std::min(65535, G - g0);
    dim3 grid(ceil_div(n_omega, 128), gn);
    integral_kernel<<<grid, 128, 0, ctx->stream>>>(g0, d, n_omega, eigvals, omega, dt, t,
                                                   reinterpret_cast<double2*>(phase_factors),
                                                   reinterpret_cast<double2*>(first_order_integral));
    FFB_LAUNCHED(ctx);
  }
  for (int g0 = 0; g0 < G; g0 += 65535) {
    const int gn = std::min(65535, G - g0);
    dim3 grid(ceil_div(n_omega, 128), n_nops * n_basis, gn);
    step_kernel<<<grid, 128, (size_t)dd * 16, ctx->stream>>>(
        g0, G, d, n_nops, n_basis, n_omega, reinterpret_cast<const double2*>(n_opers_transformed),
        reinterpret_cast<const double2*>(basis_transformed),
        reinterpret_cast<const double2*>(phase_factors),
        reinterpret_cast<const double2*>(first_order_integral), reinterpret_cast<double2*>(step));
    FFB_LAUNCHED(ctx);
  }
  {
    const size_t per_step = (size_t)n_nops * n_basis * n_omega;
    cumulate_kernel<<<(unsigned)ceil_div_sz(per_step, 256), 256, 0, ctx->stream>>>(
        G, per_step, reinterpret_cast<const double2*>(step), reinterpret_cast<double2*>(cumulative),
        reinterpret_cast<double2*>(out));
    FFB_LAUNCHED(ctx);
  }
  return FFB_OK;
}
