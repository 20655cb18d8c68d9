// K3 / K5 -- filter functions from control matrices and the infidelity integral.
//
// Replaces numeric.calculate_filter_function (numeric.py:1413-1467),
// numeric.calculate_pulse_correlation_filter_function (numeric.py:1821-1883) and the integrand +
// trapezoid of numeric.infidelity (numeric.py:2318-2320, :259-374, util.py:880-906).
//
// These stages are HBM/L2-bound streaming reductions: omega is the fastest axis of every array, so a
// warp reads 32 consecutive complex128 values (512 B) per row and the basis axis is reduced in
// registers.
#include <algorithm>
#include <cstdlib>

#include "ffb_common.cuh"
#include "ffb_peer.cuh"

namespace {

// F[(l), (r), w] = sum_k conj(B[l,k,w]) B[r,k,w];  l, r run over (pulse, noise operator) rows.
// Output index (P = n_pulses, l = (g,a), r = (h,b)): (((g*P + h)*n_nops + a)*n_nops + b)*n_omega + w
template <int RT>
__global__ void __launch_bounds__(256)
ff_fidelity_kernel(int P, int n_nops, int n_basis, int n_omega, size_t ld,
                   const double2* __restrict__ B, double2* __restrict__ F) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_omega) return;
  const int L = P * n_nops;
  const int l = blockIdx.y;
  const int r0 = blockIdx.z * RT;
  double2 acc[RT];
#pragma unroll
  for (int i = 0; i < RT; ++i) acc[i] = make_double2(0.0, 0.0);
  const double2* Bl = B + (size_t)l * n_basis * ld + w;  // rows are ld frequencies apart
  for (int k = 0; k < n_basis; ++k) {
    const double2 x = Bl[(size_t)k * ld];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      const int r = r0 + i;
      if (r < L) {
        const double2 y = B[((size_t)r * n_basis + k) * ld + w];
        // conj(x) * y
        acc[i].x += x.x * y.x + x.y * y.y;
        acc[i].y += x.x * y.y - x.y * y.x;
      }
    }
  }
  const int g = l / n_nops, a = l % n_nops;
#pragma unroll
  for (int i = 0; i < RT; ++i) {
    const int r = r0 + i;
    if (r < L) {
      const int h = r / n_nops, b = r % n_nops;
      F[((((size_t)g * P + h) * n_nops + a) * n_nops + b) * ld + w] = acc[i];
    }
  }
}

// F[(l), (r), k, m, w] = conj(B[l,k,w]) B[r,m,w]   (write-bound outer product)
__global__ void __launch_bounds__(256)
ff_generalized_kernel(int P, int n_nops, int n_basis, int n_omega, const double2* __restrict__ B,
                      double2* __restrict__ F) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_omega) return;
  const int l = blockIdx.y / n_basis, k = blockIdx.y % n_basis;
  const int r = blockIdx.z;
  const double2 x = B[((size_t)l * n_basis + k) * n_omega + w];
  const int g = l / n_nops, a = l % n_nops;
  const int h = r / n_nops, b = r % n_nops;
  double2* dst = F + ((((((size_t)g * P + h) * n_nops + a) * n_nops + b) * n_basis + k) * n_basis) *
                         (size_t)n_omega + w;
  const double2* Br = B + (size_t)r * n_basis * n_omega + w;
  for (int m = 0; m < n_basis; ++m) {
    const double2 y = Br[(size_t)m * n_omega];
    dst[(size_t)m * n_omega] = make_double2(x.x * y.x + x.y * y.y, x.x * y.y - x.y * y.x);
  }
}

// One block per output element: trapezoid of Re(F S) over omega, warp-shuffle + shared reduction in a
// fixed order (deterministic).
__global__ void __launch_bounds__(1024)
infidelity_kernel(int n_nops, int n_sel, const int* __restrict__ idx, int n_omega,
                  const double2* __restrict__ F, const double* __restrict__ spectrum,
                  int spectrum_ndim, int spectrum_is_complex, const double* __restrict__ omega,
                  double norm, double* __restrict__ out, PeerReduce peers) {
  __shared__ double warp_sums[32];
  const int o = blockIdx.x;  // output index
  int lead, a, b;
  if (spectrum_ndim == 3) {
    lead = o / (n_sel * n_sel);
    a = (o / n_sel) % n_sel;
    b = o % n_sel;
  } else {
    lead = o / n_sel;
    a = b = o % n_sel;
  }
  const int ia = idx ? idx[a] : a, ib = idx ? idx[b] : b;  // idx == nullptr: all operators, in order
  const double2* Frow = F + (((size_t)lead * n_nops + ia) * n_nops + ib) * n_omega;
  size_t s_off = 0;
  if (spectrum_ndim == 2) s_off = (size_t)a * n_omega;
  if (spectrum_ndim == 3) s_off = ((size_t)a * n_sel + b) * n_omega;
  auto integrand = [&](int i) -> double {
    const double2 f = Frow[i];
    if (spectrum_is_complex) {
      const double2 s = reinterpret_cast<const double2*>(spectrum)[s_off + i];
      return f.x * s.x - f.y * s.y;
    }
    return f.x * spectrum[s_off + i];
  };
  double sum = 0.0;
  for (int i = threadIdx.x; i < n_omega - 1; i += blockDim.x) {
    sum += (integrand(i + 1) + integrand(i)) * (omega[i + 1] - omega[i]);
  }
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total += warp_sums[i];
    // frequency-sharded run: the partial integral of this rank's block is exchanged with the peers
    // over NVLink right here (peers.world == 1 otherwise)
    out[o] = peer_allreduce_slot(peers, o, total / 2.0 / norm);
  }
}

// Few outputs on a long frequency grid (one pulse: 3-36 integrals over 1e4-5e4 frequencies): one block
// per output leaves the machine idle and walks the grid with one block's worth of loads in flight
// (9 us for config 2).  Here an output is cut into chunks of INFID_CHUNK intervals, one block each; a
// block writes its partial sum and takes a ticket, and the block that draws the last ticket adds the
// partial sums up IN CHUNK ORDER (deterministic) and resets the ticket counter for the next launch.
constexpr int INFID_CHUNK = 1024;
constexpr int INFID_MAX_CHUNKS = 64;
constexpr int INFID_MAX_OUT = 64;  // more outputs than this fill the machine with one block each
__global__ void __launch_bounds__(256)
infidelity_chunked_kernel(int n_nops, int n_sel, const int* __restrict__ idx, int n_omega,
                          const double2* __restrict__ F, const double* __restrict__ spectrum,
                          int spectrum_ndim, int spectrum_is_complex,
                          const double* __restrict__ omega, double norm, int n_chunks,
                          double* __restrict__ partial, unsigned* __restrict__ tickets,
                          double* __restrict__ out, PeerReduce peers) {
  __shared__ double warp_sums[8];
  __shared__ bool last;
  const int o = blockIdx.x, chunk = blockIdx.y;
  int lead, a, b;
  if (spectrum_ndim == 3) {
    lead = o / (n_sel * n_sel);
    a = (o / n_sel) % n_sel;
    b = o % n_sel;
  } else {
    lead = o / n_sel;
    a = b = o % n_sel;
  }
  const int ia = idx ? idx[a] : a, ib = idx ? idx[b] : b;
  const double2* Frow = F + (((size_t)lead * n_nops + ia) * n_nops + ib) * n_omega;
  size_t s_off = 0;
  if (spectrum_ndim == 2) s_off = (size_t)a * n_omega;
  if (spectrum_ndim == 3) s_off = ((size_t)a * n_sel + b) * n_omega;
  auto integrand = [&](int i) -> double {
    const double2 f = Frow[i];
    if (spectrum_is_complex) {
      const double2 s = reinterpret_cast<const double2*>(spectrum)[s_off + i];
      return f.x * s.x - f.y * s.y;
    }
    return f.x * spectrum[s_off + i];
  };
  const int per = (n_omega - 1 + n_chunks - 1) / n_chunks;  // intervals per chunk
  const int i0 = chunk * per, i1 = min(n_omega - 1, i0 + per);
  double sum = 0.0;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x)
    sum += (integrand(i + 1) + integrand(i)) * (omega[i + 1] - omega[i]);
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total += warp_sums[i];
    partial[(size_t)o * n_chunks + chunk] = total;
    __threadfence();
    last = atomicAdd(&tickets[o], 1u) == (unsigned)(n_chunks - 1);
    if (last) {
      __threadfence();
      double all = 0.0;
      for (int c = 0; c < n_chunks; ++c)
        all += reinterpret_cast<volatile double*>(partial)[(size_t)o * n_chunks + c];
      tickets[o] = 0;  // ready for the next launch (launches on one stream are ordered)
      out[o] = peer_allreduce_slot(peers, o, all / 2.0 / norm);  // sum over the frequency shards
    }
  }
}

}  // namespace

int ffbi_filter_function(ffb_ctx* ctx, int P, int n_nops, int n_basis, int n_omega,
                         const double* B, int generalized, double* F) {
  FFB_REQUIRE(ctx, P >= 1 && n_nops >= 1 && n_basis >= 1 && n_omega >= 1,
              "filter function: bad shape (P=%d, n_nops=%d, n_basis=%d, n_omega=%d)", P, n_nops,
              n_basis, n_omega);
  const int L = P * n_nops;
  FFB_REQUIRE(ctx, L <= 65535 && (long long)L * n_basis <= 65535 * 1LL,
              "filter function: too many rows (%d x %d)", L, n_basis);
  const int wt = ceil_div(n_omega, 256);
  if (!generalized) {
    constexpr int RT = 4;
    dim3 grid(wt, L, ceil_div(L, RT));
    ff_fidelity_kernel<RT><<<grid, 256, 0, ctx->stream>>>(
        P, n_nops, n_basis, n_omega, (size_t)n_omega, reinterpret_cast<const double2*>(B),
        reinterpret_cast<double2*>(F));
  } else {
    dim3 grid(wt, L * n_basis, L);
    ff_generalized_kernel<<<grid, 256, 0, ctx->stream>>>(
        P, n_nops, n_basis, n_omega, reinterpret_cast<const double2*>(B),
        reinterpret_cast<double2*>(F));
  }
  FFB_LAUNCHED(ctx);
  return FFB_OK;
}

int ffbi_filter_function_ld(ffb_ctx* ctx, int P, int n_nops, int n_basis, int n_omega, size_t ld,
                            const double* B, double* F) {
  FFB_REQUIRE(ctx, P >= 1 && n_nops >= 1 && n_basis >= 1 && n_omega >= 1 && ld >= (size_t)n_omega,
              "filter function: bad shape (P=%d, n_nops=%d, n_basis=%d, n_omega=%d)", P, n_nops,
              n_basis, n_omega);
  const int L = P * n_nops;
  FFB_REQUIRE(ctx, L <= 65535, "filter function: too many rows (%d)", L);
  constexpr int RT = 4;
  dim3 grid(ceil_div(n_omega, 256), L, ceil_div(L, RT));
  ff_fidelity_kernel<RT><<<grid, 256, 0, ctx->stream>>>(
      P, n_nops, n_basis, n_omega, ld, reinterpret_cast<const double2*>(B),
      reinterpret_cast<double2*>(F));
  FFB_LAUNCHED(ctx);
  return FFB_OK;
}

int ffbi_infidelity(ffb_ctx* ctx, int n_lead, int n_nops, int n_sel, const int* idx_dev,
                    int n_omega, const double* F, const double* spectrum, int spectrum_ndim,
                    int spectrum_is_complex, const double* omega, int d, double* out) {
  FFB_REQUIRE(ctx, spectrum_ndim >= 1 && spectrum_ndim <= 3, "infidelity: spectrum_ndim=%d",
              spectrum_ndim);
  FFB_REQUIRE(ctx, n_lead >= 1 && n_sel >= 1 && n_omega >= 1 && d >= 1,
              "infidelity: bad shape (n_lead=%d, n_sel=%d, n_omega=%d, d=%d)", n_lead, n_sel,
              n_omega, d);
  const int n_out = n_lead * (spectrum_ndim == 3 ? n_sel * n_sel : n_sel);
  const double norm = 2.0 * 3.141592653589793238462643383279502884 * d;
  // frequency-sharded run (ffb_comm_reduce_infidelity): the exchange is the epilogue of the kernel when
  // the outputs fit one exchange window, a stand-alone all-reduce behind it otherwise
  const bool reduce = ctx->comm.reduce_infidelity && ctx->comm.connected && ctx->comm.world > 1;
  const bool fused = reduce && n_out <= FFB_PEER_SLOTS;
  static const bool chunked_ok = !(getenv("FFB_INFIDELITY_CHUNKED") && atoi(getenv("FFB_INFIDELITY_CHUNKED")) == 0);
  if (chunked_ok && n_out <= INFID_MAX_OUT && n_omega - 1 >= 2 * INFID_CHUNK) {
    const int n_chunks = std::min(INFID_MAX_CHUNKS, ceil_div(n_omega - 1, INFID_CHUNK));
    if (!ctx->infid_scratch) {  // partial sums and ticket counters, allocated (and zeroed) once
      FFB_CUDA(ctx, cudaMalloc(&ctx->infid_scratch,
                               (size_t)INFID_MAX_OUT * (INFID_MAX_CHUNKS * sizeof(double) + sizeof(unsigned))));
      FFB_CUDA(ctx, cudaMemsetAsync(ctx->infid_scratch, 0,
                                    (size_t)INFID_MAX_OUT * (INFID_MAX_CHUNKS * sizeof(double) + sizeof(unsigned)),
                                    ctx->stream));
    }
    double* partial = static_cast<double*>(ctx->infid_scratch);
    unsigned* tickets = reinterpret_cast<unsigned*>(partial + (size_t)INFID_MAX_OUT * INFID_MAX_CHUNKS);
    infidelity_chunked_kernel<<<dim3(n_out, n_chunks), 256, 0, ctx->stream>>>(
        n_nops, n_sel, idx_dev, n_omega, reinterpret_cast<const double2*>(F), spectrum, spectrum_ndim,
        spectrum_is_complex, omega, norm, n_chunks, partial, tickets, out,
        fused ? ffbi_peer_next(ctx) : PeerReduce());
    FFB_LAUNCHED(ctx);
    return FFB_OK;
  }
  const int threads = n_omega > 4096 ? 1024 : 256;
  infidelity_kernel<<<n_out, threads, 0, ctx->stream>>>(
      n_nops, n_sel, idx_dev, n_omega, reinterpret_cast<const double2*>(F), spectrum,
      spectrum_ndim, spectrum_is_complex, omega, norm, out, fused ? ffbi_peer_next(ctx) : PeerReduce());
  FFB_LAUNCHED(ctx);
  if (reduce && !fused) FFB_TRY(ffbi_allreduce_sum(ctx, out, n_out));
  return FFB_OK;
}
