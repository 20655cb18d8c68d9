// Peer group of the GPUs of one node (one process per GPU): exchange windows mapped with CUDA IPC,
// the stand-alone all-reduce kernel and the all-gather of frequency blocks.  The handles travel between
// the processes through whatever the host program uses for plumbing (torch.distributed in
// filter_functions_b200/distributed.py); no data of the path goes through the host or through a
// library collective.
//
// the frequency axis shards naturally, so the only exchanges on the path are the sum of
// the per-noise-operator partial integrals (fused into the infidelity kernels, ffb_filter.cu +
// ffb_peer.cuh) and, when the caller wants the whole F(omega) on every rank, the gather of its
// column blocks (below: every rank stores its block straight into all peers' windows over NVLink).
#include <algorithm>
#include <cstring>

#include "ffb_common.cuh"
#include "ffb_peer.cuh"

namespace {

constexpr size_t WINDOW_BYTES = (size_t)2 * FFB_MAX_PEERS * FFB_PEER_SLOTS * 16;

__global__ void peer_allreduce_kernel(PeerReduce pr, double* __restrict__ data, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) data[i] = peer_allreduce_slot(pr, i, data[i]);
}

// Every rank stores its (rows x count) block of complex128 columns into all windows at column `col0`
// of a (rows x ld) array: 16-byte stores, a warp writes 512 contiguous bytes of one row.
struct Windows {
  unsigned long long p[FFB_MAX_PEERS];
};

__global__ void __launch_bounds__(256)
peer_put_kernel(Windows dst, int world, int rows, int count, size_t ld, size_t col0,
                const double2* __restrict__ local) {
  const size_t total = (size_t)rows * count;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / count, c = i % count;
    const double2 v = local[i];
    const size_t off = (r * ld + col0 + c) * sizeof(double2);
#pragma unroll 1
    for (int p = 0; p < world; ++p) *reinterpret_cast<double2*>(dst.p[p] + off) = v;
  }
  __threadfence_system();
}

int require_group(ffb_ctx* ctx) {
  if (!ctx) return FFB_EINVAL;
  if (!ctx->comm.connected)
    return ffb_fail(ctx, FFB_EINVAL, "no peer group: call ffb_comm_create / ffb_comm_connect first");
  return FFB_OK;
}

}  // namespace

PeerReduce ffbi_peer_next(ffb_ctx* ctx) {
  PeerReduce pr;
  ffb_comm& c = ctx->comm;
  if (!c.connected || c.world <= 1) return pr;
  pr.rank = c.rank;
  pr.world = c.world;
  for (int r = 0; r < c.world; ++r) pr.window[r] = reinterpret_cast<unsigned long long>(c.peer[r]);
  if (++c.seq == 0) c.seq = 1;  // 0 is the value of an untouched slot
  pr.seq = c.seq;
  pr.timeout_ms = c.timeout_ms;
  pr.error = c.err_dev;
  return pr;
}

int ffbi_allreduce_sum(ffb_ctx* ctx, double* data_dev, int n) {
  if (!ctx->comm.connected || ctx->comm.world <= 1 || n <= 0) return FFB_OK;
  for (int i0 = 0; i0 < n; i0 += FFB_PEER_SLOTS) {
    const int m = std::min(FFB_PEER_SLOTS, n - i0);
    peer_allreduce_kernel<<<ceil_div(m, 128), 128, 0, ctx->stream>>>(ffbi_peer_next(ctx),
                                                                      data_dev + i0, m);
    FFB_LAUNCHED(ctx);
  }
  return FFB_OK;
}

int ffbi_comm_fetch_error(ffb_ctx* ctx) {
  if (!ctx->comm.err_dev) return FFB_OK;
  FFB_CUDA(ctx, cudaMemcpyAsync(ctx->comm.err_host, ctx->comm.err_dev, sizeof(int),
                                cudaMemcpyDeviceToHost, ctx->stream));
  return FFB_OK;
}

int ffbi_comm_check_error(ffb_ctx* ctx) {
  if (!ctx->comm.err_dev || *ctx->comm.err_host == 0) return FFB_OK;
  const int who = *ctx->comm.err_host - 1;
  *ctx->comm.err_host = 0;
  FFB_CUDA(ctx, cudaMemsetAsync(ctx->comm.err_dev, 0, sizeof(int), ctx->stream));
  return ffb_fail(ctx, FFB_ECUDA, "peer exchange timed out after %u ms waiting for rank %d (rank %d of %d); "
                  "do all ranks issue the same collectives?", ctx->comm.timeout_ms, who, ctx->comm.rank,
                  ctx->comm.world);
}

void ffbi_comm_release(ffb_ctx* ctx) {
  ffb_comm& c = ctx->comm;
  for (int r = 0; r < c.world && r < FFB_MAX_PEERS; ++r) {
    if (r != c.rank && c.peer[r]) cudaIpcCloseMemHandle(c.peer[r]);
    if (r != c.rank && c.peer_data[r]) cudaIpcCloseMemHandle(c.peer_data[r]);
    c.peer[r] = c.peer_data[r] = nullptr;
  }
  if (c.window) cudaFree(c.window);
  if (c.data) cudaFree(c.data);
  if (c.err_dev) cudaFree(c.err_dev);
  if (c.err_host) cudaFreeHost(c.err_host);
  c = ffb_comm();
  (void)cudaGetLastError();
}

extern "C" {

int ffb_comm_create(ffb_ctx* ctx, int rank, int world, void* handle_out) {
  if (!ctx || !handle_out) return FFB_EINVAL;
  FFB_CUDA(ctx, cudaSetDevice(ctx->device));
  FFB_REQUIRE(ctx, world >= 1 && world <= FFB_MAX_PEERS && rank >= 0 && rank < world,
              "peer group: rank %d of %d (at most %d ranks)", rank, world, FFB_MAX_PEERS);
  static_assert(sizeof(cudaIpcMemHandle_t) == FFB_COMM_HANDLE_BYTES, "handle size");
  FFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ffbi_comm_release(ctx);
  ffb_comm& c = ctx->comm;
  c.rank = rank;
  c.world = world;
  if (const char* e = getenv("FFB_PEER_TIMEOUT_MS")) c.timeout_ms = (unsigned)std::max(1, atoi(e));
  FFB_CUDA(ctx, cudaMalloc(&c.window, WINDOW_BYTES));
  FFB_CUDA(ctx, cudaMemset(c.window, 0, WINDOW_BYTES));
  FFB_CUDA(ctx, cudaMalloc(&c.err_dev, sizeof(int)));
  FFB_CUDA(ctx, cudaMemset(c.err_dev, 0, sizeof(int)));
  FFB_CUDA(ctx, cudaHostAlloc(&c.err_host, sizeof(int), cudaHostAllocDefault));
  *c.err_host = 0;
  cudaIpcMemHandle_t h;
  FFB_CUDA(ctx, cudaIpcGetMemHandle(&h, c.window));
  std::memcpy(handle_out, &h, sizeof(h));
  return FFB_OK;
}

int ffb_comm_connect(ffb_ctx* ctx, const void* handles) {
  if (!ctx || !handles) return FFB_EINVAL;
  FFB_CUDA(ctx, cudaSetDevice(ctx->device));
  ffb_comm& c = ctx->comm;
  FFB_REQUIRE(ctx, c.window, "peer group: ffb_comm_create has not been called");
  for (int r = 0; r < c.world; ++r) {
    if (r == c.rank) {
      c.peer[r] = c.window;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&c.peer[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      c.peer[r] = nullptr;
      return ffb_fail(ctx, FFB_ECUDA, "peer group: cannot map the window of rank %d into rank %d (%s); "
                      "no NVLink/P2P path or CUDA IPC not permitted", r, c.rank, cudaGetErrorString(e));
    }
  }
  c.connected = true;
  return FFB_OK;
}

int ffb_comm_destroy(ffb_ctx* ctx) {
  if (!ctx) return FFB_EINVAL;
  FFB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStreamSynchronize(ctx->stream);
  ffbi_comm_release(ctx);
  return FFB_OK;
}

int ffb_comm_info(const ffb_ctx* ctx, int* rank, int* world, size_t* data_window_bytes) {
  if (!ctx) return FFB_EINVAL;
  if (rank) *rank = ctx->comm.rank;
  if (world) *world = ctx->comm.connected ? ctx->comm.world : 1;
  if (data_window_bytes) *data_window_bytes = ctx->comm.data_bytes;
  return FFB_OK;
}

int ffb_comm_reduce_infidelity(ffb_ctx* ctx, int enable) {
  if (!ctx) return FFB_EINVAL;
  if (enable) FFB_TRY(require_group(ctx));
  ctx->comm.reduce_infidelity = enable != 0;
  return FFB_OK;
}

int ffb_dev_allreduce_sum(ffb_ctx* ctx, double* data, int n) {
  FFB_TRY(require_group(ctx));
  FFB_REQUIRE(ctx, data && n >= 0, "all-reduce: bad arguments");
  return ffbi_allreduce_sum(ctx, data, n);
}

int ffb_allreduce_sum(ffb_ctx* ctx, double* data, int n) {
  FFB_TRY(require_group(ctx));
  FFB_CUDA(ctx, cudaSetDevice(ctx->device));
  FFB_REQUIRE(ctx, data && n >= 0, "all-reduce: bad arguments");
  if (n == 0 || ctx->comm.world <= 1) return FFB_OK;
  DevBuf buf;
  FFB_TRY(buf.alloc(ctx, (size_t)n * 8));
  FFB_TRY(ffb_h2d(ctx, buf.p, data, (size_t)n * 8));
  FFB_TRY(ffbi_allreduce_sum(ctx, buf.as<double>(), n));
  FFB_TRY(ffb_d2h(ctx, data, buf.p, (size_t)n * 8));
  FFB_TRY(ffbi_comm_fetch_error(ctx));
  FFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ffbi_comm_check_error(ctx);
}

// ---- symmetric data window ------------------------------------------------------------------------
int ffb_comm_data_window(ffb_ctx* ctx, size_t bytes, void* handle_out) {
  FFB_TRY(require_group(ctx));
  FFB_CUDA(ctx, cudaSetDevice(ctx->device));
  FFB_REQUIRE(ctx, handle_out && bytes > 0, "data window: bad arguments");
  ffb_comm& c = ctx->comm;
  FFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < c.world; ++r) {  // the caller has synchronised the ranks: nobody uses the old one
    if (r != c.rank && c.peer_data[r]) cudaIpcCloseMemHandle(c.peer_data[r]);
    c.peer_data[r] = nullptr;
  }
  if (c.data) FFB_CUDA(ctx, cudaFree(c.data));
  c.data = nullptr;
  c.data_bytes = 0;
  bytes = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  FFB_CUDA(ctx, cudaMalloc(&c.data, bytes));
  c.data_bytes = bytes;
  cudaIpcMemHandle_t h;
  FFB_CUDA(ctx, cudaIpcGetMemHandle(&h, c.data));
  std::memcpy(handle_out, &h, sizeof(h));
  return FFB_OK;
}

int ffb_comm_data_connect(ffb_ctx* ctx, const void* handles) {
  FFB_TRY(require_group(ctx));
  FFB_CUDA(ctx, cudaSetDevice(ctx->device));
  ffb_comm& c = ctx->comm;
  FFB_REQUIRE(ctx, handles && c.data, "data window: not allocated");
  for (int r = 0; r < c.world; ++r) {
    if (r == c.rank) {
      c.peer_data[r] = c.data;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&c.peer_data[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      c.peer_data[r] = nullptr;
      return ffb_fail(ctx, FFB_ECUDA, "data window: cannot map rank %d into rank %d (%s)", r, c.rank,
                      cudaGetErrorString(e));
    }
  }
  return FFB_OK;
}

// All-gather of column blocks over NVLink.  `local` (host, or a result array with a device shadow) is
// this rank's (rows, counts[rank]) block of complex128; `out` (host) receives (rows, sum(counts)).
int ffb_allgather_columns(ffb_ctx* ctx, int rows, const int* counts, const double* local,
                          double* out) {
  FFB_TRY(require_group(ctx));
  FFB_CUDA(ctx, cudaSetDevice(ctx->device));
  ffb_comm& c = ctx->comm;
  FFB_REQUIRE(ctx, rows >= 1 && counts && out, "all-gather: bad arguments");
  size_t ld = 0, col0 = 0;
  for (int r = 0; r < c.world; ++r) {
    FFB_REQUIRE(ctx, counts[r] >= 0, "all-gather: negative block size");
    if (r < c.rank) col0 += (size_t)counts[r];
    ld += (size_t)counts[r];
  }
  const int mine = counts[c.rank];
  FFB_REQUIRE(ctx, mine == 0 || local, "all-gather: null block");
  const size_t total_bytes = (size_t)rows * ld * 16;
  if (total_bytes == 0) return FFB_OK;
  FFB_REQUIRE(ctx, c.data && c.data_bytes >= total_bytes && c.peer_data[c.rank],
              "all-gather: the data window (%zu bytes) is smaller than the result (%zu bytes)",
              c.data_bytes, total_bytes);
  DevBuf token, block;
  FFB_TRY(token.alloc(ctx, 8));
  FFB_CUDA(ctx, cudaMemsetAsync(token.p, 0, 8, ctx->stream));
  // (1) everybody is done with the previous contents of the windows (their downloads are stream-ordered
  //     before this barrier)
  FFB_TRY(ffbi_allreduce_sum(ctx, token.as<double>(), 1));
  if (mine > 0) {
    const size_t bytes = (size_t)rows * mine * 16;
    // a block that is still mirrored on the device (the filter function the pipeline just computed) is
    // sent from there; otherwise it is uploaded first
    const double2* src = static_cast<const double2*>(ffb_shadow_lookup(ctx, local, bytes));
    if (!src) {
      FFB_TRY(block.alloc(ctx, bytes));
      FFB_TRY(ffb_h2d(ctx, block.p, local, bytes));
      src = block.as<const double2>();
    }
    Windows w;
    for (int r = 0; r < c.world; ++r) w.p[r] = reinterpret_cast<unsigned long long>(c.peer_data[r]);
    const int blocks = (int)std::min<size_t>((size_t)ctx->sm_count * 8, ceil_div_sz((size_t)rows * mine, 256));
    peer_put_kernel<<<blocks, 256, 0, ctx->stream>>>(w, c.world, rows, mine, ld, col0, src);
    FFB_LAUNCHED(ctx);
  }
  // (2) all blocks have landed everywhere
  FFB_TRY(ffbi_allreduce_sum(ctx, token.as<double>(), 1));
  FFB_TRY(ffb_d2h(ctx, out, c.data, total_bytes));
  FFB_TRY(ffbi_comm_fetch_error(ctx));
  FFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ffbi_comm_check_error(ctx);
}

}  // extern "C"
