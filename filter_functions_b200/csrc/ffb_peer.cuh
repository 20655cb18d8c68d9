// Device side of the peer group: an all-reduce of single doubles over NVLink peer memory, written so
// that it can be the EPILOGUE of a compute kernel (the thread that holds a finished partial result
// exchanges it with the other GPUs before it stores the sum) -- SURVEY.md 8e: the only exchange on the
// path is the sum of <= n_nops^2 partial integrals.
//
// Protocol (low-latency, flag-in-data, as in NCCL's LL protocol): every rank owns an exchange window
//   slot[parity][source rank][FFB_PEER_SLOTS] of 16 bytes = {value.lo, seq, value.hi, seq}
// mapped into all peers with CUDA IPC.  A rank STORES its value into slot[seq & 1][rank][i] of every
// window (its own included) and then POLLS its own window -- local memory, the stores travel over
// NVLink -- until the flags of all `world` sources equal `seq`.  Each 8-byte half carries its own flag,
// so a torn 16-byte store is harmless.  The values are added in rank order: every rank forms the same
// floating-point sum.  Two parities suffice: a rank can only be one collective ahead of a peer,
// because finishing collective k needs the peer's contribution to k, which the peer sends after it
// has finished reading k - 1.  No atomics, no barriers, no host round trip.
#pragma once

#include "ffb_common.cuh"

__device__ __forceinline__ unsigned long long ffb_globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// All-reduce (sum) of `v` with the values the other ranks pass for the same `slot` of the same
// collective.  Called by ONE thread per slot; slot < FFB_PEER_SLOTS.
__device__ __forceinline__ double peer_allreduce_slot(const PeerReduce& pr, int slot, double v) {
  if (pr.world <= 1) return v;
  const unsigned seq = pr.seq;
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  const size_t mine = ((size_t)((seq & 1u) * pr.world + pr.rank) * FFB_PEER_SLOTS + slot) * 16;
  for (int r = 0; r < pr.world; ++r) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(pr.window[r] + mine),
                 "r"(lo), "r"(seq), "r"(hi), "r"(seq)
                 : "memory");
  }
  double sum = 0.0;
  const unsigned long long t0 = ffb_globaltimer_ns();
  const unsigned long long limit = (unsigned long long)pr.timeout_ms * 1000000ull;
  for (int r = 0; r < pr.world; ++r) {
    const unsigned long long src =
        pr.window[pr.rank] + ((size_t)((seq & 1u) * pr.world + r) * FFB_PEER_SLOTS + slot) * 16;
    unsigned a, fa, b, fb;
    unsigned spins = 0;
    while (true) {
      asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb)
                   : "l"(src)
                   : "memory");
      if (fa == seq && fb == seq) break;
      if ((++spins & 1023u) == 0 && ffb_globaltimer_ns() - t0 > limit) {
        if (pr.error) atomicExch(pr.error, 1 + r);
        return __longlong_as_double(0x7ff8000000000000ll);  // NaN: the caller sees the failure
      }
    }
    sum += __hiloint2double((int)b, (int)a);
  }
  return sum;
}
