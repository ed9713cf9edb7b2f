// dl_stream.cuh -- shared-memory staged gather pipeline for the row-gather kernels (sm_100a).
//
// Why: these kernels are HBM-bound only if ~45 KB per SM are in flight at all times (6.4 TB/s x
// ~1 us / 148 SMs).  With register staging a warp alternates "issue loads / wait / compute",
// registers cap the bytes in flight at 2-4 KB per warp, and loop-carried register rotation stalls
// on the in-flight loads (ncu: the top stall sites were MOVs of just-loaded registers).  Here the
// bytes in flight live in shared memory: every warp owns a ring of DL_RING stages of DL_HS entries
// and keeps DL_RING-1 stages of neighbour rows in flight with cp.async (LDGSTS: one warp
// instruction moves one 512-byte row, completion tracked with commit / wait groups) while it
// computes on the oldest stage.
//
// A warp-specialised variant (one producer warp per CTA issuing cp.async.bulk / TMA 1-D copies
// into mbarrier-guarded stages, helpers below) was built and measured first: it is correct, but a
// bulk copy takes uniform-register operands, so per-row copies are serialised (ncu: 8 UBLKCP per
// 8-entry step in a loop), one producer warp tops out at ~17 rows/us, and the producers cost more
// issue slots per entry than the vector cp.async issued by the consuming warp itself.  TMA pays for
// large contiguous tiles, not for 512-byte random rows.
//
// Work decomposition (merge-path style): the CSR entries are cut into chunks of DL_CH consecutive
// entries regardless of row boundaries; DL_RANGE consecutive chunks form a range owned by one
// warp.  Every warp therefore does the same amount of work whatever the degree distribution -- hub
// rows need no special path -- and rows that cross a range boundary are stitched together by a
// small fix-up kernel in a fixed order (deterministic, no atomics).
#pragma once
#include "dl_common.cuh"

#define DL_CH 32          // entries per chunk
#define DL_RANGE 64       // chunks per range (2048 entries)
#define DL_HS 4           // entries per ring stage (one sub-block of 4 edges)
#define DL_QPC (DL_CH / DL_HS)
#define DL_OWNQ 2         // staged own-row slots per stage (more distinct rows -> plain loads)
#define DL_RING 3         // stages per warp ring: DL_RING-1 in flight while one is consumed

__device__ __forceinline__ unsigned dl_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void dl_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dl_smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void dl_mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void dl_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dl_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void dl_mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dl_smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void dl_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DL_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DL_DONE_%=;\n"
      "bra DL_WAIT_%=;\n"
      "DL_DONE_%=:\n"
      "}\n" ::"r"(dl_smem_u32(bar)), "r"(parity)
      : "memory");
}

// TMA 1-D bulk copy global -> shared, completion counted on `bar` (bytes must be a multiple of 16,
// both addresses 16-byte aligned)
__device__ __forceinline__ void dl_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes,
                                            unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"(dl_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(dl_smem_u32(bar))
               : "memory");
}

// 16-byte cp.async global -> shared (LDGSTS), L2-only (.cg): gathered rows are not reused via L1
#ifndef DL_CPA_L2
#define DL_CPA_L2 ""          // optional L2 prefetch-size qualifier for experiments, e.g. ".L2::64B"
#endif
__device__ __forceinline__ void dl_cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global" DL_CPA_L2 " [%0], [%1], 16;" ::"r"(dl_smem_u32(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void dl_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void dl_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ float4 dl_lds4(const void* p) { return *reinterpret_cast<const float4*>(p); }

// per-warp stream of chunks: ranges gw, gw + GW, ... of DL_RANGE consecutive chunks each
struct DlChunkStream {
  long long n_chunks, n_ranges, GW;
  __device__ __forceinline__ void init(long long nnz, long long total_warps) {
    n_chunks = (nnz + DL_CH - 1) / DL_CH;
    n_ranges = (n_chunks + DL_RANGE - 1) / DL_RANGE;
    GW = total_warps;
  }
  __device__ __forceinline__ long long first(long long gw) const { return gw < n_ranges ? gw * DL_RANGE : -1; }
  __device__ __forceinline__ long long next(long long c) const {
    if (c < 0) return -1;
    const long long c1 = c + 1;
    if (c1 % DL_RANGE != 0) return c1 < n_chunks ? c1 : -1;
    const long long rg = c / DL_RANGE + GW;      // first chunk of this warp's next range
    return rg < n_ranges ? rg * DL_RANGE : -1;
  }
};

// geometry of the chunk / range / span decomposition of the warp-specialised variant
struct DlSpanGeom {
  long long n_chunks, n_ranges, n_spans;
  __host__ __device__ static DlSpanGeom make(long long nnz, int nc) {
    DlSpanGeom s;
    s.n_chunks = (nnz + DL_CH - 1) / DL_CH;
    s.n_ranges = (s.n_chunks + DL_RANGE - 1) / DL_RANGE;
    s.n_spans = (s.n_ranges + nc - 1) / nc;
    return s;
  }
  // chunk id handled by consumer w of span sp at step j, or -1
  __device__ __forceinline__ long long chunk(long long sp, int nc, int w, int j) const {
    const long long rg = sp * nc + w;
    if (rg >= n_ranges) return -1;
    const long long c = rg * DL_RANGE + j;
    return c < n_chunks ? c : -1;
  }
};
