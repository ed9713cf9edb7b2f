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
// TMA was measured for this access pattern twice and is not used.  Round 1: a warp-specialised variant
// (one producer warp per CTA issuing 1-D cp.async.bulk copies into mbarrier-guarded stages) -- a bulk
// copy takes uniform-register operands, so per-row copies are serialised (ncu: 8 UBLKCP per 8-entry step),
// one producer warp tops out at ~17 rows/us and costs more issue slots per entry than the vector cp.async
// of the consuming warp.  Round 2: tools/gather_probe.cu times cp.async.bulk.tensor ... tile::gather4
// (UTMALDG.2D.GATHER4, 4 rows per instruction) against ld.global / cp.async on random objects out of a
// 25.6 GB array (profiles/r02_gather_probe.jsonl): 3.0e10 objects/s against 3.65e10 for objects <= 128 B,
// the same 6.7-6.9 TB/s for 512-byte rows.  TMA pays for large contiguous tiles, not for random rows.
//
// Work decomposition (merge-path style): the CSR entries are cut into chunks of DL_CH consecutive
// entries regardless of row boundaries; DL_RANGE consecutive chunks form a range owned by one
// warp.  Every warp therefore does the same amount of work whatever the degree distribution -- hub
// rows need no special path -- and rows that cross a range boundary are stitched together by a
// small fix-up kernel in a fixed order (deterministic, no atomics).
#pragma once
#include "dl_common.cuh"

#define DL_CH 32          // entries per chunk
#define DL_RANGE (1LL << g.range_shift)   // chunks per range of the graph `g` in scope: 64 (2048 entries) on large
                                          // graphs, fewer on small ones (dl_range_shift, dl_common.cuh)
#define DL_HS 4           // entries per ring stage (one sub-block of 4 edges)
#define DL_QPC (DL_CH / DL_HS)
#define DL_OWNQ 2         // staged own-row slots per stage (more distinct rows -> plain loads)
#define DL_RING 3         // stages per warp ring: DL_RING-1 in flight while one is consumed

__device__ __forceinline__ unsigned dl_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
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
  int sh;                                        // log2(chunks per range)
  __device__ __forceinline__ void init(long long nnz, long long total_warps, int range_shift) {
    sh = range_shift;
    n_chunks = (nnz + DL_CH - 1) / DL_CH;
    n_ranges = (n_chunks + (1LL << sh) - 1) >> sh;
    GW = total_warps;
  }
  __device__ __forceinline__ long long first(long long gw) const { return gw < n_ranges ? (gw << sh) : -1; }
  __device__ __forceinline__ long long next(long long c) const {
    if (c < 0) return -1;
    const long long c1 = c + 1;
    if ((c1 & ((1LL << sh) - 1)) != 0) return c1 < n_chunks ? c1 : -1;
    const long long rg = (c >> sh) + GW;         // first chunk of this warp's next range
    return rg < n_ranges ? (rg << sh) : -1;
  }
};
