// dl_common.cuh -- shared device code of libdisenlink_b200.so (sm_100a).
//
// Canonical arithmetic.  Hard routing (argmax over factors, reference model.py:61) is
// discontinuous, so everything that feeds it is evaluated in ONE fixed fp32 order, the same in
// every kernel, independent of which endpoint is "the row" and of how rows are split over warps
// or GPUs.  That makes kstar / w bitwise symmetric ((i,j) vs (j,i)), which the atomic-free
// backward relies on, and run-to-run / partition-to-partition deterministic.
//   dot   : elements in float4 chunks (scalars when d % 4 != 0); a chunk is an FMA chain seeded
//           with +0; chunk partials are combined by a balanced binary tree, adjacent pairs first,
//           zero padded to a power of two (a lane butterfly with xor offsets 1,2,4,... is exactly
//           that tree because fp addition is commutative).
//   q     : dot / T  (IEEE division, model.py:56 divides by the temperature)
//   exp   : dl_expf  (Cody-Waite + degree-5 polynomial, explicit __fmaf_rn -- no MUFU, so the
//           value does not depend on the approximate-unit implementation)
//   sum_k : sequential k = 0..K-1;  a_k = e_k / sum  (IEEE division)
//   argmax: first strict maximum; NaN counts as maximum (torch.argmax semantics)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/disenlink_b200.h"

#define DL_FULL 0xffffffffu
#define DL_MAX_PEER_OUT 15
#define DL_WARPS_PER_CTA 8
#define DL_CTA (DL_WARPS_PER_CTA * 32)
// the register-heavy, double-buffered row-gather kernels use small CTAs so the register file is
// filled at a finer granularity
#define DL_WARPS_PER_CTA_S 4
#define DL_CTA_S (DL_WARPS_PER_CTA_S * 32)

#define DL_CUDA_TRY(expr)                    \
  do {                                       \
    cudaError_t _e = (expr);                 \
    if (_e != cudaSuccess) return (int)_e;   \
  } while (0)

#define DL_LAUNCH_CHECK()                    \
  do {                                       \
    cudaError_t _e = cudaPeekAtLastError();  \
    if (_e != cudaSuccess) return (int)_e;   \
  } while (0)

// ---- device copy of dl_graph --------------------------------------------------------------
struct DlGraphDev {
  long long N, nnz;
  const long long* __restrict__ rowptr;
  const int* __restrict__ col;
  const int* __restrict__ perm;
  long long n_hub, n_hub_items;
  const long long* __restrict__ hub_seg_ptr;
  const int* __restrict__ item_hub;
  const int* __restrict__ erow;
  long long row_base;
  // Node-partitioned runs: the peers' copies of the row array the kernel produces (H for the
  // aggregation, dH for the decoder backward), full-size and indexed by global node id like the
  // local one.  A kernel that finishes an owned row stores it into every peer as well, so the
  // all-gather rides on the kernel (NVLink stores) instead of following it.  Set by the entry points
  // from their peer arguments; 0 everywhere else.
  float* peer_out[DL_MAX_PEER_OUT];
  int n_peer_out;
  // Streaming kernels: the entries are cut into 32-entry chunks and (1 << range_shift) consecutive chunks
  // form a range owned by one warp.  64 chunks per range on large graphs; fewer on small ones, so that
  // there are always enough ranges for every warp of the device (a 10^4-entry graph cut into 2048-entry
  // ranges would leave 5 warps walking them serially at one DRAM / L2 round trip per 4 entries).
  int range_shift;
};

// a pure function of nnz: every kernel and every scratch-size query of one graph must agree on it
static inline __host__ __device__ int dl_range_shift(long long nnz) {
  const long long n_chunks = (nnz + 31) / 32;
  int sh = 6;
  while (sh > 0 && (n_chunks >> sh) < 8192) --sh;
  return sh;
}

static inline DlGraphDev dl_graph_dev(const dl_graph* g) {
  DlGraphDev o;
  o.N = g->N; o.nnz = g->nnz;
  o.rowptr = (const long long*)g->rowptr;
  o.col = g->col; o.perm = g->perm;
  o.n_hub = g->n_hub; o.n_hub_items = g->n_hub_items;
  o.hub_seg_ptr = (const long long*)g->hub_seg_ptr;
  o.item_hub = g->item_hub;
  o.erow = g->erow;
  o.row_base = g->row_base;
  o.n_peer_out = 0;
  for (int q = 0; q < DL_MAX_PEER_OUT; ++q) o.peer_out[q] = nullptr;
  o.range_shift = dl_range_shift(g->nnz);
  return o;
}

static inline int dl_set_peer_out(DlGraphDev& g, float* const* peers, int n_peers) {
  if (n_peers < 0 || n_peers > DL_MAX_PEER_OUT || (n_peers > 0 && !peers)) return 0;
  for (int q = 0; q < n_peers; ++q) {
    if (!peers[q]) return 0;
    g.peer_out[q] = peers[q];
  }
  g.n_peer_out = n_peers;
  return 1;
}

static inline int dl_graph_ok(const dl_graph* g) {
  if (!g || g->N < 0 || g->nnz < 0) return 0;
  if (g->N > 0 && (!g->rowptr || !g->perm)) return 0;
  if (g->nnz > 0 && !g->col) return 0;
  if (g->n_hub < 0 || g->n_hub > g->N || g->n_hub_items < 0) return 0;
  if (g->n_hub > 0 && (!g->hub_seg_ptr || !g->item_hub)) return 0;
  if (g->row_base < 0) return 0;
  return 1;
}

// One work item = one row, or one DL_SEG-edge segment of a hub row.
struct DlItem {
  int row;             // local row (indexes rowptr)
  long long node;      // global node id = row_base + row (indexes Z, H, s, r, G, dZ)
  long long e0, e1;
  long long hub_slot;  // >= 0: partial result goes to scratch slot hub_slot; -1: direct
};

__device__ __forceinline__ long long dl_num_items(const DlGraphDev& g) {
  return g.n_hub_items + (g.N - g.n_hub);
}

__device__ __forceinline__ DlItem dl_decode_item(const DlGraphDev& g, long long t) {
  DlItem it;
  if (t < g.n_hub_items) {
    int h = __ldg(g.item_hub + t);
    it.row = __ldg(g.perm + h);
    long long seg = t - __ldg(g.hub_seg_ptr + h);
    long long r0 = __ldg(g.rowptr + it.row), r1 = __ldg(g.rowptr + it.row + 1);
    it.e0 = r0 + seg * DL_SEG;
    it.e1 = min(r1, it.e0 + (long long)DL_SEG);
    it.hub_slot = t;
  } else {
    long long idx = t - g.n_hub_items + g.n_hub;
    it.row = __ldg(g.perm + idx);
    it.e0 = __ldg(g.rowptr + it.row);
    it.e1 = __ldg(g.rowptr + it.row + 1);
    it.hub_slot = -1;
  }
  it.node = g.row_base + it.row;
  return it;
}

// Software-pipelined work-item iterator.  Hub segments come first (they are long, the decode cost is
// amortised), then the regular rows in NATURAL order with a warp stride, skipping hub rows.  The row
// bounds of the next two rows of this warp are always in flight, so the rowptr -> col -> gather
// dependency chain of one row overlaps the gathers of the previous one.
struct DlRowIter {
  long long W, t, r, a0, a1, b0, b1;
  __device__ __forceinline__ void init(const DlGraphDev& g, long long warp0, long long nwarps) {
    W = nwarps; t = warp0; r = warp0;
    a0 = a1 = b0 = b1 = 0;
    if (r < g.N) { a0 = __ldg(g.rowptr + r); a1 = __ldg(g.rowptr + r + 1); }
    if (r + W < g.N) { b0 = __ldg(g.rowptr + r + W); b1 = __ldg(g.rowptr + r + W + 1); }
  }
  __device__ __forceinline__ bool next(const DlGraphDev& g, DlItem& it) {
    if (t < g.n_hub_items) {
      it = dl_decode_item(g, t);
      t += W;
      return true;
    }
    while (r < g.N) {
      const long long cr = r, c0 = a0, c1 = a1;
      r += W; a0 = b0; a1 = b1;
      if (r + W < g.N) { b0 = __ldg(g.rowptr + r + W); b1 = __ldg(g.rowptr + r + W + 1); }
      if (c1 - c0 >= (long long)DL_SEG) continue;   // hub row: handled as segments above
      it.row = (int)cr; it.node = g.row_base + cr; it.e0 = c0; it.e1 = c1; it.hub_slot = -1;
      return true;
    }
    return false;
  }
};

// ---- canonical exp ------------------------------------------------------------------------
__device__ __forceinline__ float dl_expf(float x) {
  // clamps written as selects so that a NaN argument flows through to a NaN result without a branch
  x = (x > 89.0f) ? 89.0f : x;
  x = (x < -104.0f) ? -104.0f : x;
  float t = __fmul_rn(x, 1.44269504088896341f);
  float n = rintf(t);
  float r = __fmaf_rn(n, -0.693359375f, x);
  r = __fmaf_rn(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = __fmaf_rn(p, r, 1.3981999507e-3f);
  p = __fmaf_rn(p, r, 8.3334519073e-3f);
  p = __fmaf_rn(p, r, 4.1665795894e-2f);
  p = __fmaf_rn(p, r, 1.6666665459e-1f);
  p = __fmaf_rn(p, r, 5.0000001201e-1f);
  float r2 = __fmul_rn(r, r);
  p = __fmaf_rn(p, r2, r);
  p = __fadd_rn(p, 1.0f);
  int ni = (int)n;
  int n1 = ni >> 1;
  int n2 = ni - n1;
  float s1 = __int_as_float((n1 + 127) << 23);
  float s2 = __int_as_float((n2 + 127) << 23);
  return __fmul_rn(__fmul_rn(p, s1), s2);
}

__device__ __forceinline__ float dl_sigmoid(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, dl_expf(-x)));
}

__device__ __forceinline__ float dl_chunk_dot(const float4& a, const float4& b) {
  float p = __fmaf_rn(a.x, b.x, 0.0f);
  p = __fmaf_rn(a.y, b.y, p);
  p = __fmaf_rn(a.z, b.z, p);
  p = __fmaf_rn(a.w, b.w, p);
  return p;
}

// first strict maximum over a[0..K), NaN counts as the maximum (torch.argmax)
template <int K>
__device__ __forceinline__ int dl_first_argmax(const float (&a)[K]) {
  int best = 0;
  float bv = a[0];
#pragma unroll
  for (int k = 1; k < K; ++k) {
    float v = a[k];
    if (v > bv || (v != v && bv == bv)) { bv = v; best = k; }
  }
  return best;
}

__device__ __forceinline__ float4 dl_ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ float4 dl_zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ void dl_fma4(float4& acc, float c, const float4& v) {
  acc.x = __fmaf_rn(c, v.x, acc.x);
  acc.y = __fmaf_rn(c, v.y, acc.y);
  acc.z = __fmaf_rn(c, v.z, acc.z);
  acc.w = __fmaf_rn(c, v.w, acc.w);
}

// ---- lane mapping for a compile-time (K, d) with d % 4 == 0, d <= 64 -----------------------
// A row of D = K*d floats is K*L float4 chunks, L = d/4.  LP = L padded to a power of two lanes.
// FPP = 32/LP factors fit in one pass over the warp; NP = ceil(K/FPP) passes.  In pass t lane
// (slot = lane/LP, g = lane%LP) owns chunk g of factor k = t*FPP + slot (idle if k>=K or g>=L).
// EB = min(4, LP) edges are processed together ("sub-block"): after a reduce-scatter over the
// LP-lane group, lane g holds the finished dot of edge dl_edge_of_lane(g).
template <int K_, int d_>
struct DlMap {
  static constexpr int K = K_;
  static constexpr int d = d_;
  static constexpr int D = K_ * d_;
  static constexpr int L = d_ / 4;
  static constexpr int LP = (L <= 1) ? 1 : (L <= 2) ? 2 : (L <= 4) ? 4 : (L <= 8) ? 8 : (L <= 16) ? 16 : 32;
  static constexpr int FPP = 32 / LP;
  static constexpr int NP = (K_ + FPP - 1) / FPP;
  static constexpr int EB = (LP < 4) ? LP : 4;
  static_assert(d_ % 4 == 0 && d_ >= 4 && d_ <= 64, "fast path needs d % 4 == 0, 4 <= d <= 64");
  static_assert(K_ >= 1 && K_ <= DL_MAX_K, "K out of range");

  __device__ static __forceinline__ int slot(int lane) { return lane / LP; }
  __device__ static __forceinline__ int g(int lane) { return lane % LP; }
  // factor owned in pass t (may be >= K: idle)
  __device__ static __forceinline__ int factor(int lane, int t) { return t * FPP + slot(lane); }
  __device__ static __forceinline__ bool active(int lane, int t) {
    return factor(lane, t) < K && g(lane) < L;
  }
  // float offset of the lane's chunk inside a row, pass t
  __device__ static __forceinline__ int offset(int lane, int t) {
    return factor(lane, t) * d + 4 * g(lane);
  }
  // which of the EB edges of a sub-block this lane holds after the reduce-scatter
  __device__ static __forceinline__ int edge_of_lane(int lane) {
    if (EB == 4) return 2 * (lane & 1) + ((lane >> 1) & 1);
    if (EB == 2) return lane & 1;
    return 0;
  }
  // lane (inside a group) that holds edge e after the reduce-scatter (inverse of edge_of_lane)
  __device__ static __forceinline__ int lane_of_edge(int e) {
    if (EB == 4) return ((e >> 1) & 1) | ((e & 1) << 1);
    return e;
  }
};

// Reduce-scatter EB per-edge chunk partials over the LP-lane group in canonical tree order.
// On return every lane holds the complete dot of edge M::edge_of_lane(lane) (replicated over
// the LP/EB lanes that share the low bits).
template <class M>
__device__ __forceinline__ float dl_reduce_scatter(const float (&p)[M::EB], int lane) {
  float v;
  if (M::EB == 4) {
    const bool b0 = lane & 1, b1 = lane & 2;
    float s0 = b0 ? p[0] : p[2];
    float s1 = b0 ? p[1] : p[3];
    float r0 = __shfl_xor_sync(DL_FULL, s0, 1);
    float r1 = __shfl_xor_sync(DL_FULL, s1, 1);
    float a0 = __fadd_rn(b0 ? p[2] : p[0], r0);
    float a1 = __fadd_rn(b0 ? p[3] : p[1], r1);
    float s2 = b1 ? a0 : a1;
    float r2 = __shfl_xor_sync(DL_FULL, s2, 2);
    v = __fadd_rn(b1 ? a1 : a0, r2);
#pragma unroll
    for (int off = 4; off < M::LP; off <<= 1) v = __fadd_rn(v, __shfl_xor_sync(DL_FULL, v, off));
  } else if (M::EB == 2) {
    const bool b0 = lane & 1;
    float r0 = __shfl_xor_sync(DL_FULL, b0 ? p[0] : p[1], 1);
    v = __fadd_rn(b0 ? p[1] : p[0], r0);
  } else {
    v = p[0];
  }
  return v;
}

// Plain butterfly over the LP-lane group (all lanes get the sum), canonical tree order.
template <class M>
__device__ __forceinline__ float dl_group_sum(float v) {
#pragma unroll
  for (int off = 1; off < M::LP; off <<= 1) v = __fadd_rn(v, __shfl_xor_sync(DL_FULL, v, off));
  return v;
}

// ---- launch geometry ----------------------------------------------------------------------
// Persistent-style grid: (#SMs) x (resident CTAs per SM) CTAs of DL_CTA threads, warps stride
// over the work items.
template <class Kernel>
static inline int dl_grid_for(Kernel kern, long long n_items, int* grid_out, size_t smem_bytes = 0,
                              int cta_threads = DL_CTA) {
  int dev = 0, sms = 0, per_sm = 0;
  DL_CUDA_TRY(cudaGetDevice(&dev));
  DL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, cta_threads, smem_bytes));
  if (per_sm < 1) per_sm = 1;
  const int wpc = cta_threads / 32;
  long long want = (n_items + wpc - 1) / wpc;
  long long cap = (long long)sms * per_sm;
  long long grid = want < cap ? want : cap;
  if (grid < 1) grid = 1;
  *grid_out = (int)grid;
  return 0;
}
