// factor_fwd.cu -- forward of DisenLink's factor-aware message passing on a CSR.
//
//   k_edge_attn_fwd   [ref: model.py:56-73]  per entry (i,j): q_k = z_i^k.z_j^k / T, softmax over
//                     the K factors, hard routing kstar = first argmax, w = a[kstar]; per row the
//                     routed sums s[i,k] (zeros -> 1).  Everything stays in registers; only
//                     (kstar: u8, w: f32) per entry and s [N,K] are written.
//   k_factor_gather   [ref: model.py:75]     H[i,k] = beta Z[i,k] + (1-beta) sum_{j: kstar=k}
//                     (w_ij / s[j,k]) Z[j,k]  -- gather / segment-sum, no atomics.  The same kernel
//                     (MODE 1) is pass 1 of the backward: T_[i,k] = (1-beta)/s[i,k] sum w G[j,k].
//
// Work distribution: one warp per work item (a row, or a DL_SEG-edge segment of a hub row), rows in
// natural order with a warp stride, software-pipelined (DlRowIter + per-kernel prefetch of the next
// item's first column block / row vectors) so the rowptr -> col -> gather dependency chain of one
// row overlaps the gathers of the previous row.
//
// Attention mapping (DlMap<K,d>): a row of D = K*d floats is spread over the lanes as float4
// chunks, so a neighbour row is one coalesced 128-bit-per-lane gather.  Four edges are in flight
// per warp step; their chunk partials are reduce-scattered over the d/4-lane factor group
// (3 shuffles per 4 edges instead of 8), giving lane (k, g) the finished dot of edge g, factor k.
//
// Slice-gather mapping: only the kstar slice (d floats) of a neighbour is needed, so the warp is
// split into NG = 32/LP lane groups that fetch NG different edges per load instruction (8 x 64 B
// at d = 16); each lane keeps K float4 accumulators (predicated on the edge's factor), and the NG
// group partials are combined in a fixed order through shared memory at the end of the row.
//
// HBM bytes per entry (D=128, K=8, d=16): attention 4 (col) + 512 (z_j) + 5 (kstar, w);
// aggregation 4 + 5 + 4 (s[j,k]) + 64 (z_j^k slice).  Per row: z_i, s / H.  DESIGN.md section 4.
#include <math_constants.h>
#include <stdlib.h>

#include "dl_dispatch.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// attention, fast path
// ---------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(DL_CTA_S)
k_edge_attn_fwd(DlGraphDev g, const float* __restrict__ Z, float T, unsigned char* __restrict__ kstar,
                float* __restrict__ w, float* __restrict__ s, float* __restrict__ hub_ws) {
  constexpr int K = M::K, D = M::D, NP = M::NP, EB = M::EB, LP = M::LP, FPP = M::FPP;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA_S + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA_S;

  int off[NP];
  bool act[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { off[p] = M::offset(lane, p); act[p] = M::active(lane, p); }
  const int my_e = M::edge_of_lane(lane);
  const int gsrc = lane & (EB - 1);
  const bool primary = (lane % LP) < EB;  // one replica per (factor, edge) accumulates s
  const bool unit_T = (T == 1.0f);

  // gather the EB neighbour rows of one sub-block: columns idx0.. of the 32-column block held in
  // `colreg` (one column per lane), valid while idx < lim
  auto fetch = [&](float4 (&dst)[EB][NP], int colreg, int idx0, int lim) {
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      const int idx = idx0 + e;
      const int c = __shfl_sync(DL_FULL, colreg, idx & 31);
      const bool valid = idx < lim;
#pragma unroll
      for (int p = 0; p < NP; ++p)
        dst[e][p] = (valid && act[p]) ? dl_ldg4(Z + (long long)c * D + off[p]) : dl_zero4();
    }
  };

  DlRowIter itr;
  itr.init(g, warp0, nwarps);
  DlItem it, nit;
  it = nit = DlItem{0, 0, 0, 0, -1};
  bool have = itr.next(g, it);
  int colA = 0;
  if (have && lane < it.e1 - it.e0) colA = __ldg(g.col + it.e0 + lane);
  // zjA always holds the rows of the sub-block about to be computed; while it is being computed
  // the rows of the NEXT sub-block in stream order (same block / next block of the row / first
  // block of the next item) are in flight into zjB.
  float4 zjA[EB][NP], zjB[EB][NP];
  bool pre = false;   // zjA already holds the first sub-block of `it`

  while (have) {
    const bool have_n = itr.next(g, nit);
    const int ncntB = have_n ? (int)min(32LL, nit.e1 - nit.e0) : 0;
    int colB = 0;
    if (lane < ncntB) colB = __ldg(g.col + nit.e0 + lane);
    float4 zi[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) zi[p] = act[p] ? dl_ldg4(Z + it.node * D + off[p]) : dl_zero4();
    if (!pre && it.e1 > it.e0) fetch(zjA, colA, 0, (int)min(32LL, it.e1 - it.e0));
    bool pre_n = false;

    float sacc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) sacc[p] = 0.0f;
    int mycol = colA, colN = 0;
    for (long long base = it.e0; base < it.e1; base += 32) {
      const int cnt = (int)min(32LL, it.e1 - base);
      const bool more = base + 32 < it.e1;
      const int ncntN = more ? (int)min(32LL, it.e1 - base - 32) : 0;
      if (lane < ncntN) colN = __ldg(g.col + base + 32 + lane);
      int out_ks = 0;
      float out_w = 0.0f;
      const int nsub = (cnt + EB - 1) / EB;
      for (int sb = 0; sb < nsub; ++sb) {
        {   // prefetch the next sub-block in stream order
          int pc, pi0, plim;
          if (sb + 1 < nsub) { pc = mycol; pi0 = (sb + 1) * EB; plim = cnt; }
          else if (more) { pc = colN; pi0 = 0; plim = ncntN; }
          else { pc = colB; pi0 = 0; plim = ncntB; pre_n = ncntB > 0; }
          fetch(zjB, pc, pi0, plim);
        }
        float ev[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          float part[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) part[e] = dl_chunk_dot(zi[p], zjA[e][p]);
          float q = dl_reduce_scatter<M>(part, lane);
          if (!unit_T) q = __fdiv_rn(q, T);
          ev[p] = dl_expf(q);
        }
        // all-gather the K exponentials of my edge
        float a[K];
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          a[k] = __shfl_sync(DL_FULL, ev[k / FPP], (k % FPP) * LP + gsrc);
          sum = (k == 0) ? a[0] : __fadd_rn(sum, a[k]);
        }
        // routing = first argmax of a_k = e_k / sum.  Division is monotone, so when the largest
        // exponential is separated from every other one by more than 2^-22 relative (>= 2 float
        // spacings) its quotient is strictly the largest and one division suffices; exact ties in e
        // resolve to the first index either way.  Anything closer (or non-finite) takes the
        // literal path.  Both paths return the same bits.
        int ks = 0;
        float emax = a[0];
#pragma unroll
        for (int k = 1; k < K; ++k)
          if (a[k] > emax) { emax = a[k]; ks = k; }
        const float thr = __fmul_rn(emax, 0.99999976158142089844f);  // 1 - 2^-22
        bool slow = !(sum < CUDART_INF_F);
#pragma unroll
        for (int k = 0; k < K; ++k) slow = slow || (a[k] > thr && a[k] != emax);
        float wv;
        if (!slow) {
          wv = __fdiv_rn(emax, sum);
        } else {
          ks = 0;
          wv = 0.0f;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            float v = __fdiv_rn(a[k], sum);
            if (k == 0) { wv = v; }
            else if (v > wv || (v != v && wv == wv)) { wv = v; ks = k; }
          }
        }
        const bool valid = (sb * EB + my_e) < cnt;
#pragma unroll
        for (int p = 0; p < NP; ++p)
          if (valid && primary && ks == M::factor(lane, p)) sacc[p] = __fadd_rn(sacc[p], wv);
        if (sb == lane / EB) { out_ks = ks; out_w = wv; }
#pragma unroll
        for (int e = 0; e < EB; ++e)
#pragma unroll
          for (int p = 0; p < NP; ++p) zjA[e][p] = zjB[e][p];
      }
      const int oi = (lane & ~(EB - 1)) + my_e;
      if (oi < cnt) {
        kstar[base + oi] = (unsigned char)out_ks;
        w[base + oi] = out_w;
      }
      mycol = colN;
    }
    // per-row routed sums: combine the EB interleaved chains, fixed tree
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float v = sacc[p];
#pragma unroll
      for (int o = 1; o < EB; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(DL_FULL, v, o));
      const int k = M::factor(lane, p);
      if (k < K && (lane % LP) == 0) {
        if (it.hub_slot >= 0) hub_ws[it.hub_slot * K + k] = v;
        else s[it.node * K + k] = (v == 0.0f) ? 1.0f : v;
      }
    }
    it = nit;
    have = have_n;
    colA = colB;
    pre = pre_n;
  }
}

// ---------------------------------------------------------------------------------------------
// attention, runtime-generic path: one warp per item, one entry at a time
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DL_CTA)
k_edge_attn_fwd_generic(DlGraphDev g, const float* __restrict__ Z, int K, int d, float T,
                        unsigned char* __restrict__ kstar, float* __restrict__ w,
                        float* __restrict__ s, float* __restrict__ hub_ws) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long n_items = dl_num_items(g);
  const long long D = (long long)K * d;
  float e[DL_MAX_K], a[DL_MAX_K], sacc[DL_MAX_K];
  for (long long t = warp0; t < n_items; t += nwarps) {
    const DlItem it = dl_decode_item(g, t);
    for (int k = 0; k < K; ++k) sacc[k] = 0.0f;
    const float* zi = Z + it.node * D;
    for (long long p = it.e0; p < it.e1; ++p) {
      const float* zj = Z + (long long)__ldg(g.col + p) * D;
      int ks = dl_generic_route(zi, zj, K, d, T, lane, e, a);
      float wv = a[ks];
      sacc[ks] = __fadd_rn(sacc[ks], wv);
      if (lane == 0) { kstar[p] = (unsigned char)ks; w[p] = wv; }
    }
    if (lane == 0) {
      for (int k = 0; k < K; ++k) {
        if (it.hub_slot >= 0) hub_ws[it.hub_slot * K + k] = sacc[k];
        else s[it.node * K + k] = (sacc[k] == 0.0f) ? 1.0f : sacc[k];
      }
    }
  }
}

// hub rows: s[row,k] = sum over the row's segments (in order) of the partials, zeros -> 1
__global__ void k_attn_hub_fixup(DlGraphDev g, int K, const float* __restrict__ hub_ws,
                                 float* __restrict__ s) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < g.n_hub * K; x += stride) {
    long long h = x / K;
    int k = (int)(x % K);
    long long a = g.hub_seg_ptr[h], b = g.hub_seg_ptr[h + 1];
    float v = 0.0f;
    for (long long sg = a; sg < b; ++sg) v = __fadd_rn(v, hub_ws[sg * K + k]);
    s[(g.row_base + g.perm[h]) * K + k] = (v == 0.0f) ? 1.0f : v;
  }
}

// aggregation, runtime-generic path.  Direct rows accumulate in H itself (one warp owns the row),
// hub segments in their scratch slot.
__global__ void __launch_bounds__(DL_CTA)
k_factor_spmm_fwd_generic(DlGraphDev g, const float* __restrict__ Z,
                          const unsigned char* __restrict__ kstar, const float* __restrict__ w,
                          const float* __restrict__ s, int K, int d, float beta, float omb,
                          float* __restrict__ H, float* __restrict__ hub_ws) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long n_items = dl_num_items(g);
  const long long D = (long long)K * d;
  for (long long t = warp0; t < n_items; t += nwarps) {
    const DlItem it = dl_decode_item(g, t);
    float* acc = it.hub_slot >= 0 ? hub_ws + it.hub_slot * D : H + it.node * D;
    for (long long x = lane; x < D; x += 32) acc[x] = 0.0f;
    __syncwarp();
    for (long long p = it.e0; p < it.e1; ++p) {
      const long long j = __ldg(g.col + p);
      const int k = __ldg(kstar + p);
      const float coef = __fdiv_rn(__ldg(w + p), __ldg(s + j * K + k));
      const float* zj = Z + j * D + (long long)k * d;
      float* ak = acc + (long long)k * d;
      for (int x = lane; x < d; x += 32) ak[x] = __fmaf_rn(coef, zj[x], ak[x]);
    }
    __syncwarp();
    if (it.hub_slot < 0) {
      const float* zi = Z + it.node * D;
      for (long long x = lane; x < D; x += 32)
        acc[x] = __fadd_rn(__fmul_rn(beta, zi[x]), __fmul_rn(omb, acc[x]));
    }
  }
}

// hub rows: H[row] = beta z + (1-beta) * (sum of segment partials, in order)
__global__ void k_spmm_hub_fixup(DlGraphDev g, long long D, const float* __restrict__ Z,
                                 const float* __restrict__ hub_ws, float beta, float omb,
                                 float* __restrict__ H) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < g.n_hub * D; x += stride) {
    long long h = x / D, o = x % D;
    long long a = g.hub_seg_ptr[h], b = g.hub_seg_ptr[h + 1];
    float v = 0.0f;
    for (long long sg = a; sg < b; ++sg) v = __fadd_rn(v, hub_ws[sg * D + o]);
    long long row = g.row_base + g.perm[h];
    H[row * D + o] = __fadd_rn(__fmul_rn(beta, Z[row * D + o]), __fmul_rn(omb, v));
  }
}

template <class M>
int launch_attn(const DlGraphDev& g, long long n_items, const float* Z, float T, uint8_t* kstar,
                float* w, float* s, float* hub_ws, cudaStream_t st) {
  int grid = 1;
  int rc = dl_grid_for(k_edge_attn_fwd<M>, n_items, &grid, 0, DL_CTA_S);
  if (rc) return rc;
  k_edge_attn_fwd<M><<<grid, DL_CTA_S, 0, st>>>(g, Z, T, kstar, w, s, hub_ws);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

inline int fixup_blocks(long long n) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

}  // namespace

// Zs[i,k,:] = Z[i,k,:] * (1 / s[i,k]): one streaming pass (8*N*D bytes at full bandwidth) that takes the
// s[col,k] gather -- one of the two DRAM transactions per entry -- out of the aggregation kernel
static __global__ void __launch_bounds__(256)
k_scale_rows(const float* __restrict__ Z, const float* __restrict__ s, long long n_rows, int K, int d,
             float* __restrict__ Zs) {
  // one warp per row, lanes over the row's float4 chunks; s[row,k] is inverted once per chunk
  const int lane = threadIdx.x & 31;
  const int D4 = K * d / 4, d4 = d / 4;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows; row += nwarps) {
    const float4* zr = reinterpret_cast<const float4*>(Z) + row * D4;
    float4* out = reinterpret_cast<float4*>(Zs) + row * D4;
    for (int c = lane; c < D4; c += 32) {
      const float rs = __fdiv_rn(1.0f, __ldg(s + row * K + c / d4));
      float4 v = __ldg(zr + c);
      v.x = __fmul_rn(v.x, rs); v.y = __fmul_rn(v.y, rs); v.z = __fmul_rn(v.z, rs); v.w = __fmul_rn(v.w, rs);
      out[c] = v;
    }
  }
}

// per-entry copy of s[col, kstar] (sj_out) for the paths that do not write it on the way
static __global__ void k_entry_sj(DlGraphDev g, const unsigned char* __restrict__ kstar, const float* __restrict__ s,
                           int K, float* __restrict__ sj) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < g.nnz; e += stride)
    sj[e] = __ldg(s + (long long)__ldg(g.col + e) * K + __ldg(kstar + e));
}

extern "C" {

size_t dl_hub_scratch_floats(const dl_graph* g_host, int64_t width) {
  if (!g_host || width < 0) return 0;
  // hub-segment partials of the row-per-warp kernels, or range carries + chain scratch of the
  // streaming kernels (3 records per 2048-entry range), whichever is larger
  const size_t hub = (size_t)g_host->n_hub_items * (size_t)width;
  const long long RE = 32LL << dl_range_shift(g_host->nnz);      // entries per range of the streaming kernels
  const size_t stream = (size_t)((g_host->nnz + RE - 1) / RE) * 3 * (size_t)width;
  return hub > stream ? hub : stream;
}

static int attn_fwd_impl(const dl_graph* g_host, const float* Z, int K, int d, float T,
                         uint8_t* kstar, float* w, float* s, float* hub_ws, float* const* s_peers, int n_peers,
                         dl_stream_t stream) {
  if (!dl_graph_ok(g_host) || !dl_shape_ok(K, d)) return DL_EINVAL;
  if (g_host->N == 0) return DL_OK;
  if (!Z || !s || (g_host->nnz > 0 && (!kstar || !w))) return DL_EINVAL;
  if (g_host->n_hub_items > 0 && !hub_ws) return DL_EINVAL;
  if (!(T == T) || T == 0.0f) return DL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned flags = g_host->flags;
  const bool no_stream = (flags & DL_F_NO_STREAM) != 0;
  DlGraphDev g = dl_graph_dev(g_host);
  if (!dl_set_peer_out(g, s_peers, n_peers)) return DL_EINVAL;
  const long long n_items = g.n_hub_items + (g.N - g.n_hub);
  int rc = -1000;
  bool stream_tried = false;
  // factor-per-lane kernel (attn_fl.cu): routing and row sums in one launch
  if (g.erow && g.nnz > 0 && !(flags & (DL_F_NO_STREAM | DL_F_NO_FL | DL_F_NO_FL_ATTN))) {
    rc = dl_launch_attn_fl(g, Z, K, d, T, kstar, w, s, hub_ws, st);
    if (rc == DL_OK) return DL_OK;
    if (rc != -1000) return rc;
  }
  if (g.erow && g.nnz > 0 && !no_stream) {
    rc = dl_launch_attn_stream(g, g.erow, Z, K, d, T, kstar, w, s, hub_ws, st);
    stream_tried = (rc != -1000);
  }
  if (rc == -1000) {
#define BODY_MACRO(M) rc = launch_attn<M>(g, n_items, Z, T, kstar, w, s, hub_ws, st);
    DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  }
  if (rc == -1000) {
    int grid = 1;
    rc = dl_grid_for(k_edge_attn_fwd_generic, n_items, &grid);
    if (rc) return rc;
    k_edge_attn_fwd_generic<<<grid, DL_CTA, 0, st>>>(g, Z, K, d, T, kstar, w, s, hub_ws);
    DL_LAUNCH_CHECK();
    rc = DL_OK;
  }
  bool streamed = false;
  if (rc == -1001) { rc = DL_OK; }           // streamed routing, row sums by the row-per-warp kernel
  else if (rc == DL_OK && g.erow && g.nnz > 0 && !no_stream && stream_tried) streamed = true;
  if (rc) return rc;
  if (g.n_hub > 0 && !streamed) {
    k_attn_hub_fixup<<<fixup_blocks(g.n_hub * K), 256, 0, st>>>(g, K, hub_ws, s);
    DL_LAUNCH_CHECK();
  }
  if (n_peers > 0 && !streamed) {              // the row-per-warp row sums do not push: one copy kernel does
    void* dst[DL_MAX_PEER_OUT];
    const long long off = g.row_base * (long long)K;
    for (int q = 0; q < n_peers; ++q) dst[q] = s_peers[q] + off;
    return dl_push_slice(s + off, dst, n_peers, (int64_t)g.N * K * 4, stream);
  }
  return DL_OK;
}

int dl_edge_attn_fwd(const dl_graph* g_host, const float* Z, int K, int d, float T,
                     uint8_t* kstar, float* w, float* s, float* hub_ws, dl_stream_t stream) {
  return attn_fwd_impl(g_host, Z, K, d, T, kstar, w, s, hub_ws, nullptr, 0, stream);
}

int dl_edge_attn_fwd_push(const dl_graph* g_host, const float* Z, int K, int d, float T,
                          uint8_t* kstar, float* w, float* s, float* hub_ws, float* const* s_peers, int n_peers,
                          dl_stream_t stream) {
  return attn_fwd_impl(g_host, Z, K, d, T, kstar, w, s, hub_ws, s_peers, n_peers, stream);
}

static int spmm_fwd_impl(const dl_graph* g_host, const float* Z, const uint8_t* kstar,
                         const float* w, const float* s, int K, int d, float beta,
                         float one_minus_beta, float* H, float* sj_out, float* zs_scratch, int64_t zs_rows,
                         float* hub_ws, float* const* H_peers, int n_peers, dl_stream_t stream) {
  if (!dl_graph_ok(g_host) || !dl_shape_ok(K, d)) return DL_EINVAL;
  if (g_host->N == 0) return DL_OK;
  if (!Z || !s || !H || (g_host->nnz > 0 && (!kstar || !w))) return DL_EINVAL;
  if (g_host->n_hub_items > 0 && !hub_ws) return DL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned flags = g_host->flags;
  DlGraphDev g = dl_graph_dev(g_host);
  if (!dl_set_peer_out(g, H_peers, n_peers)) return DL_EINVAL;
  const long long n_items = g.n_hub_items + (g.N - g.n_hub);
  int rc = -1000;
  // pre-scaled path: needs every gathered row inside [0, N) (a graph that is not row-partitioned)
  if (zs_scratch && (sj_out || g.row_base != 0)) return DL_EINVAL;
  if (zs_scratch && g.nnz > 0 && d % 4 == 0 && g.erow && !(flags & (DL_F_NO_STREAM | DL_F_NO_PRESCALE))) {
    k_scale_rows<<<148 * 16, 256, 0, st>>>(Z, s, zs_rows > g.N ? zs_rows : g.N, K, d, zs_scratch);
    DL_LAUNCH_CHECK();
    rc = dl_launch_gather_stream(0, g, Z, zs_scratch, kstar, w, nullptr, K, d, beta, one_minus_beta, H, nullptr,
                                 hub_ws, st);
    if (rc == DL_OK) return DL_OK;
    if (rc != -1000) return rc;                // -1000: no streaming instantiation, the scratch goes unused
  }
  if (!(flags & DL_F_NO_STREAM))
    rc = dl_launch_gather_stream(0, g, Z, Z, kstar, w, s, K, d, beta, one_minus_beta, H, sj_out, hub_ws, st);
  if (rc == DL_OK) return DL_OK;             // carries, chained rows and empty rows all handled
  if (rc != -1000) return rc;
  if (sj_out && g.nnz > 0) {                 // the other paths do not produce sj on the way
    k_entry_sj<<<fixup_blocks(g.nnz), 256, 0, st>>>(g, kstar, s, K, sj_out);
    DL_LAUNCH_CHECK();
  }
  rc = dl_launch_slice_gather(0, g, n_items, Z, Z, kstar, w, s, K, d, beta, one_minus_beta, H, nullptr,
                              hub_ws, st);
  if (rc == -1000) {
    int grid = 1;
    rc = dl_grid_for(k_factor_spmm_fwd_generic, n_items, &grid);
    if (rc) return rc;
    k_factor_spmm_fwd_generic<<<grid, DL_CTA, 0, st>>>(g, Z, kstar, w, s, K, d, beta, one_minus_beta,
                                                       H, hub_ws);
    DL_LAUNCH_CHECK();
    rc = DL_OK;
  }
  if (rc) return rc;
  if (g.n_hub > 0) {
    const long long D = (long long)K * d;
    k_spmm_hub_fixup<<<fixup_blocks(g.n_hub * D), 256, 0, st>>>(g, D, Z, hub_ws, beta, one_minus_beta, H);
    DL_LAUNCH_CHECK();
  }
  if (n_peers > 0) {                           // the row-per-warp paths do not push: one copy kernel does
    void* dst[DL_MAX_PEER_OUT];
    const long long off = g.row_base * (long long)K * d;
    for (int q = 0; q < n_peers; ++q) dst[q] = H_peers[q] + off;
    return dl_push_slice(H + off, dst, n_peers, (int64_t)g.N * K * d * 4, stream);
  }
  return DL_OK;
}

int dl_factor_spmm_fwd(const dl_graph* g_host, const float* Z, const uint8_t* kstar,
                       const float* w, const float* s, int K, int d, float beta,
                       float one_minus_beta, float* H, float* sj_out, float* zs_scratch, int64_t zs_rows,
                       float* hub_ws, dl_stream_t stream) {
  return spmm_fwd_impl(g_host, Z, kstar, w, s, K, d, beta, one_minus_beta, H, sj_out, zs_scratch, zs_rows, hub_ws,
                       nullptr, 0, stream);
}

int dl_factor_spmm_fwd_push(const dl_graph* g_host, const float* Z, const uint8_t* kstar,
                            const float* w, const float* s, int K, int d, float beta,
                            float one_minus_beta, float* H, float* sj_out, float* hub_ws,
                            float* const* H_peers, int n_peers, dl_stream_t stream) {
  return spmm_fwd_impl(g_host, Z, kstar, w, s, K, d, beta, one_minus_beta, H, sj_out, nullptr, 0, hub_ws, H_peers,
                       n_peers, stream);
}

}  // extern "C"