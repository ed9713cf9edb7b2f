// factor_fwd.cu -- forward of DisenLink's factor-aware message passing on a CSR.
//
//   k_edge_attn_fwd   [ref: model.py:56-73]  per entry (i,j): q_k = z_i^k.z_j^k / T, softmax over
//                     the K factors, hard routing kstar = first argmax, w = a[kstar]; per row the
//                     routed sums s[i,k] (zeros -> 1).  Everything stays in registers; only
//                     (kstar: u8, w: f32) per entry and s [N,K] are written.
//   k_factor_spmm_fwd [ref: model.py:75]     H[i,k] = beta Z[i,k] + (1-beta) sum_{j: kstar=k}
//                     (w_ij / s[j,k]) Z[j,k]  -- gather / segment-sum in column order, no atomics.
//
// Mapping (fast path, DlMap<K,d>): one warp per work item (a row, or a DL_SEG-edge segment of a
// hub row).  A row of D = K*d floats is spread over the lanes as float4 chunks, so a neighbour
// row is one coalesced 128-bit-per-lane gather.  Four edges are in flight per warp step; their
// chunk partials are reduce-scattered over the d/4-lane factor group (3 shuffles per 4 edges
// instead of 8), giving lane (k, g) the finished dot of edge g for factor k.
//
// HBM bytes per entry (D=128, K=8, d=16): attention 4 (col) + 512 (z_j) + 5 (kstar, w);
// aggregation 4 + 5 + 4 (s[j,k]) + 64 (z_j^k slice).  Per row: z_i, s / H.  DESIGN.md section 4.
#include "dl_dispatch.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// attention, fast path
// ---------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(DL_CTA)
k_edge_attn_fwd(DlGraphDev g, const float* __restrict__ Z, float T, unsigned char* __restrict__ kstar,
                float* __restrict__ w, float* __restrict__ s, float* __restrict__ hub_ws) {
  constexpr int K = M::K, D = M::D, NP = M::NP, EB = M::EB, LP = M::LP, FPP = M::FPP;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long n_items = dl_num_items(g);

  int off[NP];
  bool act[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { off[p] = M::offset(lane, p); act[p] = M::active(lane, p); }
  const int my_e = M::edge_of_lane(lane);
  const int gsrc = lane & (EB - 1);
  const bool primary = (lane % LP) < EB;  // one replica per (factor, edge) accumulates s

  for (long long t = warp0; t < n_items; t += nwarps) {
    const DlItem it = dl_decode_item(g, t);
    float4 zi[NP];
    float sacc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      zi[p] = act[p] ? dl_ldg4(Z + it.node * D + off[p]) : dl_zero4();
      sacc[p] = 0.0f;
    }
    for (long long base = it.e0; base < it.e1; base += 32) {
      const int cnt = (int)min(32LL, it.e1 - base);
      const int mycol = lane < cnt ? __ldg(g.col + base + lane) : 0;
      int out_ks = 0;
      float out_w = 0.0f;
      const int nsub = (cnt + EB - 1) / EB;
      for (int sb = 0; sb < nsub; ++sb) {
        float4 zj[EB][NP];
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          const int idx = sb * EB + e;
          const int c = __shfl_sync(DL_FULL, mycol, idx & 31);
          const bool valid = idx < cnt;
#pragma unroll
          for (int p = 0; p < NP; ++p)
            zj[e][p] = (valid && act[p]) ? dl_ldg4(Z + (long long)c * D + off[p]) : dl_zero4();
        }
        float ev[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          float part[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) part[e] = dl_chunk_dot(zi[p], zj[e][p]);
          float q = __fdiv_rn(dl_reduce_scatter<M>(part, lane), T);
          ev[p] = dl_expf(q);
        }
        // all-gather the K exponentials of my edge, then the softmax / argmax in registers
        float a[K];
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          a[k] = __shfl_sync(DL_FULL, ev[k / FPP], (k % FPP) * LP + gsrc);
          sum = (k == 0) ? a[0] : __fadd_rn(sum, a[k]);
        }
        int ks = 0;
        float wv = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float v = __fdiv_rn(a[k], sum);
          if (k == 0) { wv = v; }
          else if (v > wv || (v != v && wv == wv)) { wv = v; ks = k; }
        }
        const bool valid = (sb * EB + my_e) < cnt;
#pragma unroll
        for (int p = 0; p < NP; ++p)
          if (valid && primary && ks == M::factor(lane, p)) sacc[p] = __fadd_rn(sacc[p], wv);
        if (sb == lane / EB) { out_ks = ks; out_w = wv; }
      }
      const int oi = (lane & ~(EB - 1)) + my_e;
      if (oi < cnt) {
        kstar[base + oi] = (unsigned char)out_ks;
        w[base + oi] = out_w;
      }
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

// ---------------------------------------------------------------------------------------------
// aggregation, fast path
// ---------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(DL_CTA)
k_factor_spmm_fwd(DlGraphDev g, const float* __restrict__ Z, const unsigned char* __restrict__ kstar,
                  const float* __restrict__ w, const float* __restrict__ s, float beta, float omb,
                  float* __restrict__ H, float* __restrict__ hub_ws) {
  constexpr int K = M::K, d = M::d, D = M::D, NP = M::NP, L = M::L, FPP = M::FPP;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long n_items = dl_num_items(g);
  const int slot = M::slot(lane), gg = M::g(lane);
  const bool glane = gg < L;

  for (long long t = warp0; t < n_items; t += nwarps) {
    const DlItem it = dl_decode_item(g, t);
    float4 acc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] = dl_zero4();
    for (long long base = it.e0; base < it.e1; base += 32) {
      const int cnt = (int)min(32LL, it.e1 - base);
      int c = 0, k = 255;
      float coef = 0.0f;
      if (lane < cnt) {
        c = __ldg(g.col + base + lane);
        k = __ldg(kstar + base + lane);
        float wv = __ldg(w + base + lane);
        float sj = __ldg(s + (long long)c * K + k);
        coef = __fdiv_rn(wv, sj);
      }
      for (int i0 = 0; i0 < cnt; i0 += 8) {
        float4 z[8];
        float cf[8];
        int pk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = i0 + u;  // < 32; lanes >= cnt carry k = 255 and match nothing
          const int cc = __shfl_sync(DL_FULL, c, idx);
          const int kk = __shfl_sync(DL_FULL, k, idx);
          cf[u] = __shfl_sync(DL_FULL, coef, idx);
          const bool m = glane && (kk % FPP) == slot && kk < K;
          pk[u] = m ? kk / FPP : -1;
          z[u] = m ? dl_ldg4(Z + (long long)cc * D + kk * d + 4 * gg) : dl_zero4();
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
          for (int p = 0; p < NP; ++p)
            if (pk[u] == p) dl_fma4(acc[p], cf[u], z[u]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (!M::active(lane, p)) continue;
      const int o = M::offset(lane, p);
      if (it.hub_slot >= 0) {
        *reinterpret_cast<float4*>(hub_ws + it.hub_slot * D + o) = acc[p];
      } else {
        const float4 zi = dl_ldg4(Z + it.node * D + o);
        float4 h;
        h.x = __fadd_rn(__fmul_rn(beta, zi.x), __fmul_rn(omb, acc[p].x));
        h.y = __fadd_rn(__fmul_rn(beta, zi.y), __fmul_rn(omb, acc[p].y));
        h.z = __fadd_rn(__fmul_rn(beta, zi.z), __fmul_rn(omb, acc[p].z));
        h.w = __fadd_rn(__fmul_rn(beta, zi.w), __fmul_rn(omb, acc[p].w));
        *reinterpret_cast<float4*>(H + it.node * D + o) = h;
      }
    }
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
  int rc = dl_grid_for(k_edge_attn_fwd<M>, n_items, &grid);
  if (rc) return rc;
  k_edge_attn_fwd<M><<<grid, DL_CTA, 0, st>>>(g, Z, T, kstar, w, s, hub_ws);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

template <class M>
int launch_spmm(const DlGraphDev& g, long long n_items, const float* Z, const uint8_t* kstar,
                const float* w, const float* s, float beta, float omb, float* H, float* hub_ws,
                cudaStream_t st) {
  int grid = 1;
  int rc = dl_grid_for(k_factor_spmm_fwd<M>, n_items, &grid);
  if (rc) return rc;
  k_factor_spmm_fwd<M><<<grid, DL_CTA, 0, st>>>(g, Z, kstar, w, s, beta, omb, H, hub_ws);
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

extern "C" {

size_t dl_hub_scratch_floats(const dl_graph* g_host, int64_t width) {
  if (!g_host || width < 0) return 0;
  return (size_t)g_host->n_hub_items * (size_t)width;
}

int dl_edge_attn_fwd(const dl_graph* g_host, const float* Z, int K, int d, float T,
                     uint8_t* kstar, float* w, float* s, float* hub_ws, dl_stream_t stream) {
  if (!dl_graph_ok(g_host) || !dl_shape_ok(K, d)) return DL_EINVAL;
  if (g_host->N == 0) return DL_OK;
  if (!Z || !s || (g_host->nnz > 0 && (!kstar || !w))) return DL_EINVAL;
  if (g_host->n_hub_items > 0 && !hub_ws) return DL_EINVAL;
  if (!(T == T) || T == 0.0f) return DL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const DlGraphDev g = dl_graph_dev(g_host);
  const long long n_items = g.n_hub_items + (g.N - g.n_hub);
  int rc = -1000;
#define BODY_MACRO(M) rc = launch_attn<M>(g, n_items, Z, T, kstar, w, s, hub_ws, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  if (rc == -1000) {
    int grid = 1;
    rc = dl_grid_for(k_edge_attn_fwd_generic, n_items, &grid);
    if (rc) return rc;
    k_edge_attn_fwd_generic<<<grid, DL_CTA, 0, st>>>(g, Z, K, d, T, kstar, w, s, hub_ws);
    DL_LAUNCH_CHECK();
    rc = DL_OK;
  }
  if (rc) return rc;
  if (g.n_hub > 0) {
    k_attn_hub_fixup<<<fixup_blocks(g.n_hub * K), 256, 0, st>>>(g, K, hub_ws, s);
    DL_LAUNCH_CHECK();
  }
  return DL_OK;
}

int dl_factor_spmm_fwd(const dl_graph* g_host, const float* Z, const uint8_t* kstar,
                       const float* w, const float* s, int K, int d, float beta,
                       float one_minus_beta, float* H, float* hub_ws, dl_stream_t stream) {
  if (!dl_graph_ok(g_host) || !dl_shape_ok(K, d)) return DL_EINVAL;
  if (g_host->N == 0) return DL_OK;
  if (!Z || !s || !H || (g_host->nnz > 0 && (!kstar || !w))) return DL_EINVAL;
  if (g_host->n_hub_items > 0 && !hub_ws) return DL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const DlGraphDev g = dl_graph_dev(g_host);
  const long long n_items = g.n_hub_items + (g.N - g.n_hub);
  int rc = -1000;
#define BODY_MACRO(M) rc = launch_spmm<M>(g, n_items, Z, kstar, w, s, beta, one_minus_beta, H, hub_ws, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
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
  return DL_OK;
}

}  // extern "C"
