// pair_score.cu -- factor-weighted link-pair scoring over explicit (u,v) batches, and its backward.
//
//   [ref: model.py:109-113]  link_pred = sigmoid( sum_k exp(Z_k Z_k^T / T)[u,v] * (H_k H_k^T)[u,v] )
//   The reference evaluates all N^2 pairs and the training script reads a few thousand of them
//   through boolean masks (main_disentangled.py:195,202,217); here only the listed pairs are
//   evaluated: 4 row gathers (z_u, z_v, h_u, h_v) and one float out per pair.
//
//   Backward is node-major over incidence lists (dl_pair_incidence) so every dZ / dH row is written
//   by exactly one warp (or reduced from hub segments in a fixed order): no atomics, deterministic.
//
// HBM bytes per pair (D=128): forward 8 (ids) + 4*512 + 4; backward 2 * (4 + 4 + 4 + 2*512) per
// pair (each pair is visited from both endpoints) + per node 2*512 in, 2*512 out.
#include <stdlib.h>

#include "dl_dispatch.cuh"

namespace {

template <class M>
__global__ void __launch_bounds__(DL_CTA)
k_pair_score_fwd(const int* __restrict__ u, const int* __restrict__ v, long long P,
                 const float* __restrict__ Z, const float* __restrict__ H, float T,
                 float* __restrict__ logit, float* __restrict__ prob) {
  constexpr int D = M::D, NP = M::NP, EB = M::EB, LP = M::LP;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long nblk = (P + 31) / 32;
  const int my_e = M::edge_of_lane(lane);

  for (long long b = warp0; b < nblk; b += nwarps) {
    const long long base = b * 32;
    const int cnt = (int)min(32LL, P - base);
    const int myu = lane < cnt ? __ldg(u + base + lane) : 0;
    const int myv = lane < cnt ? __ldg(v + base + lane) : 0;
    float outS = 0.0f;
    const int nsub = (cnt + EB - 1) / EB;
    for (int sb = 0; sb < nsub; ++sb) {
      int cu[EB], cv[EB];
      bool valid[EB];
#pragma unroll
      for (int e = 0; e < EB; ++e) {
        const int idx = sb * EB + e;
        cu[e] = __shfl_sync(DL_FULL, myu, idx & 31);
        cv[e] = __shfl_sync(DL_FULL, myv, idx & 31);
        valid[e] = idx < cnt;
      }
      float S = 0.0f;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const bool act = M::active(lane, p);
        const int off = M::offset(lane, p);
        float4 zu[EB], zv[EB], hu[EB], hv[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          const bool ld = valid[e] && act;
          zu[e] = ld ? dl_ldg4(Z + (long long)cu[e] * D + off) : dl_zero4();
          zv[e] = ld ? dl_ldg4(Z + (long long)cv[e] * D + off) : dl_zero4();
          hu[e] = ld ? dl_ldg4(H + (long long)cu[e] * D + off) : dl_zero4();
          hv[e] = ld ? dl_ldg4(H + (long long)cv[e] * D + off) : dl_zero4();
        }
        float pz[EB], ph[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) { pz[e] = dl_chunk_dot(zu[e], zv[e]); ph[e] = dl_chunk_dot(hu[e], hv[e]); }
        const float q = __fdiv_rn(dl_reduce_scatter<M>(pz, lane), T);
        const float gh = dl_reduce_scatter<M>(ph, lane);
        // idle factor slots: exp(0) * 0 = 0
        const float tv = __fmul_rn(dl_expf(q), gh);
        S = (p == 0) ? tv : __fadd_rn(S, tv);
      }
      // sum over the factor slots of the warp (butterfly across groups)
#pragma unroll
      for (int o = LP; o < 32; o <<= 1) S = __fadd_rn(S, __shfl_xor_sync(DL_FULL, S, o));
      if (sb == lane / EB) outS = S;
    }
    const int oi = (lane & ~(EB - 1)) + my_e;
    if (oi < cnt) {
      if (logit) logit[base + oi] = outS;
      if (prob) prob[base + oi] = dl_sigmoid(outS);
    }
  }
}

__global__ void __launch_bounds__(DL_CTA)
k_pair_score_fwd_generic(const int* __restrict__ u, const int* __restrict__ v, long long P,
                         const float* __restrict__ Z, const float* __restrict__ H, int K, int d,
                         float T, float* __restrict__ logit, float* __restrict__ prob) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long D = (long long)K * d;
  for (long long p = warp0; p < P; p += nwarps) {
    const float* zu = Z + (long long)u[p] * D;
    const float* zv = Z + (long long)v[p] * D;
    const float* hu = H + (long long)u[p] * D;
    const float* hv = H + (long long)v[p] * D;
    float S = 0.0f;
    for (int k = 0; k < K; ++k) {
      float q = __fdiv_rn(dl_generic_dot(zu + k * d, zv + k * d, d, lane), T);
      float gh = dl_generic_dot(hu + k * d, hv + k * d, d, lane);
      float tv = __fmul_rn(dl_expf(q), gh);
      S = (k == 0) ? tv : __fadd_rn(S, tv);
    }
    if (lane == 0) {
      if (logit) logit[p] = S;
      if (prob) prob[p] = dl_sigmoid(S);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, node-major.  g is the incidence "graph": rowptr = inc_ptr, col = inc_other.
//   dH[n,k] = sum_t dS e_k H[o,k] ;  dZ[n,k] = sum_t dS e_k <H[n,k],H[o,k]> / T  Z[o,k]
// ---------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(DL_CTA)
k_pair_score_bwd(DlGraphDev g, const int* __restrict__ inc_pair, const float* __restrict__ Z,
                 const float* __restrict__ H, const float* __restrict__ dS, float T,
                 float* __restrict__ dZ, float* __restrict__ dH, float* __restrict__ hub_ws) {
  constexpr int D = M::D, NP = M::NP, EB = M::EB, LP = M::LP;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long n_items = dl_num_items(g);
  const int my_e = M::edge_of_lane(lane);
  const int gbase = lane & ~(LP - 1);

  for (long long t = warp0; t < n_items; t += nwarps) {
    const DlItem it = dl_decode_item(g, t);
    float4 zn[NP], hn[NP], az[NP], ah[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const bool act = M::active(lane, p);
      zn[p] = act ? dl_ldg4(Z + it.node * D + M::offset(lane, p)) : dl_zero4();
      hn[p] = act ? dl_ldg4(H + it.node * D + M::offset(lane, p)) : dl_zero4();
      az[p] = dl_zero4();
      ah[p] = dl_zero4();
    }
    for (long long base = it.e0; base < it.e1; base += 32) {
      const int cnt = (int)min(32LL, it.e1 - base);
      int myo = 0;
      float myds = 0.0f;
      if (lane < cnt) {
        myo = __ldg(g.col + base + lane);
        myds = __ldg(dS + __ldg(inc_pair + base + lane));
      }
      const int nsub = (cnt + EB - 1) / EB;
      for (int sb = 0; sb < nsub; ++sb) {
        int co[EB];
        bool valid[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          const int idx = sb * EB + e;
          co[e] = __shfl_sync(DL_FULL, myo, idx & 31);
          valid[e] = idx < cnt;
        }
        const int my_idx = sb * EB + my_e;
        const float ds = __shfl_sync(DL_FULL, myds, my_idx & 31);  // 0 beyond cnt
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const bool act = M::active(lane, p);
          const int off = M::offset(lane, p);
          float4 zo[EB], ho[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            const bool ld = valid[e] && act;
            zo[e] = ld ? dl_ldg4(Z + (long long)co[e] * D + off) : dl_zero4();
            ho[e] = ld ? dl_ldg4(H + (long long)co[e] * D + off) : dl_zero4();
          }
          float pz[EB], ph[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) { pz[e] = dl_chunk_dot(zn[p], zo[e]); ph[e] = dl_chunk_dot(hn[p], ho[e]); }
          const float q = __fdiv_rn(dl_reduce_scatter<M>(pz, lane), T);
          const float hh = dl_reduce_scatter<M>(ph, lane);
          const float ek = dl_expf(q);
          const float ch = __fmul_rn(ds, ek);
          const float cz = __fdiv_rn(__fmul_rn(__fmul_rn(ds, ek), hh), T);
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            const int src = gbase + M::lane_of_edge(e);
            const float che = __shfl_sync(DL_FULL, ch, src);
            const float cze = __shfl_sync(DL_FULL, cz, src);
            dl_fma4(ah[p], che, ho[e]);
            dl_fma4(az[p], cze, zo[e]);
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (!M::active(lane, p)) continue;
      const int o = M::offset(lane, p);
      if (it.hub_slot >= 0) {
        *reinterpret_cast<float4*>(hub_ws + it.hub_slot * 2 * D + o) = az[p];
        *reinterpret_cast<float4*>(hub_ws + it.hub_slot * 2 * D + D + o) = ah[p];
      } else {
        *reinterpret_cast<float4*>(dZ + it.node * D + o) = az[p];
        *reinterpret_cast<float4*>(dH + it.node * D + o) = ah[p];
      }
    }
  }
}

__global__ void __launch_bounds__(DL_CTA)
k_pair_score_bwd_generic(DlGraphDev g, const int* __restrict__ inc_pair, const float* __restrict__ Z,
                         const float* __restrict__ H, const float* __restrict__ dS, int K, int d,
                         float T, float* __restrict__ dZ, float* __restrict__ dH,
                         float* __restrict__ hub_ws) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long n_items = dl_num_items(g);
  const long long D = (long long)K * d;
  for (long long t = warp0; t < n_items; t += nwarps) {
    const DlItem it = dl_decode_item(g, t);
    float* az = it.hub_slot >= 0 ? hub_ws + it.hub_slot * 2 * D : dZ + it.node * D;
    float* ah = it.hub_slot >= 0 ? hub_ws + it.hub_slot * 2 * D + D : dH + it.node * D;
    for (long long x = lane; x < D; x += 32) { az[x] = 0.0f; ah[x] = 0.0f; }
    __syncwarp();
    const float* zn = Z + it.node * D;
    const float* hn = H + it.node * D;
    for (long long e = it.e0; e < it.e1; ++e) {
      const long long o = __ldg(g.col + e);
      const float ds = __ldg(dS + __ldg(inc_pair + e));
      const float* zo = Z + o * D;
      const float* ho = H + o * D;
      for (int k = 0; k < K; ++k) {
        float q = __fdiv_rn(dl_generic_dot(zn + k * d, zo + k * d, d, lane), T);
        float hh = dl_generic_dot(hn + k * d, ho + k * d, d, lane);
        float ek = dl_expf(q);
        float ch = __fmul_rn(ds, ek);
        float cz = __fdiv_rn(__fmul_rn(__fmul_rn(ds, ek), hh), T);
        for (int x = lane; x < d; x += 32) {
          ah[k * d + x] = __fmaf_rn(ch, ho[k * d + x], ah[k * d + x]);
          az[k * d + x] = __fmaf_rn(cz, zo[k * d + x], az[k * d + x]);
        }
      }
    }
    __syncwarp();
  }
}

__global__ void k_pair_bwd_hub_fixup(DlGraphDev g, long long D, const float* __restrict__ hub_ws,
                                     float* __restrict__ dZ, float* __restrict__ dH) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < g.n_hub * D; x += stride) {
    long long h = x / D, o = x % D;
    long long a = g.hub_seg_ptr[h], b = g.hub_seg_ptr[h + 1];
    float vz = 0.0f, vh = 0.0f;
    for (long long sg = a; sg < b; ++sg) {
      vz = __fadd_rn(vz, hub_ws[sg * 2 * D + o]);
      vh = __fadd_rn(vh, hub_ws[sg * 2 * D + D + o]);
    }
    long long row = g.row_base + g.perm[h];
    dZ[row * D + o] = vz;
    dH[row * D + o] = vh;
  }
}

template <class M>
int launch_pair_fwd(const int* u, const int* v, long long P, const float* Z, const float* H, float T,
                    float* logit, float* prob, cudaStream_t st) {
  int grid = 1;
  int rc = dl_grid_for(k_pair_score_fwd<M>, (P + 31) / 32, &grid);
  if (rc) return rc;
  k_pair_score_fwd<M><<<grid, DL_CTA, 0, st>>>(u, v, P, Z, H, T, logit, prob);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

template <class M>
int launch_pair_bwd(const DlGraphDev& g, long long n_items, const int* inc_pair, const float* Z,
                    const float* H, const float* dS, float T, float* dZ, float* dH, float* hub_ws,
                    cudaStream_t st) {
  int grid = 1;
  int rc = dl_grid_for(k_pair_score_bwd<M>, n_items, &grid);
  if (rc) return rc;
  k_pair_score_bwd<M><<<grid, DL_CTA, 0, st>>>(g, inc_pair, Z, H, dS, T, dZ, dH, hub_ws);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

}  // namespace

extern "C" {

int dl_pair_score_fwd(const int32_t* u, const int32_t* v, int64_t P, const float* Z, const float* H,
                      int64_t N, int K, int d, float T, float* logit, float* prob,
                      dl_stream_t stream) {
  if (P < 0 || N < 0 || !dl_shape_ok(K, d)) return DL_EINVAL;
  if (P == 0) return DL_OK;
  if (!u || !v || !Z || !H || (!logit && !prob)) return DL_EINVAL;
  if (!(T == T) || T == 0.0f) return DL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = -1000;
#define BODY_MACRO(M) rc = launch_pair_fwd<M>(u, v, P, Z, H, T, logit, prob, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  if (rc == -1000) {
    int grid = 1;
    rc = dl_grid_for(k_pair_score_fwd_generic, P, &grid);
    if (rc) return rc;
    k_pair_score_fwd_generic<<<grid, DL_CTA, 0, st>>>(u, v, P, Z, H, K, d, T, logit, prob);
    DL_LAUNCH_CHECK();
    rc = DL_OK;
  }
  return rc;
}

static int pair_bwd_impl(const dl_graph* inc_host, const int32_t* inc_pair, const float* Z,
                         const float* H, const float* dS, int K, int d, float T, float* dZ, float* dH,
                         float* hub_ws, float* const* dH_peers, int n_peers, dl_stream_t stream) {
  if (!dl_graph_ok(inc_host) || !dl_shape_ok(K, d)) return DL_EINVAL;
  if (inc_host->N == 0) return DL_OK;
  if (!Z || !H || !dZ || !dH) return DL_EINVAL;
  if (inc_host->nnz > 0 && (!inc_pair || !dS)) return DL_EINVAL;
  if (inc_host->n_hub_items > 0 && !hub_ws) return DL_EINVAL;
  if (!(T == T) || T == 0.0f) return DL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  DlGraphDev g = dl_graph_dev(inc_host);
  if (!dl_set_peer_out(g, dH_peers, n_peers)) return DL_EINVAL;
  const long long n_items = g.n_hub_items + (g.N - g.n_hub);
  int rc = -1000;
  if (!(inc_host->flags & DL_F_NO_STREAM))
    rc = dl_launch_pair_bwd_stream(g, inc_pair, Z, H, dS, K, d, T, dZ, dH, hub_ws, st);
  if (rc == DL_OK) return DL_OK;
  if (rc != -1000) return rc;
#define BODY_MACRO(M) rc = launch_pair_bwd<M>(g, n_items, inc_pair, Z, H, dS, T, dZ, dH, hub_ws, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  if (rc == -1000) {
    int grid = 1;
    rc = dl_grid_for(k_pair_score_bwd_generic, n_items, &grid);
    if (rc) return rc;
    k_pair_score_bwd_generic<<<grid, DL_CTA, 0, st>>>(g, inc_pair, Z, H, dS, K, d, T, dZ, dH, hub_ws);
    DL_LAUNCH_CHECK();
    rc = DL_OK;
  }
  if (rc) return rc;
  if (g.n_hub > 0) {
    const long long D = (long long)K * d;
    long long n = g.n_hub * D;
    long long b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    k_pair_bwd_hub_fixup<<<(int)b, 256, 0, st>>>(g, D, hub_ws, dZ, dH);
    DL_LAUNCH_CHECK();
  }
  if (n_peers > 0) {                           // the row-per-warp paths do not push: one copy kernel does
    void* dst[DL_MAX_PEER_OUT];
    const long long off = g.row_base * (long long)K * d;
    for (int q = 0; q < n_peers; ++q) dst[q] = dH_peers[q] + off;
    return dl_push_slice(dH + off, dst, n_peers, (int64_t)g.N * K * d * 4, stream);
  }
  return DL_OK;
}

int dl_pair_score_bwd(const dl_graph* inc_host, const int32_t* inc_pair, const float* Z,
                      const float* H, const float* dS, int K, int d, float T, float* dZ, float* dH,
                      float* hub_ws, dl_stream_t stream) {
  return pair_bwd_impl(inc_host, inc_pair, Z, H, dS, K, d, T, dZ, dH, hub_ws, nullptr, 0, stream);
}

int dl_pair_score_bwd_push(const dl_graph* inc_host, const int32_t* inc_pair, const float* Z,
                           const float* H, const float* dS, int K, int d, float T, float* dZ, float* dH,
                           float* hub_ws, float* const* dH_peers, int n_peers, dl_stream_t stream) {
  return pair_bwd_impl(inc_host, inc_pair, Z, H, dS, K, d, T, dZ, dH, hub_ws, dH_peers, n_peers, stream);
}

}  // extern "C"