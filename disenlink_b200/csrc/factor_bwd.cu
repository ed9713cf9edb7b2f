// factor_bwd.cu -- backward of the factor attention + aggregation w.r.t. Z, given G = dL/dH.
//
// [ref: autograd of model.py:56-75].  Closed form, per CSR entry e=(i,j), k = kstar[e]:
//   pass 1 (k_factor_bwd_gather), per row i:
//     T_[i,k,:] = (1-beta)/s[i,k] * sum_{e in row i, kstar=k} w[e] * G[col_e,k,:]
//     r[i,k]    = <Z[i,k], T_[i,k]> / s[i,k]
//     dZ[i,:]  += beta * G[i,:] + T_[i,:]
//   pass 2 (k_factor_bwd_edges), per row i, per entry (i,j):
//     dwsum = c_ij/s[j,k] + c_ji/s[i,k] - r[i,k] - r[j,k],   c_ij = (1-beta) <G[i,k], Z[j,k]>
//     dZ[i,kk,:] += dwsum * w / T * ((kk==k) - a[kk]) * Z[j,kk,:]          for all kk
// The adjacency is symmetric and kstar / w / a are bitwise symmetric (canonical arithmetic,
// dl_common.cuh), so the term an entry sends to its column node is picked up by the mirrored entry
// in that node's own row: both passes are gathers, no scatter, no atomics.  a[] is recomputed from
// Z (the z_j row is needed anyway); kstar is read back so the G[j,k] slice, s[j,k] and r[j,k]
// gathers can be issued together with the z_j gather.
//
// HBM bytes per entry (D=128, d=16): pass 1  4 + 5 + 64;  pass 2  4 + 1 + 512 + 64 + 4 + 4.
#include <stdlib.h>

#include "dl_dispatch.cuh"

namespace {

struct EMeta {
  int c, kk;
  float sj, rj;
};

template <class M>
__global__ void __launch_bounds__(DL_CTA, (M::NP == 1 ? 2 : 1))
k_factor_bwd_edges(DlGraphDev g, const float* __restrict__ Z, const float* __restrict__ G,
                   const unsigned char* __restrict__ kstar, const float* __restrict__ s,
                   const float* __restrict__ r, float omb, float T, float* __restrict__ dZ,
                   float* __restrict__ hub_ws) {
  constexpr int K = M::K, d = M::d, D = M::D, NP = M::NP, EB = M::EB, LP = M::LP, FPP = M::FPP, L = M::L;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const int my_e = M::edge_of_lane(lane);
  const int gsrc = lane & (EB - 1);
  const int gbase = lane & ~(LP - 1);
  const int slot = M::slot(lane), gg = M::g(lane);
  const bool glane = gg < L;
  const bool unit_T = (T == 1.0f);

  auto load_ck = [&](const DlItem& it, EMeta& m) {
    m.c = 0; m.kk = 255; m.sj = 1.0f; m.rj = 0.0f;
    if (lane < it.e1 - it.e0) {
      m.c = __ldg(g.col + it.e0 + lane);
      m.kk = __ldg(kstar + it.e0 + lane);
    }
  };
  auto load_sr = [&](EMeta& m) {
    if (m.kk != 255) {
      m.sj = __ldg(s + (long long)m.c * K + m.kk);
      m.rj = __ldg(r + (long long)m.c * K + m.kk);
    }
  };
  auto load_row = [&](const DlItem& it, float4 (&zi)[NP], float4 (&gi)[NP]) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const bool act = M::active(lane, p);
      zi[p] = act ? dl_ldg4(Z + it.node * D + M::offset(lane, p)) : dl_zero4();
      gi[p] = act ? dl_ldg4(G + it.node * D + M::offset(lane, p)) : dl_zero4();
    }
  };

  DlRowIter itr;
  itr.init(g, warp0, nwarps);
  DlItem it0, it1, it2;
  EMeta m0, m1, m2;
  it0 = it1 = it2 = DlItem{0, 0, 0, 0, -1};
  m0 = m1 = m2 = EMeta{0, 255, 1.0f, 0.0f};
  float4 zi[NP], gi[NP], ziB[NP], giB[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) zi[p] = gi[p] = ziB[p] = giB[p] = dl_zero4();
  bool h0 = itr.next(g, it0), h1 = false, h2 = false;
  if (h0) { load_ck(it0, m0); load_sr(m0); load_row(it0, zi, gi); }
  h1 = h0 && itr.next(g, it1);
  if (h1) load_ck(it1, m1);

  while (h0) {
    if (h1) { load_sr(m1); load_row(it1, ziB, giB); }   // in flight during this item
    h2 = h1 && itr.next(g, it2);
    if (h2) load_ck(it2, m2);

    float4 dz[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) dz[p] = dl_zero4();
    for (long long base = it0.e0; base < it0.e1; base += 32) {
      const int cnt = (int)min(32LL, it0.e1 - base);
      int c = m0.c, kk = m0.kk;
      float sj = m0.sj, rj = m0.rj;
      if (base != it0.e0) {
        c = 0; kk = 255; sj = 1.0f; rj = 0.0f;
        if (lane < cnt) {
          c = __ldg(g.col + base + lane);
          kk = __ldg(kstar + base + lane);
          sj = __ldg(s + (long long)c * K + kk);
          rj = __ldg(r + (long long)c * K + kk);
        }
      }
      const int nsub = (cnt + EB - 1) / EB;
      for (int sb = 0; sb < nsub; ++sb) {
        float4 zj[EB][NP], gje[EB];
        int gpass[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          const int idx = sb * EB + e;
          const int cc = __shfl_sync(DL_FULL, c, idx & 31);
          const int ke = __shfl_sync(DL_FULL, kk, idx & 31);   // 255 beyond cnt
          const bool valid = idx < cnt;
#pragma unroll
          for (int p = 0; p < NP; ++p)
            zj[e][p] = (valid && M::active(lane, p)) ? dl_ldg4(Z + (long long)cc * D + M::offset(lane, p))
                                                     : dl_zero4();
          const bool m = valid && glane && ke < K && (ke % FPP) == slot;
          gpass[e] = m ? ke / FPP : -1;
          gje[e] = m ? dl_ldg4(G + (long long)cc * D + ke * d + 4 * gg) : dl_zero4();
        }
        // softmax over the factors recomputed in canonical arithmetic (bit-identical to forward)
        float ev[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          float part[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) part[e] = dl_chunk_dot(zi[p], zj[e][p]);
          float q = dl_reduce_scatter<M>(part, lane);
          if (!unit_T) q = __fdiv_rn(q, T);
          ev[p] = dl_expf(q);
        }
        const int my_idx = sb * EB + my_e;
        const bool valid = my_idx < cnt;
        int ks = __shfl_sync(DL_FULL, kk, my_idx & 31);        // stored routing of my edge
        ks = valid ? ks : 0;
        float sum = 0.0f, eks = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float ek = __shfl_sync(DL_FULL, ev[k / FPP], (k % FPP) * LP + gsrc);
          sum = (k == 0) ? ek : __fadd_rn(sum, ek);
          if (k == ks) eks = ek;
        }
        const float wv = __fdiv_rn(eks, sum);                  // = w[e] of the forward, same bits
        // c_ij = (1-beta) <G[i,ks], Z[j,ks]>, c_ji = (1-beta) <G[j,ks], Z[i,ks]>: partials exist only
        // on the lanes that own factor kstar(e) of edge e
        float pij[EB], pji[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          float x = 0.0f, y = 0.0f;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            if (gpass[e] == p) {
              x = dl_chunk_dot(gi[p], zj[e][p]);
              y = dl_chunk_dot(gje[e], zi[p]);
            }
          }
          pij[e] = x;
          pji[e] = y;
        }
        const float rij = dl_reduce_scatter<M>(pij, lane);
        const float rji = dl_reduce_scatter<M>(pji, lane);
        const int ksrc = (ks % FPP) * LP + gsrc;
        const float cij = __fmul_rn(omb, __shfl_sync(DL_FULL, rij, ksrc));
        const float cji = __fmul_rn(omb, __shfl_sync(DL_FULL, rji, ksrc));
        const float sjv = __shfl_sync(DL_FULL, sj, my_idx & 31);
        const float rjv = __shfl_sync(DL_FULL, rj, my_idx & 31);
        const float siv = __ldg(s + it0.node * K + ks);
        const float riv = __ldg(r + it0.node * K + ks);
        float dws = __fadd_rn(__fdiv_rn(cij, sjv), __fdiv_rn(cji, siv));
        dws = __fsub_rn(dws, riv);
        dws = __fsub_rn(dws, rjv);
        float basec = __fmul_rn(dws, wv);
        if (!unit_T) basec = __fdiv_rn(basec, T);
        basec = valid ? basec : 0.0f;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const float a_own = __fdiv_rn(ev[p], sum);
          const float ind = (M::factor(lane, p) == ks) ? 1.0f : 0.0f;
          const float coef_own = __fmul_rn(basec, __fsub_rn(ind, a_own));
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            const float cf = __shfl_sync(DL_FULL, coef_own, gbase + M::lane_of_edge(e));
            dl_fma4(dz[p], cf, zj[e][p]);
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (!M::active(lane, p)) continue;
      const int o = M::offset(lane, p);
      if (it0.hub_slot >= 0) {
        *reinterpret_cast<float4*>(hub_ws + it0.hub_slot * D + o) = dz[p];
      } else {
        float4* dst = reinterpret_cast<float4*>(dZ + it0.node * D + o);
        float4 cur = *dst;
        cur.x = __fadd_rn(cur.x, dz[p].x); cur.y = __fadd_rn(cur.y, dz[p].y);
        cur.z = __fadd_rn(cur.z, dz[p].z); cur.w = __fadd_rn(cur.w, dz[p].w);
        *dst = cur;
      }
    }
    it0 = it1; m0 = m1; h0 = h1;
    it1 = it2; m1 = m2; h1 = h2;
#pragma unroll
    for (int p = 0; p < NP; ++p) { zi[p] = ziB[p]; gi[p] = giB[p]; }
  }
}

// hub rows of pass 2: dZ[row] += sum of segment partials (in order)
__global__ void k_bwd_edges_hub_fixup(DlGraphDev g, long long D, const float* __restrict__ hub_ws,
                                      float* __restrict__ dZ) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < g.n_hub * D; x += stride) {
    long long h = x / D, o = x % D;
    long long a = g.hub_seg_ptr[h], b = g.hub_seg_ptr[h + 1];
    float v = 0.0f;
    for (long long sg = a; sg < b; ++sg) v = __fadd_rn(v, hub_ws[sg * D + o]);
    long long row = g.row_base + g.perm[h];
    dZ[row * D + o] = __fadd_rn(dZ[row * D + o], v);
  }
}

// ---------------------------------------------------------------------------------------------
// runtime-generic path (any K, d): one warp per item, one entry at a time
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(DL_FULL, v, o));
  return v;
}

// pass 1: per factor the row's entries are rescanned so the accumulator fits in registers
__global__ void __launch_bounds__(DL_CTA)
k_factor_bwd_gather_generic(DlGraphDev g, const float* __restrict__ Z, const float* __restrict__ G,
                            const unsigned char* __restrict__ kstar, const float* __restrict__ w,
                            const float* __restrict__ s, int K, int d, float beta, float omb,
                            float* __restrict__ dZ, float* __restrict__ r, float* __restrict__ hub_ws) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long n_items = dl_num_items(g);
  const long long D = (long long)K * d;
  constexpr int R = DL_MAX_D / 32;
  for (long long t = warp0; t < n_items; t += nwarps) {
    const DlItem it = dl_decode_item(g, t);
    for (int k = 0; k < K; ++k) {
      float acc[R];
#pragma unroll
      for (int x = 0; x < R; ++x) acc[x] = 0.0f;
      for (long long p = it.e0; p < it.e1; ++p) {
        if (__ldg(kstar + p) != k) continue;
        const float wv = __ldg(w + p);
        const float* gj = G + (long long)__ldg(g.col + p) * D + (long long)k * d;
#pragma unroll
        for (int x = 0; x < R; ++x) {
          int e = lane + 32 * x;
          if (e < d) acc[x] = __fmaf_rn(wv, gj[e], acc[x]);
        }
      }
      if (it.hub_slot >= 0) {
#pragma unroll
        for (int x = 0; x < R; ++x) {
          int e = lane + 32 * x;
          if (e < d) hub_ws[it.hub_slot * D + (long long)k * d + e] = acc[x];
        }
      } else {
        const long long row = it.node;
        const float sk = __ldg(s + row * K + k);
        const float scale = __fdiv_rn(omb, sk);
        float part = 0.0f;
#pragma unroll
        for (int x = 0; x < R; ++x) {
          int e = lane + 32 * x;
          if (e < d) {
            const float tv = __fmul_rn(scale, acc[x]);
            part = __fmaf_rn(Z[row * D + (long long)k * d + e], tv, part);
            float* dst = dZ + row * D + (long long)k * d + e;
            *dst = __fadd_rn(*dst, __fmaf_rn(beta, G[row * D + (long long)k * d + e], tv));
          }
        }
        part = warp_sum(part);
        if (lane == 0) r[row * K + k] = __fdiv_rn(part, sk);
      }
    }
  }
}

__global__ void __launch_bounds__(DL_CTA)
k_factor_bwd_gather_hub_generic(DlGraphDev g, const float* __restrict__ Z, const float* __restrict__ G,
                                const float* __restrict__ s, int K, int d, float beta, float omb,
                                float* __restrict__ dZ, float* __restrict__ r,
                                const float* __restrict__ hub_ws) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long D = (long long)K * d;
  for (long long h = warp0; h < g.n_hub; h += nwarps) {
    const long long a = g.hub_seg_ptr[h], b = g.hub_seg_ptr[h + 1];
    const long long row = g.row_base + g.perm[h];
    for (int k = 0; k < K; ++k) {
      const float sk = __ldg(s + row * K + k);
      const float scale = __fdiv_rn(omb, sk);
      float part = 0.0f;
      for (int e = lane; e < d; e += 32) {
        float acc = 0.0f;
        for (long long sg = a; sg < b; ++sg) acc = __fadd_rn(acc, hub_ws[sg * D + (long long)k * d + e]);
        const float tv = __fmul_rn(scale, acc);
        part = __fmaf_rn(Z[row * D + (long long)k * d + e], tv, part);
        float* dst = dZ + row * D + (long long)k * d + e;
        *dst = __fadd_rn(*dst, __fmaf_rn(beta, G[row * D + (long long)k * d + e], tv));
      }
      part = warp_sum(part);
      if (lane == 0) r[row * K + k] = __fdiv_rn(part, sk);
    }
  }
}

// pass 2: direct rows accumulate into their own dZ row, hub segments into their scratch slot
__global__ void __launch_bounds__(DL_CTA)
k_factor_bwd_edges_generic(DlGraphDev g, const float* __restrict__ Z, const float* __restrict__ G,
                           const float* __restrict__ s, const float* __restrict__ r, int K, int d,
                           float omb, float T, float* __restrict__ dZ, float* __restrict__ hub_ws) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long n_items = dl_num_items(g);
  const long long D = (long long)K * d;
  float e[DL_MAX_K], a[DL_MAX_K];
  for (long long t = warp0; t < n_items; t += nwarps) {
    const DlItem it = dl_decode_item(g, t);
    const long long i = it.node;
    float* acc = it.hub_slot >= 0 ? hub_ws + it.hub_slot * D : dZ + i * D;
    if (it.hub_slot >= 0) {
      for (long long x = lane; x < D; x += 32) acc[x] = 0.0f;
    }
    __syncwarp();
    const float* zi = Z + i * D;
    const float* gi = G + i * D;
    for (long long p = it.e0; p < it.e1; ++p) {
      const long long j = __ldg(g.col + p);
      const float* zj = Z + j * D;
      const float* gj = G + j * D;
      const int k = dl_generic_route(zi, zj, K, d, T, lane, e, a);
      const float wk = a[k];
      float cij = 0.0f, cji = 0.0f;
      for (int x = lane; x < d; x += 32) {
        cij = __fmaf_rn(gi[k * d + x], zj[k * d + x], cij);
        cji = __fmaf_rn(gj[k * d + x], zi[k * d + x], cji);
      }
      cij = __fmul_rn(omb, warp_sum(cij));
      cji = __fmul_rn(omb, warp_sum(cji));
      float dws = __fadd_rn(__fdiv_rn(cij, __ldg(s + j * K + k)), __fdiv_rn(cji, __ldg(s + i * K + k)));
      dws = __fsub_rn(dws, __ldg(r + i * K + k));
      dws = __fsub_rn(dws, __ldg(r + j * K + k));
      const float basec = __fdiv_rn(__fmul_rn(dws, wk), T);
      for (int kk = 0; kk < K; ++kk) {
        const float coef = __fmul_rn(basec, __fsub_rn((kk == k) ? 1.0f : 0.0f, a[kk]));
        for (int x = lane; x < d; x += 32)
          acc[kk * d + x] = __fmaf_rn(coef, zj[kk * d + x], acc[kk * d + x]);
      }
    }
    __syncwarp();
  }
}

inline int fixup_blocks(long long n) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

template <class M>
int launch_bwd_edges(const DlGraphDev& g, long long n_items, const float* Z, const float* G,
                     const uint8_t* kstar, const float* s, const float* r, float omb, float T,
                     float* dZ, float* hub_ws, cudaStream_t st) {
  int grid = 1;
  int rc = dl_grid_for(k_factor_bwd_edges<M>, n_items, &grid);
  if (rc) return rc;
  k_factor_bwd_edges<M><<<grid, DL_CTA, 0, st>>>(g, Z, G, kstar, s, r, omb, T, dZ, hub_ws);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

}  // namespace

extern "C" {

static int bwd_gather_impl(const dl_graph* g_host, const float* Z, const float* G,
                           const uint8_t* kstar, const float* w, const float* s, int K, int d,
                           float beta, float one_minus_beta, float* dZ, float* r, float* x, const int32_t* x_index,
                           uint8_t* ku_out, int* x_valid_out, float* hub_ws, float* const* r_peers, int n_peers,
                           dl_stream_t stream) {
  if (x_valid_out) *x_valid_out = 0;
  if (!dl_graph_ok(g_host) || !dl_shape_ok(K, d)) return DL_EINVAL;
  if (g_host->N == 0) return DL_OK;
  if (!Z || !G || !s || !dZ || !r || (g_host->nnz > 0 && (!kstar || !w))) return DL_EINVAL;
  if (g_host->n_hub_items > 0 && !hub_ws) return DL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned flags = g_host->flags;
  DlGraphDev g = dl_graph_dev(g_host);
  if (!dl_set_peer_out(g, r_peers, n_peers)) return DL_EINVAL;
  const long long n_items = g.n_hub_items + (g.N - g.n_hub);
  int rc = -1000;
  if (!(flags & DL_F_NO_STREAM)) {
    // the streaming pass 1 leaves x[e] = <G[j,k*], Z[i,k*]> for pass 2 when asked to
    float* xo = (x && !(flags & DL_F_NO_XDOT) && dl_gather_stream_has_x(K, d)) ? x : nullptr;
    rc = dl_launch_gather_stream(1, g, Z, G, kstar, w, s, K, d, beta, one_minus_beta, dZ, r, hub_ws, st, xo,
                                 xo ? x_index : nullptr, (xo && x_index) ? ku_out : nullptr);
    if (rc == DL_OK && xo && x_valid_out) *x_valid_out = 1;
  }
  if (rc == DL_OK) return DL_OK;
  if (rc != -1000) return rc;
  rc = dl_launch_slice_gather(1, g, n_items, Z, G, kstar, w, s, K, d, beta, one_minus_beta, dZ, r,
                              hub_ws, st);
  if (rc == -1000) {
    int grid = 1;
    rc = dl_grid_for(k_factor_bwd_gather_generic, n_items, &grid);
    if (rc) return rc;
    k_factor_bwd_gather_generic<<<grid, DL_CTA, 0, st>>>(g, Z, G, kstar, w, s, K, d, beta,
                                                         one_minus_beta, dZ, r, hub_ws);
    DL_LAUNCH_CHECK();
    if (g.n_hub > 0) {
      rc = dl_grid_for(k_factor_bwd_gather_hub_generic, g.n_hub, &grid);
      if (rc) return rc;
      k_factor_bwd_gather_hub_generic<<<grid, DL_CTA, 0, st>>>(g, Z, G, s, K, d, beta, one_minus_beta,
                                                             dZ, r, hub_ws);
      DL_LAUNCH_CHECK();
    }
    rc = DL_OK;
  }
  if (rc) return rc;
  if (n_peers > 0) {                           // the row-per-warp paths do not push: one copy kernel does
    void* dst[DL_MAX_PEER_OUT];
    const long long off = g.row_base * (long long)K;
    for (int q = 0; q < n_peers; ++q) dst[q] = r_peers[q] + off;
    return dl_push_slice(r + off, dst, n_peers, (int64_t)g.N * K * 4, stream);
  }
  return DL_OK;
}

int dl_factor_bwd_gather(const dl_graph* g_host, const float* Z, const float* G,
                         const uint8_t* kstar, const float* w, const float* s, int K, int d,
                         float beta, float one_minus_beta, float* dZ, float* r, float* x, const int32_t* x_index,
                         uint8_t* ku_out, int* x_valid_out, float* hub_ws, dl_stream_t stream) {
  return bwd_gather_impl(g_host, Z, G, kstar, w, s, K, d, beta, one_minus_beta, dZ, r, x, x_index, ku_out,
                         x_valid_out, hub_ws, nullptr, 0, stream);
}

int dl_factor_bwd_gather_push(const dl_graph* g_host, const float* Z, const float* G,
                              const uint8_t* kstar, const float* w, const float* s, int K, int d,
                              float beta, float one_minus_beta, float* dZ, float* r, float* x, int* x_valid_out,
                              float* hub_ws, float* const* r_peers, int n_peers, dl_stream_t stream) {
  return bwd_gather_impl(g_host, Z, G, kstar, w, s, K, d, beta, one_minus_beta, dZ, r, x, nullptr, nullptr,
                         x_valid_out, hub_ws, r_peers, n_peers, stream);
}

int dl_factor_bwd_edges(const dl_graph* g_host, const float* Z, const float* G,
                        const uint8_t* kstar, const float* w, const float* s, const float* r,
                        const float* sj, float* sr_scratch, int64_t n_nodes, const float* x, int K, int d,
                        float one_minus_beta, float T, float* dZ, float* hub_ws, dl_stream_t stream) {
  if (!dl_graph_ok(g_host) || !dl_shape_ok(K, d)) return DL_EINVAL;
  if (g_host->N == 0) return DL_OK;
  if (!Z || !G || !s || !dZ || !r || (g_host->nnz > 0 && !kstar)) return DL_EINVAL;
  if (g_host->n_hub_items > 0 && !hub_ws) return DL_EINVAL;
  if (!(T == T) || T == 0.0f) return DL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned flags = g_host->flags;
  const DlGraphDev g = dl_graph_dev(g_host);
  const long long n_items = g.n_hub_items + (g.N - g.n_hub);
  const long long D = (long long)K * d;
  int rc = -1000;
  // factor-per-lane kernel (bwd_fl.cu) for the K <= 8, d <= 16 shape class.  It reads the per-entry
  // dot x pass 1 left behind (no second gather of the routed G slice) and gathers (s, r) of the
  // neighbour packed in one 8-byte access; flags switch either off for A/B runs.
  if (!(flags & (DL_F_NO_STREAM | DL_F_NO_FL)))
    rc = dl_launch_bwd_edges_fl(g, Z, G, kstar, s, r, (flags & DL_F_NO_SJ) ? nullptr : sj,
                                (flags & DL_F_NO_SR) ? nullptr : sr_scratch, n_nodes,
                                (flags & DL_F_NO_XDOT) ? nullptr : x, K, d, one_minus_beta, T, dZ, hub_ws, st);
  if (rc == DL_OK) return DL_OK;
  if (rc != -1000) return rc;
  if (!(flags & DL_F_NO_STREAM))
    rc = dl_launch_bwd_edges_stream(g, Z, G, kstar, s, r, K, d, one_minus_beta, T, dZ, hub_ws, st);
  if (rc == DL_OK) return DL_OK;
  if (rc != -1000) return rc;
#define BODY_MACRO(M) \
  rc = launch_bwd_edges<M>(g, n_items, Z, G, kstar, s, r, one_minus_beta, T, dZ, hub_ws, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  if (rc == -1000) {
    int grid = 1;
    rc = dl_grid_for(k_factor_bwd_edges_generic, n_items, &grid);
    if (rc) return rc;
    k_factor_bwd_edges_generic<<<grid, DL_CTA, 0, st>>>(g, Z, G, s, r, K, d, one_minus_beta, T, dZ, hub_ws);
    DL_LAUNCH_CHECK();
    rc = DL_OK;
  }
  if (rc) return rc;
  if (g.n_hub > 0) {
    k_bwd_edges_hub_fixup<<<fixup_blocks(g.n_hub * D), 256, 0, st>>>(g, D, hub_ws, dZ);
    DL_LAUNCH_CHECK();
  }
  return DL_OK;
}

int dl_factor_bwd(const dl_graph* g_host, const float* Z, const float* G, const uint8_t* kstar,
                  const float* w, const float* s, const float* sj, float* sr_scratch, int64_t n_nodes,
                  float* x_scratch, int K, int d, float beta, float one_minus_beta, float T, float* dZ,
                  float* r, float* hub_ws, dl_stream_t stream) {
  if (!(T == T) || T == 0.0f) return DL_EINVAL;
  int x_valid = 0;
  int rc = dl_factor_bwd_gather(g_host, Z, G, kstar, w, s, K, d, beta, one_minus_beta, dZ, r, x_scratch, nullptr,
                                nullptr, &x_valid, hub_ws, stream);
  if (rc) return rc;
  return dl_factor_bwd_edges(g_host, Z, G, kstar, w, s, r, sj, sr_scratch, n_nodes,
                             x_valid ? x_scratch : nullptr, K, d, one_minus_beta, T, dZ, hub_ws, stream);
}

}  // extern "C"
