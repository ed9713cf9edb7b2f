// slice_gather.cu -- per-factor gather / segment-sum over a destination-sorted CSR.
//
//   MODE 0  aggregation forward  [ref: model.py:75]
//           H[i,k] = beta Z[i,k] + (1-beta) sum_{e in row i, kstar=k} (w[e] / s[col_e,k]) Z[col_e,k]
//   MODE 1  backward pass 1      [ref: autograd of model.py:70-75]
//           T_[i,k] = (1-beta)/s[i,k] sum_{e in row i, kstar=k} w[e] G[col_e,k]
//           r[i,k] = <Z[i,k],T_[i,k]>/s[i,k] ;  dZ[i] += beta G[i] + T_[i]
//
// Only the routed factor's slice of a neighbour row is needed (d floats = LP lanes of float4), so
// the warp is split into NG = 32/LP lane groups and ONE load instruction fetches the slices of NG
// different edges (8 x 64 B at d = 16).  Each lane keeps K float4 accumulators, predicated on the
// factor of the edge its group is handling; at the end of the row the NG group partials are
// combined in group order through shared memory, which also transposes them into the row layout
// (lane (k, g) owns chunk g of factor k) for a coalesced store.  No atomics; the summation order
// of a row is fixed by the CSR order alone (edges e = g mod NG in sequence, then groups 0..NG-1).
//
// Latency: items come from DlRowIter (natural row order, row bounds prefetched); the first
// (col, kstar, w) block of the item two ahead and the s[col,kstar] gather of the next item are in
// flight while the current item is processed.
#include "dl_dispatch.cuh"

namespace {

struct GMeta {
  int c, k;
  float wv, sj;
};

// epilogue of backward pass 1 for one row held in DlMap lane layout
template <class M>
__device__ __forceinline__ void bwd_gather_epilogue(int lane, long long row, const float4 (&acc)[M::NP],
                                                    const float* __restrict__ Z, const float* __restrict__ G,
                                                    const float* __restrict__ s, float beta, float omb,
                                                    float* __restrict__ dZ, float* __restrict__ r) {
  constexpr int K = M::K, D = M::D, NP = M::NP, LP = M::LP;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int k = M::factor(lane, p);
    const bool act = M::active(lane, p);
    const int o = M::offset(lane, p);
    const float sk = (k < K) ? __ldg(s + row * K + k) : 1.0f;
    const float scale = __fdiv_rn(omb, sk);
    float4 tv;
    tv.x = __fmul_rn(scale, acc[p].x); tv.y = __fmul_rn(scale, acc[p].y);
    tv.z = __fmul_rn(scale, acc[p].z); tv.w = __fmul_rn(scale, acc[p].w);
    const float4 zi = act ? dl_ldg4(Z + row * D + o) : dl_zero4();
    const float dotzt = dl_group_sum<M>(dl_chunk_dot(zi, tv));
    if (k < K && (lane % LP) == 0) r[row * K + k] = __fdiv_rn(dotzt, sk);
    if (act) {
      const float4 gi = dl_ldg4(G + row * D + o);
      float4* dst = reinterpret_cast<float4*>(dZ + row * D + o);
      float4 cur = *dst;
      cur.x = __fadd_rn(cur.x, __fmaf_rn(beta, gi.x, tv.x));
      cur.y = __fadd_rn(cur.y, __fmaf_rn(beta, gi.y, tv.y));
      cur.z = __fadd_rn(cur.z, __fmaf_rn(beta, gi.z, tv.z));
      cur.w = __fadd_rn(cur.w, __fmaf_rn(beta, gi.w, tv.w));
      *dst = cur;
    }
  }
}

template <class M>
struct GatherCfg {
  static constexpr int NG = 32 / M::LP;          // lane groups = edges per load instruction
  static constexpr int GS = M::K * M::LP + 4;    // group stride in float4 units; +4 breaks bank aliasing
  static constexpr size_t SMEM = (size_t)DL_WARPS_PER_CTA * NG * GS * sizeof(float4);
};

template <class M, int MODE>
__global__ void __launch_bounds__(DL_CTA)
k_factor_gather(DlGraphDev g, const float* __restrict__ Z, const float* __restrict__ SRC,
                const unsigned char* __restrict__ kstar, const float* __restrict__ w,
                const float* __restrict__ s, float beta, float omb, float* __restrict__ OUT,
                float* __restrict__ r, float* __restrict__ hub_ws) {
  constexpr int K = M::K, d = M::d, D = M::D, NP = M::NP, L = M::L, LP = M::LP;
  constexpr int NG = GatherCfg<M>::NG, GS = GatherCfg<M>::GS;
  constexpr int UNR = LP < 4 ? LP : 4;
  extern __shared__ float4 dl_smem_f4[];
  const int lane = threadIdx.x & 31;
  float4* sm = dl_smem_f4 + (threadIdx.x >> 5) * (NG * GS);
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const int grp = lane / LP, gg = lane % LP;
  const bool glane = gg < L;

  auto load_ckw = [&](const DlItem& it, GMeta& m) {
    m.c = 0; m.k = 255; m.wv = 0.0f; m.sj = 1.0f;
    if (lane < it.e1 - it.e0) {
      m.c = __ldg(g.col + it.e0 + lane);
      m.k = __ldg(kstar + it.e0 + lane);
      m.wv = __ldg(w + it.e0 + lane);
    }
  };
  auto load_sj = [&](GMeta& m) {
    if (MODE == 0 && m.k != 255) m.sj = __ldg(s + (long long)m.c * K + m.k);
  };

  DlRowIter itr;
  itr.init(g, warp0, nwarps);
  DlItem it0, it1, it2;
  GMeta m0, m1, m2;
  it1 = it2 = it0 = DlItem{0, 0, 0, 0, -1};
  m0 = m1 = m2 = GMeta{0, 255, 0.0f, 1.0f};
  bool h0 = itr.next(g, it0), h1 = false, h2 = false;
  if (h0) { load_ckw(it0, m0); load_sj(m0); }
  h1 = h0 && itr.next(g, it1);
  if (h1) load_ckw(it1, m1);

  while (h0) {
    if (h1) load_sj(m1);                       // col/kstar of the next item arrived an iteration ago
    h2 = h1 && itr.next(g, it2);
    if (h2) load_ckw(it2, m2);                 // in flight for two iterations

    float4 acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = dl_zero4();
    for (long long base = it0.e0; base < it0.e1; base += 32) {
      const int cnt = (int)min(32LL, it0.e1 - base);
      int c = m0.c, k = m0.k;
      float wv = m0.wv, sj = m0.sj;
      if (base != it0.e0) {
        c = 0; k = 255; wv = 0.0f; sj = 1.0f;
        if (lane < cnt) {
          c = __ldg(g.col + base + lane);
          k = __ldg(kstar + base + lane);
          wv = __ldg(w + base + lane);
          if (MODE == 0) sj = __ldg(s + (long long)c * K + k);
        }
      }
      const float coef = (MODE == 0) ? __fdiv_rn(wv, sj) : wv;
      // rounds in batches of UNR: all gathers of a batch are issued before any of them is used, so
      // a warp keeps UNR * NG neighbour slices (32 at d = 16) in flight
      for (int rb = 0; rb * NG < cnt; rb += UNR) {
        float4 z[UNR];
        float cf[UNR];
        int kk[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int idx = ((rb + u) * NG + grp) & 31;
          const int cc = __shfl_sync(DL_FULL, c, idx);
          kk[u] = __shfl_sync(DL_FULL, k, idx);
          cf[u] = __shfl_sync(DL_FULL, coef, idx);
          const bool valid = glane && kk[u] < K && (rb + u) < LP;   // lanes beyond cnt carry k = 255
          kk[u] = valid ? kk[u] : 255;
          z[u] = valid ? dl_ldg4(SRC + (long long)cc * D + kk[u] * d + 4 * gg) : dl_zero4();
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if ((rb + u) * NG < cnt) {             // warp-uniform
#pragma unroll
            for (int kq = 0; kq < K; ++kq)
              if (kk[u] == kq) dl_fma4(acc[kq], cf[u], z[u]);
          }
        }
      }
    }
    // combine the NG group partials in group order; lane (slot, g) ends up with chunk g of factor
    // p*FPP + slot, the row layout
    if (glane) {
#pragma unroll
      for (int kq = 0; kq < K; ++kq) sm[grp * GS + kq * LP + gg] = acc[kq];
    }
    __syncwarp();
    float4 tot[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      tot[p] = dl_zero4();
      if (M::active(lane, p)) {
        const int f = M::factor(lane, p);
#pragma unroll
        for (int sg = 0; sg < NG; ++sg) {
          const float4 v = sm[sg * GS + f * LP + gg];
          tot[p].x = __fadd_rn(tot[p].x, v.x); tot[p].y = __fadd_rn(tot[p].y, v.y);
          tot[p].z = __fadd_rn(tot[p].z, v.z); tot[p].w = __fadd_rn(tot[p].w, v.w);
        }
      }
    }
    __syncwarp();
    if (it0.hub_slot >= 0) {
#pragma unroll
      for (int p = 0; p < NP; ++p)
        if (M::active(lane, p))
          *reinterpret_cast<float4*>(hub_ws + it0.hub_slot * D + M::offset(lane, p)) = tot[p];
    } else if (MODE == 0) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (!M::active(lane, p)) continue;
        const int o = M::offset(lane, p);
        const float4 zi = dl_ldg4(Z + it0.node * D + o);
        float4 h;
        h.x = __fadd_rn(__fmul_rn(beta, zi.x), __fmul_rn(omb, tot[p].x));
        h.y = __fadd_rn(__fmul_rn(beta, zi.y), __fmul_rn(omb, tot[p].y));
        h.z = __fadd_rn(__fmul_rn(beta, zi.z), __fmul_rn(omb, tot[p].z));
        h.w = __fadd_rn(__fmul_rn(beta, zi.w), __fmul_rn(omb, tot[p].w));
        *reinterpret_cast<float4*>(OUT + it0.node * D + o) = h;
      }
    } else {
      bwd_gather_epilogue<M>(lane, it0.node, tot, Z, SRC, s, beta, omb, OUT, r);
    }
    it0 = it1; m0 = m1; h0 = h1;
    it1 = it2; m1 = m2; h1 = h2;
  }
}

// hub rows of backward pass 1: one warp per hub row sums the segment partials (in order), then the
// epilogue
template <class M>
__global__ void __launch_bounds__(DL_CTA)
k_factor_bwd_gather_hub(DlGraphDev g, const float* __restrict__ Z, const float* __restrict__ G,
                        const float* __restrict__ s, float beta, float omb, float* __restrict__ dZ,
                        float* __restrict__ r, const float* __restrict__ hub_ws) {
  constexpr int D = M::D, NP = M::NP;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  for (long long h = warp0; h < g.n_hub; h += nwarps) {
    const long long a = g.hub_seg_ptr[h], b = g.hub_seg_ptr[h + 1];
    float4 acc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] = dl_zero4();
    for (long long sg = a; sg < b; ++sg) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (!M::active(lane, p)) continue;
        const float4 v = *reinterpret_cast<const float4*>(hub_ws + sg * D + M::offset(lane, p));
        acc[p].x = __fadd_rn(acc[p].x, v.x); acc[p].y = __fadd_rn(acc[p].y, v.y);
        acc[p].z = __fadd_rn(acc[p].z, v.z); acc[p].w = __fadd_rn(acc[p].w, v.w);
      }
    }
    bwd_gather_epilogue<M>(lane, g.row_base + g.perm[h], acc, Z, G, s, beta, omb, dZ, r);
  }
}

template <class M, int MODE>
int launch_gather(const DlGraphDev& g, long long n_items, const float* Z, const float* SRC,
                  const unsigned char* kstar, const float* w, const float* s, float beta, float omb,
                  float* OUT, float* r, float* hub_ws, cudaStream_t st) {
  constexpr size_t smem = GatherCfg<M>::SMEM;
  if (smem > 48 * 1024)
    DL_CUDA_TRY(cudaFuncSetAttribute(k_factor_gather<M, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  int grid = 1;
  int rc = dl_grid_for(k_factor_gather<M, MODE>, n_items, &grid, smem);
  if (rc) return rc;
  k_factor_gather<M, MODE><<<grid, DL_CTA, smem, st>>>(g, Z, SRC, kstar, w, s, beta, omb, OUT, r, hub_ws);
  DL_LAUNCH_CHECK();
  if (MODE == 1 && g.n_hub > 0) {
    rc = dl_grid_for(k_factor_bwd_gather_hub<M>, g.n_hub, &grid);
    if (rc) return rc;
    k_factor_bwd_gather_hub<M><<<grid, DL_CTA, 0, st>>>(g, Z, SRC, s, beta, omb, OUT, r, hub_ws);
    DL_LAUNCH_CHECK();
  }
  return DL_OK;
}

}  // namespace

int dl_launch_slice_gather(int mode, const DlGraphDev& g, long long n_items, const float* Z,
                           const float* SRC, const unsigned char* kstar, const float* w,
                           const float* s, int K, int d, float beta, float omb, float* OUT, float* r,
                           float* hub_ws, cudaStream_t st) {
  int rc = -1000;
#define BODY_MACRO(M)                                                                                      \
  rc = (mode == 0) ? launch_gather<M, 0>(g, n_items, Z, SRC, kstar, w, s, beta, omb, OUT, r, hub_ws, st)  \
                   : launch_gather<M, 1>(g, n_items, Z, SRC, kstar, w, s, beta, omb, OUT, r, hub_ws, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  return rc;
}
