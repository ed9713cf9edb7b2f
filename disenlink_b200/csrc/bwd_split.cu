// bwd_split.cu -- backward pass 2 of the factor attention as two lean streaming kernels.
//
// [ref: autograd of model.py:56-75]  per CSR entry e = (i,j), k = kstar(e):
//   (A) k_bwd_base_stream   base[e] = (c_ij/s[j,k] + c_ji/s[i,k] - r[i,k] - r[j,k]) * w[e] / T
//                            c_ij = (1-beta) <G[i,k], Z[j,k]>,  c_ji = (1-beta) <G[j,k], Z[i,k]>
//       needs only the routed d-float slices Z[j,k], G[j,k] (2 x 64 B) and scalars -> a
//       slice-gather kernel; the scalar arithmetic is done for 32 entries at once, one per lane.
//   (B) k_bwd_accum_stream  dZ[i,kk,:] += base[e] * ((kk==k) - a[e,kk]) * Z[j,kk,:]  for all kk
//       with a[] = softmax over factors recomputed from Z (canonical arithmetic): the shape of the
//       attention kernel (one 512-B row gather per entry) plus a per-row register accumulator.
// The fused single-kernel version (bwd_stream.cu) is issue-bound at its register-limited 16 warps
// per SM; splitting removes the G traffic, the two extra reduce-scatters and ~40 registers from
// the heavy kernel.  Cost: 8 extra bytes per entry (base written and read once).
#include "dl_dispatch.cuh"
#include "dl_stream.cuh"

namespace {

constexpr int BA_WARPS = 8;    // 64 KB of slice tiles per CTA -> 3 CTAs = 24 warps per SM
constexpr int BB_RING = 2;
constexpr int BB_OWN = 2;

// ============================================================================ (A) base[e]
template <class M>
struct BaseCfg {
  static constexpr int SLB = M::d * 4;
  static constexpr int TILE_B = DL_CH * SLB;                         // one chunk of slices
  static constexpr size_t SMEM = (size_t)BA_WARPS * 2 * 2 * TILE_B;  // Z and G tiles, double buffered
};

struct AMeta {
  int row, col, ks;
  float wv, sj, rj, si, ri;
};

template <class M>
__global__ void __launch_bounds__(BA_WARPS * 32)
k_bwd_base_stream(DlGraphDev g, const float* __restrict__ Z, const float* __restrict__ G,
                  const unsigned char* __restrict__ kstar, const float* __restrict__ w,
                  const float* __restrict__ s, const float* __restrict__ r, float omb, float T,
                  float* __restrict__ base) {
  using C = BaseCfg<M>;
  constexpr int K = M::K, d = M::d, D = M::D, NP = M::NP, L = M::L, LP = M::LP, FPP = M::FPP;
  constexpr int NG = 32 / LP, SLB = C::SLB, TILE_B = C::TILE_B;
  extern __shared__ __align__(128) unsigned char dl_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* tiles = dl_smem_raw + (size_t)warp * 4 * TILE_B;   // [buf][Z|G][TILE_B]
  const long long gw = (long long)blockIdx.x * BA_WARPS + warp;
  const int grp = lane / LP, gg = lane % LP, slot = M::slot(lane);
  const bool glane = gg < L;

  DlChunkStream cs;
  cs.init(g.nnz, (long long)gridDim.x * BA_WARPS);

  auto load_meta = [&](long long cc, AMeta& m) {
    m.row = -1; m.col = 0; m.ks = 255; m.wv = 0.0f; m.sj = 1.0f; m.rj = 0.0f; m.si = 1.0f; m.ri = 0.0f;
    if (cc >= 0) {
      const long long e = cc * DL_CH + lane;
      if (e < g.nnz) {
        m.row = __ldg(g.erow + e); m.col = __ldg(g.col + e); m.ks = __ldg(kstar + e); m.wv = __ldg(w + e);
      }
    }
  };
  auto load_scalars = [&](AMeta& m) {
    if (m.row >= 0) {
      const long long node = g.row_base + m.row;
      m.sj = __ldg(s + (long long)m.col * K + m.ks);
      m.rj = __ldg(r + (long long)m.col * K + m.ks);
      m.si = __ldg(s + node * K + m.ks);
      m.ri = __ldg(r + node * K + m.ks);
    }
  };
  auto issue_slices = [&](unsigned char* buf, const AMeta& m) {
#pragma unroll
    for (int rd = 0; rd < LP; ++rd) {
      const int idx = rd * NG + grp;
      const int cc = __shfl_sync(DL_FULL, m.col, idx);
      const int kk = __shfl_sync(DL_FULL, m.ks, idx);
      if (glane && kk < K) {
        const long long o = (long long)cc * D + kk * d + gg * 4;
        dl_cp_async16(buf + idx * SLB + gg * 16, Z + o);
        dl_cp_async16(buf + TILE_B + idx * SLB + gg * 16, G + o);
      }
    }
  };

  long long c = cs.first(gw);
  AMeta mA, mB, mC;
  load_meta(c, mA);
  load_scalars(mA);
  long long cn = cs.next(c);
  load_meta(cn, mB);
  int buf = 0;
  issue_slices(tiles, mA);
  dl_cp_async_commit();
  float4 zi[NP], gi[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) zi[p] = gi[p] = dl_zero4();
  int cur_row = -1;

  while (c >= 0) {
    const long long cnn = cs.next(cn);
    load_meta(cnn, mC);
    load_scalars(mB);
    issue_slices(tiles + (buf ^ 1) * 2 * TILE_B, mB);
    dl_cp_async_commit();
    dl_cp_async_wait<1>();
    __syncwarp();

    const unsigned vmask = __ballot_sync(DL_FULL, mA.row >= 0);
    const int cnt = __popc(vmask);
    const unsigned char* zt = tiles + buf * 2 * TILE_B;
    const unsigned char* gt = zt + TILE_B;
    const int prow = __shfl_up_sync(DL_FULL, mA.row, 1);
    const bool st = mA.row >= 0 && (lane == 0 ? mA.row != cur_row : mA.row != prow);
    const unsigned starts = __ballot_sync(DL_FULL, st);
    float cx = 0.0f, cy = 0.0f;      // <G[i,k],Z[j,k]> and <G[j,k],Z[i,k]> of entry `lane`
    int idx = 0;
    while (idx < cnt) {
      if ((starts >> idx) & 1u) {    // new row run: its own rows (Z[i], G[i]) into registers
        cur_row = __shfl_sync(DL_FULL, mA.row, idx);
        const long long node = g.row_base + cur_row;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const bool act = M::active(lane, p);
          zi[p] = act ? dl_ldg4(Z + node * D + M::offset(lane, p)) : dl_zero4();
          gi[p] = act ? dl_ldg4(G + node * D + M::offset(lane, p)) : dl_zero4();
        }
      }
      const unsigned rest = (idx < 31) ? (starts & ~((2u << idx) - 1u)) : 0u;
      const int end = rest ? (__ffs(rest) - 1) : cnt;
      for (; idx < end; ++idx) {
        const unsigned ke = (unsigned)__shfl_sync(DL_FULL, mA.ks, idx);
        float x = 0.0f, y = 0.0f;
        if (glane && (ke & (FPP - 1)) == (unsigned)slot) {
          const float4 vz = dl_lds4(zt + idx * SLB + gg * 16);
          const float4 vg = dl_lds4(gt + idx * SLB + gg * 16);
          const unsigned pe = ke / FPP;
#pragma unroll
          for (int p = 0; p < NP; ++p)
            if (pe == (unsigned)p) { x = dl_chunk_dot(gi[p], vz); y = dl_chunk_dot(vg, zi[p]); }
        }
        x = dl_group_sum<M>(x);
        y = dl_group_sum<M>(y);
        const int src = (int)(ke & (FPP - 1)) * LP;
        const float xv = __shfl_sync(DL_FULL, x, src);
        const float yv = __shfl_sync(DL_FULL, y, src);
        if (lane == idx) { cx = xv; cy = yv; }
      }
    }
    // 32 entries at once, one per lane
    if (lane < cnt) {
      float dws = __fadd_rn(__fdiv_rn(__fmul_rn(omb, cx), mA.sj), __fdiv_rn(__fmul_rn(omb, cy), mA.si));
      dws = __fsub_rn(dws, mA.ri);
      dws = __fsub_rn(dws, mA.rj);
      base[c * DL_CH + lane] = __fdiv_rn(__fmul_rn(dws, mA.wv), T);
    }
    __syncwarp();
    buf ^= 1;
    c = cn; cn = cnn;
    mA = mB; mB = mC;
  }
  dl_cp_async_wait<0>();
}

// ============================================================================ (B) dZ accumulation
template <class M>
struct AccumCfg {
  static constexpr int ROWB = M::D * 4;
  static constexpr int STAGE_B = (DL_HS + BB_OWN) * ROWB;
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int NW_RAW = BUDGET / (BB_RING * STAGE_B);
  static constexpr bool OK = NW_RAW >= 4 && M::EB == 4;
  static constexpr int NW = NW_RAW >= 24 ? 24 : (NW_RAW >= 4 ? NW_RAW : 4);
  static constexpr int THREADS = NW * 32;
  static constexpr size_t SMEM = (size_t)NW * BB_RING * STAGE_B;
};

template <int ROWB>
__device__ __forceinline__ void bb_stage_row(unsigned char* dst, const float* src, int lane) {
#pragma unroll
  for (int t = 0; t * 32 < ROWB / 16; ++t) {
    const int piece = t * 32 + lane;
    if (piece < ROWB / 16) dl_cp_async16(dst + piece * 16, src + piece * 4);
  }
}

struct CMeta {
  int row, col, ks;
  float base;
};

template <class M>
__global__ void __launch_bounds__(AccumCfg<M>::THREADS, 1)
k_bwd_accum_stream(DlGraphDev g, const float* __restrict__ Z, const unsigned char* __restrict__ kstar,
                   const float* __restrict__ base, float T, float* __restrict__ dZ,
                   float* __restrict__ carry) {
  using C = AccumCfg<M>;
  constexpr int K = M::K, D = M::D, NP = M::NP, EB = M::EB, LP = M::LP, FPP = M::FPP;
  constexpr int ROWB = C::ROWB, STAGE_B = C::STAGE_B;
  constexpr bool DENSE = (M::L == M::LP) && (M::K % M::FPP == 0);
  static_assert(DL_HS == 4, "one stage = one sub-block of 4 entries");
  extern __shared__ __align__(128) unsigned char dl_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = dl_smem_raw + (size_t)warp * BB_RING * STAGE_B;
  const long long gw = (long long)blockIdx.x * C::NW + warp;
  const long long RE = (long long)DL_CH * DL_RANGE;

  int off[NP];
  bool act[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { off[p] = M::offset(lane, p); act[p] = M::active(lane, p); }
  const int my_e = M::edge_of_lane(lane);
  const int gsrc = lane & (EB - 1);
  const int gbase = lane & ~(LP - 1);
  const bool unit_T = (T == 1.0f);

  DlChunkStream cs;
  cs.init(g.nnz, (long long)gridDim.x * C::NW);
  auto load_meta = [&](long long cc, CMeta& m) {
    m.row = -1; m.col = 0; m.ks = 0; m.base = 0.0f;
    if (cc >= 0) {
      const long long e = cc * DL_CH + lane;
      if (e < g.nnz) {
        m.row = __ldg(g.erow + e); m.col = __ldg(g.col + e); m.ks = __ldg(kstar + e); m.base = __ldg(base + e);
      }
    }
  };
  auto issue_stage = [&](unsigned char* st, const CMeta& m, int q) {
#pragma unroll
    for (int e = 0; e < DL_HS; ++e) {
      const int rr = __shfl_sync(DL_FULL, m.row, q * DL_HS + e);
      const int cc = __shfl_sync(DL_FULL, m.col, q * DL_HS + e);
      if (rr >= 0) bb_stage_row<ROWB>(st + e * ROWB, Z + (long long)cc * D, lane);
    }
    const int prev = __shfl_up_sync(DL_FULL, m.row, 1);
    const bool start = (lane / DL_HS) == q && m.row >= 0 && ((lane % DL_HS) == 0 || prev != m.row);
    unsigned smask = __ballot_sync(DL_FULL, start);
#pragma unroll
    for (int o = 0; o < BB_OWN; ++o) {
      if (smask) {
        const int pos = __ffs(smask) - 1;
        smask &= smask - 1;
        const long long node = g.row_base + __shfl_sync(DL_FULL, m.row, pos);
        bb_stage_row<ROWB>(st + (DL_HS + o) * ROWB, Z + node * D, lane);
      }
    }
  };

  float4 dz[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) dz[p] = dl_zero4();
  int cur_row = -1;
  bool first_run = true, head_open = false, tail_open = false;
  long long cur_range = -1;
  auto flush = [&](bool at_range_end) {
    if (cur_row >= 0) {
      const bool to_head = first_run && head_open;
      const bool to_tail = !to_head && at_range_end && tail_open;
      if (to_head || to_tail) {
        float* dst = carry + (cur_range * 2 + (to_tail ? 1 : 0)) * D;
#pragma unroll
        for (int p = 0; p < NP; ++p)
          if (act[p]) *reinterpret_cast<float4*>(dst + off[p]) = dz[p];
      } else {
        const long long node = g.row_base + cur_row;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          if (!act[p]) continue;
          float4* dp = reinterpret_cast<float4*>(dZ + node * D + off[p]);
          float4 cur = *dp;
          cur.x = __fadd_rn(cur.x, dz[p].x); cur.y = __fadd_rn(cur.y, dz[p].y);
          cur.z = __fadd_rn(cur.z, dz[p].z); cur.w = __fadd_rn(cur.w, dz[p].w);
          *dp = cur;
        }
      }
      first_run = false;
    }
    cur_row = -1;
#pragma unroll
    for (int p = 0; p < NP; ++p) dz[p] = dl_zero4();
  };

  long long c = cs.first(gw);
  CMeta mA, mB;
  load_meta(c, mA);
  long long cn = cs.next(c);
  load_meta(cn, mB);
#pragma unroll
  for (int pq = 0; pq < BB_RING - 1; ++pq) {
    issue_stage(ring + pq * STAGE_B, mA, pq);
    dl_cp_async_commit();
  }
  int rslot = 0;

  while (c >= 0) {
    const long long rg = c / DL_RANGE;
    if (rg != cur_range) {
      if (cur_range >= 0) flush(true);
      cur_range = rg;
      first_run = true;
      const long long R0 = rg * RE, R1 = min(R0 + RE, g.nnz);
      head_open = R0 > 0 && __ldg(g.erow + R0 - 1) == __ldg(g.erow + R0);
      tail_open = R1 < g.nnz && __ldg(g.erow + R1) == __ldg(g.erow + R1 - 1);
    }
    const int prevA = __shfl_up_sync(DL_FULL, mA.row, 1);
    const bool startA = mA.row >= 0 && ((lane % DL_HS) == 0 || prevA != mA.row);
    const unsigned smaskA = __ballot_sync(DL_FULL, startA);
    const unsigned qbits = ((1u << DL_HS) - 1u) << ((lane / DL_HS) * DL_HS);
    const int rankA = __popc(smaskA & qbits & (0xffffffffu >> (31 - lane))) - 1;
    const unsigned vmaskA = __ballot_sync(DL_FULL, mA.row >= 0);
    const bool allownA = __all_sync(DL_FULL, rankA < BB_OWN);

#pragma unroll 1
    for (int q = 0; q < DL_QPC; ++q) {
      int islot = rslot + (BB_RING - 1);
      if (islot >= BB_RING) islot -= BB_RING;
      if (q < DL_QPC - (BB_RING - 1)) issue_stage(ring + islot * STAGE_B, mA, q + (BB_RING - 1));
      else issue_stage(ring + islot * STAGE_B, mB, q + (BB_RING - 1) - DL_QPC);
      dl_cp_async_commit();
      dl_cp_async_wait<BB_RING - 1>();
      __syncwarp();
      const unsigned char* st = ring + rslot * STAGE_B;
      const int cnt = __popc((vmaskA >> (q * DL_HS)) & ((1u << DL_HS) - 1u));
      if (cnt > 0) {
        int re[EB], rk[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          re[e] = __shfl_sync(DL_FULL, mA.row, q * DL_HS + e);
          rk[e] = __shfl_sync(DL_FULL, rankA, q * DL_HS + e);
        }
        float4 zj[EB][NP];
        float ev[NP];
        const bool full = allownA && cnt == EB;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          float part[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            float4 zi = dl_zero4();
            zj[e][p] = dl_zero4();
            if (full) {
              if (DENSE || act[p]) {
                zj[e][p] = dl_lds4(st + e * ROWB + off[p] * 4);
                zi = dl_lds4(st + (DL_HS + rk[e]) * ROWB + off[p] * 4);
              }
            } else if (e < cnt && act[p]) {
              zj[e][p] = dl_lds4(st + e * ROWB + off[p] * 4);
              if (rk[e] < BB_OWN) zi = dl_lds4(st + (DL_HS + rk[e]) * ROWB + off[p] * 4);
              else zi = dl_ldg4(Z + (g.row_base + re[e]) * D + off[p]);
            }
            part[e] = dl_chunk_dot(zi, zj[e][p]);
          }
          float qv = dl_reduce_scatter<M>(part, lane);
          if (!unit_T) qv = __fdiv_rn(qv, T);
          ev[p] = dl_expf(qv);
        }
        const bool valid = my_e < cnt;
        const int ks = __shfl_sync(DL_FULL, mA.ks, (q * DL_HS + my_e) & 31);
        float basec = __shfl_sync(DL_FULL, mA.base, (q * DL_HS + my_e) & 31);
        basec = valid ? basec : 0.0f;
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float ek = __shfl_sync(DL_FULL, ev[k / FPP], (k % FPP) * LP + gsrc);
          sum = (k == 0) ? ek : __fadd_rn(sum, ek);
        }
        const float rsum = __fdiv_rn(1.0f, sum);
        float cfe[NP][EB];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const float a_own = __fmul_rn(ev[p], rsum);
          const float ind = (M::factor(lane, p) == ks) ? 1.0f : 0.0f;
          const float coef_own = __fmul_rn(basec, __fsub_rn(ind, a_own));
#pragma unroll
          for (int e = 0; e < EB; ++e) cfe[p][e] = __shfl_sync(DL_FULL, coef_own, gbase + M::lane_of_edge(e));
        }
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          if (e < cnt) {
            if (re[e] != cur_row) { flush(false); cur_row = re[e]; }
#pragma unroll
            for (int p = 0; p < NP; ++p) dl_fma4(dz[p], cfe[p][e], zj[e][p]);
          }
        }
      }
      __syncwarp();
      rslot = (rslot + 1 == BB_RING) ? 0 : rslot + 1;
    }
    c = cn;
    mA = mB;
    cn = cs.next(c);
    load_meta(cn, mB);
  }
  if (cur_range >= 0) flush(true);
  dl_cp_async_wait<0>();
}

template <class M>
int launch_bwd_split(const DlGraphDev& g, const float* Z, const float* G, const unsigned char* kstar,
                     const float* w, const float* s, const float* r, float omb, float T, float* dZ,
                     float* carry, float* base, cudaStream_t st) {
  using A = BaseCfg<M>;
  using B = AccumCfg<M>;
  if (!B::OK || A::SMEM > 200 * 1024) return -1000;
  int dev = 0, sms = 0, per_sm = 0;
  DL_CUDA_TRY(cudaGetDevice(&dev));
  DL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long n_chunks = (g.nnz + DL_CH - 1) / DL_CH;
  const long long n_ranges = (n_chunks + DL_RANGE - 1) / DL_RANGE;
  {
    if (A::SMEM > 48 * 1024)
      DL_CUDA_TRY(cudaFuncSetAttribute(k_bwd_base_stream<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)A::SMEM));
    DL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bwd_base_stream<M>, BA_WARPS * 32,
                                                              A::SMEM));
    if (per_sm < 1) per_sm = 1;
    long long grid = (n_ranges + BA_WARPS - 1) / BA_WARPS;
    if (grid > (long long)sms * per_sm) grid = (long long)sms * per_sm;
    if (grid < 1) grid = 1;
    k_bwd_base_stream<M><<<(int)grid, BA_WARPS * 32, A::SMEM, st>>>(g, Z, G, kstar, w, s, r, omb, T, base);
    DL_LAUNCH_CHECK();
  }
  {
    DL_CUDA_TRY(cudaFuncSetAttribute(k_bwd_accum_stream<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)B::SMEM));
    long long grid = (n_ranges + B::NW - 1) / B::NW;
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    k_bwd_accum_stream<M><<<(int)grid, B::THREADS, B::SMEM, st>>>(g, Z, kstar, base, T, dZ, carry);
    DL_LAUNCH_CHECK();
  }
  return DL_OK;
}

}  // namespace

// scratch layout: [carries + chain scratch: 3 * n_ranges * D floats][base: nnz floats]
int dl_launch_bwd_edges_split(const DlGraphDev& g, const float* Z, const float* G,
                              const unsigned char* kstar, const float* w, const float* s, const float* r,
                              int K, int d, float omb, float T, float* dZ, float* scratch, cudaStream_t st) {
  if (!g.erow || g.nnz == 0 || !scratch || !w) return -1000;
  const long long RE = (long long)DL_CH * DL_RANGE;
  const long long n_ranges = (g.nnz + RE - 1) / RE;
  float* base = scratch + (size_t)n_ranges * 3 * K * d;
  int rc = -1000;
#define BODY_MACRO(M) rc = launch_bwd_split<M>(g, Z, G, kstar, w, s, r, omb, T, dZ, scratch, base, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  if (rc != DL_OK) return rc;
  return dl_gather_chain_add(g, K, d, scratch, dZ, st);
}
