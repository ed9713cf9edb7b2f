// bwd_stream.cu -- backward pass 2 of the factor attention (streaming version).
//
// [ref: autograd of model.py:56-75]  per CSR entry (i,j), k = kstar(i,j):
//   dwsum = c_ij/s[j,k] + c_ji/s[i,k] - r[i,k] - r[j,k],   c_ij = (1-beta) <G[i,k], Z[j,k]>
//   dZ[i,kk,:] += dwsum * w / T * ((kk==k) - a[kk]) * Z[j,kk,:]          for all kk
// with a[] = softmax over factors recomputed from Z in canonical arithmetic (bit-identical to the
// forward) and kstar read back.  Pass 1 (gather_stream.cu, MODE 1) has already written r and
// beta*G + T_ into dZ.
//
// Structure: the same balanced chunk streams as attn_stream.cu.  Per 4-entry stage the warp stages
// with cp.async: the neighbour rows Z[j] (512 B), the routed slices G[j,k] (64 B), and -- at the
// first entry of a row run -- the own rows Z[i], G[i] and s[i,:], r[i,:].  The per-entry gathers
// s[j,k], r[j,k] ride with the chunk metadata, one chunk ahead.  The row's dZ accumulator lives in
// registers (lane (k, g) owns chunk g of factor k), is accumulated in CSR order and is added to
// dZ[i] when the row id changes; rows cut by a range boundary go through the carry / chain
// mechanism of gather_stream.cu (mode 3).
//
// HBM bytes per entry (D=128, d=16): 4 + 4 + 1 (col, row, kstar) + 512 + 64 + 4 + 4.
#include "dl_dispatch.cuh"
#include "dl_stream.cuh"

namespace {

constexpr int BW_OWN = 2;   // staged own-row slots per stage
constexpr int BW_RING = 2;  // stages per warp ring (registers cap the kernel at 16 warps/SM anyway)

template <class M>
struct BwdStreamCfg {
  static constexpr int ROWB = M::D * 4;
  static constexpr int SLB = M::d * 4;
  static constexpr int SRB = ((2 * M::K * 4 + 15) / 16) * 16;             // s[i,:], r[i,:]
  static constexpr int OWN_B = 2 * ROWB + SRB;                             // Z[i], G[i], s/r
  static constexpr int STAGE_B = DL_HS * (ROWB + SLB) + BW_OWN * OWN_B;
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int NW_RAW = BUDGET / (BW_RING * STAGE_B);
  static constexpr bool OK = NW_RAW >= 4 && M::EB == 4 && 2 * M::K <= 32;
  static constexpr int NW = NW_RAW >= 16 ? 16 : (NW_RAW >= 4 ? NW_RAW : 4);
  static constexpr int THREADS = NW * 32;
  static constexpr size_t SMEM = (size_t)NW * BW_RING * STAGE_B;
};

__device__ __forceinline__ void dl_cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dl_smem_u32(smem_dst)), "l"(gmem_src)
               : "memory");
}

template <int ROWB>
__device__ __forceinline__ void bw_stage_row(unsigned char* dst, const float* src, int lane) {
#pragma unroll
  for (int t = 0; t * 32 < ROWB / 16; ++t) {
    const int piece = t * 32 + lane;
    if (piece < ROWB / 16) dl_cp_async16(dst + piece * 16, src + piece * 4);
  }
}

struct BMeta {
  int row, col, ks;
  float sj, rj;
};

template <class M>
__global__ void __launch_bounds__(BwdStreamCfg<M>::THREADS, 1)
k_bwd_edges_stream(DlGraphDev g, const float* __restrict__ Z, const float* __restrict__ G,
                   const unsigned char* __restrict__ kstar, const float* __restrict__ s,
                   const float* __restrict__ r, float omb, float T, float* __restrict__ dZ,
                   float* __restrict__ carry) {
  using C = BwdStreamCfg<M>;
  constexpr int K = M::K, d = M::d, D = M::D, NP = M::NP, EB = M::EB, LP = M::LP, FPP = M::FPP, L = M::L;
  constexpr int ROWB = C::ROWB, SLB = C::SLB, STAGE_B = C::STAGE_B, OWN_B = C::OWN_B;
  constexpr int NBG_OFF = DL_HS * ROWB;                 // routed G slices of the stage
  constexpr int OWN_OFF = DL_HS * (ROWB + SLB);         // own-row slots
  static_assert(DL_HS == 4, "one stage = one sub-block of 4 entries");
  constexpr bool DENSE = (M::L == M::LP) && (M::K % M::FPP == 0);   // every lane active in every pass
  extern __shared__ __align__(128) unsigned char dl_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = dl_smem_raw + (size_t)warp * BW_RING * STAGE_B;
  const long long gw = (long long)blockIdx.x * C::NW + warp;
  const long long RE = (long long)DL_CH * DL_RANGE;

  int off[NP];
  bool act[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { off[p] = M::offset(lane, p); act[p] = M::active(lane, p); }
  const int my_e = M::edge_of_lane(lane);
  const int gsrc = lane & (EB - 1);
  const int gbase = lane & ~(LP - 1);
  const int slot = M::slot(lane), gg = M::g(lane), grp = lane / LP;
  const bool glane = gg < L;
  const bool unit_T = (T == 1.0f);

  DlChunkStream cs;
  cs.init(g.nnz, (long long)gridDim.x * C::NW, g.range_shift);

  auto load_meta = [&](long long cc, BMeta& m) {
    m.row = -1; m.col = 0; m.ks = 255; m.sj = 1.0f; m.rj = 0.0f;
    if (cc >= 0) {
      const long long e = cc * DL_CH + lane;
      if (e < g.nnz) { m.row = __ldg(g.erow + e); m.col = __ldg(g.col + e); m.ks = __ldg(kstar + e); }
    }
  };
  auto load_sr = [&](BMeta& m) {
    if (m.row >= 0) {
      m.sj = __ldg(s + (long long)m.col * K + m.ks);
      m.rj = __ldg(r + (long long)m.col * K + m.ks);
    }
  };
  // stage q of the chunk whose metadata is m: 4 neighbour rows, 4 routed G slices, and the own
  // rows (Z, G, s, r) of the first BW_OWN row runs of the stage
  auto issue_stage = [&](unsigned char* st, const BMeta& m, int q) {
#pragma unroll
    for (int e = 0; e < DL_HS; ++e) {
      const int rr = __shfl_sync(DL_FULL, m.row, q * DL_HS + e);
      const int cc = __shfl_sync(DL_FULL, m.col, q * DL_HS + e);
      if (rr >= 0) bw_stage_row<ROWB>(st + e * ROWB, Z + (long long)cc * D, lane);
    }
    {   // the routed slices: lane group grp copies the slice of entry rd*NG + grp (all 4 entries
        // with one instruction when d <= 32)
      constexpr int NG = 32 / LP;
#pragma unroll
      for (int rd = 0; rd * NG < DL_HS; ++rd) {
        const int ent = rd * NG + grp;
        const int src = q * DL_HS + (ent < DL_HS ? ent : 0);
        const int rr = __shfl_sync(DL_FULL, m.row, src);
        const int cc = __shfl_sync(DL_FULL, m.col, src);
        const int kk = __shfl_sync(DL_FULL, m.ks, src);
        if (ent < DL_HS && glane && rr >= 0)
          dl_cp_async16(st + NBG_OFF + ent * SLB + gg * 16, G + (long long)cc * D + kk * d + gg * 4);
      }
    }
    const int prev = __shfl_up_sync(DL_FULL, m.row, 1);
    const bool start = (lane / DL_HS) == q && m.row >= 0 && ((lane % DL_HS) == 0 || prev != m.row);
    unsigned smask = __ballot_sync(DL_FULL, start);
#pragma unroll
    for (int o = 0; o < BW_OWN; ++o) {
      if (smask) {
        const int pos = __ffs(smask) - 1;
        smask &= smask - 1;
        const long long node = g.row_base + __shfl_sync(DL_FULL, m.row, pos);
        unsigned char* ow = st + OWN_OFF + o * OWN_B;
        bw_stage_row<ROWB>(ow, Z + node * D, lane);
        bw_stage_row<ROWB>(ow + ROWB, G + node * D, lane);
        if (lane < K) dl_cp_async4(ow + 2 * ROWB + lane * 4, s + node * K + lane);
        else if (lane < 2 * K) dl_cp_async4(ow + 2 * ROWB + lane * 4, r + node * K + (lane - K));
      }
    }
  };

  // per-row accumulator and range bookkeeping (see gather_stream.cu)
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
  BMeta mA, mB, mC;
  load_meta(c, mA);
  load_sr(mA);
  long long cn = cs.next(c);
  load_meta(cn, mB);
#pragma unroll
  for (int pq = 0; pq < BW_RING - 1; ++pq) {
    issue_stage(ring + pq * STAGE_B, mA, pq);
    dl_cp_async_commit();
  }
  int rslot = 0;

  while (c >= 0) {
    const long long cnn = cs.next(cn);
    load_meta(cnn, mC);       // two chunks ahead
    load_sr(mB);              // s[j,k], r[j,k] of the next chunk (its col / kstar arrived a chunk ago)

    const long long rg = c >> g.range_shift;
    if (rg != cur_range) {
      if (cur_range >= 0) flush(true);
      cur_range = rg;
      first_run = true;
      const long long R0 = rg * RE, R1 = min(R0 + RE, g.nnz);
      head_open = R0 > 0 && __ldg(g.erow + R0 - 1) == __ldg(g.erow + R0);
      tail_open = R1 < g.nnz && __ldg(g.erow + R1) == __ldg(g.erow + R1 - 1);
    }
    // own-row slot of every entry inside its stage
    const int prevA = __shfl_up_sync(DL_FULL, mA.row, 1);
    const bool startA = mA.row >= 0 && ((lane % DL_HS) == 0 || prevA != mA.row);
    const unsigned smaskA = __ballot_sync(DL_FULL, startA);
    const unsigned qbits = ((1u << DL_HS) - 1u) << ((lane / DL_HS) * DL_HS);
    const int rankA = __popc(smaskA & qbits & (0xffffffffu >> (31 - lane))) - 1;
    const unsigned vmaskA = __ballot_sync(DL_FULL, mA.row >= 0);
    const bool allownA = __all_sync(DL_FULL, rankA < BW_OWN);

#pragma unroll 1
    for (int q = 0; q < DL_QPC; ++q) {
      int islot = rslot + (BW_RING - 1);
      if (islot >= BW_RING) islot -= BW_RING;
      if (q < DL_QPC - (BW_RING - 1)) issue_stage(ring + islot * STAGE_B, mA, q + (BW_RING - 1));
      else issue_stage(ring + islot * STAGE_B, mB, q + (BW_RING - 1) - DL_QPC);
      dl_cp_async_commit();
      dl_cp_async_wait<BW_RING - 1>();
      __syncwarp();
      const unsigned char* st = ring + rslot * STAGE_B;
      const int cnt = __popc((vmaskA >> (q * DL_HS)) & ((1u << DL_HS) - 1u));
      if (cnt > 0) {
        // per-entry scalars of the 4 entries of this stage
        int re[EB], ke[EB], rk[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          re[e] = __shfl_sync(DL_FULL, mA.row, q * DL_HS + e);
          ke[e] = __shfl_sync(DL_FULL, mA.ks, q * DL_HS + e);
          rk[e] = __shfl_sync(DL_FULL, rankA, q * DL_HS + e);
        }
        float4 zj[EB][NP];
        float ev[NP];
        float pij[EB], pji[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) { pij[e] = 0.0f; pji[e] = 0.0f; }
        // fast path: a full stage whose own rows are all staged needs no per-entry validity handling
        const bool full = allownA && cnt == EB;
        if (full) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            float part[EB];
#pragma unroll
            for (int e = 0; e < EB; ++e) {
              float4 zi = dl_zero4(), gi = dl_zero4();
              zj[e][p] = dl_zero4();
              if (DENSE || act[p]) {
                const unsigned char* ow = st + OWN_OFF + rk[e] * OWN_B + off[p] * 4;
                zj[e][p] = dl_lds4(st + e * ROWB + off[p] * 4);
                zi = dl_lds4(ow);
                if (ke[e] == M::factor(lane, p)) {
                  gi = dl_lds4(ow + ROWB);
                  const float4 gje = dl_lds4(st + NBG_OFF + e * SLB + gg * 16);
                  pij[e] = dl_chunk_dot(gi, zj[e][p]);
                  pji[e] = dl_chunk_dot(gje, zi);
                }
              }
              part[e] = dl_chunk_dot(zi, zj[e][p]);
            }
            float qv = dl_reduce_scatter<M>(part, lane);
            if (!unit_T) qv = __fdiv_rn(qv, T);
            ev[p] = dl_expf(qv);
          }
        } else {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            float part[EB];
#pragma unroll
            for (int e = 0; e < EB; ++e) {
              float4 zi = dl_zero4(), gi = dl_zero4();
              zj[e][p] = dl_zero4();
              if (e < cnt && act[p]) {
                zj[e][p] = dl_lds4(st + e * ROWB + off[p] * 4);
                if (rk[e] < BW_OWN) {
                  zi = dl_lds4(st + OWN_OFF + rk[e] * OWN_B + off[p] * 4);
                  gi = dl_lds4(st + OWN_OFF + rk[e] * OWN_B + ROWB + off[p] * 4);
                } else {
                  zi = dl_ldg4(Z + (g.row_base + re[e]) * D + off[p]);
                  gi = dl_ldg4(G + (g.row_base + re[e]) * D + off[p]);
                }
                // c_ij, c_ji partials exist only on the lanes that own the routed factor of entry e
                if (ke[e] == M::factor(lane, p)) {
                  const float4 gje = dl_lds4(st + NBG_OFF + e * SLB + gg * 16);
                  pij[e] = dl_chunk_dot(gi, zj[e][p]);
                  pji[e] = dl_chunk_dot(gje, zi);
                }
              }
              part[e] = dl_chunk_dot(zi, zj[e][p]);
            }
            float qv = dl_reduce_scatter<M>(part, lane);
            if (!unit_T) qv = __fdiv_rn(qv, T);
            ev[p] = dl_expf(qv);
          }
        }
        const bool valid = my_e < cnt;
        int ks = __shfl_sync(DL_FULL, mA.ks, (q * DL_HS + my_e) & 31);
        ks = valid ? ks : 0;
        float sum = 0.0f, eks = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float ek = __shfl_sync(DL_FULL, ev[k / FPP], (k % FPP) * LP + gsrc);
          sum = (k == 0) ? ek : __fadd_rn(sum, ek);
          if (k == ks) eks = ek;
        }
        const float rsum = __fdiv_rn(1.0f, sum);
        const float wv = __fmul_rn(eks, rsum);           // w[e] up to one rounding
        const float rij = dl_reduce_scatter<M>(pij, lane);
        const float rji = dl_reduce_scatter<M>(pji, lane);
        const int ksrc = (ks % FPP) * LP + gsrc;
        const float cij = __fmul_rn(omb, __shfl_sync(DL_FULL, rij, ksrc));
        const float cji = __fmul_rn(omb, __shfl_sync(DL_FULL, rji, ksrc));
        const float sjv = __shfl_sync(DL_FULL, mA.sj, (q * DL_HS + my_e) & 31);
        const float rjv = __shfl_sync(DL_FULL, mA.rj, (q * DL_HS + my_e) & 31);
        // s[i,ks], r[i,ks] of my entry's own row
        const int myrk = __shfl_sync(DL_FULL, rankA, (q * DL_HS + my_e) & 31);
        const int myrow = __shfl_sync(DL_FULL, mA.row, (q * DL_HS + my_e) & 31);
        float siv = 1.0f, riv = 0.0f;
        if (valid) {
          if (myrk < BW_OWN) {
            const float* sr = reinterpret_cast<const float*>(st + OWN_OFF + myrk * OWN_B + 2 * ROWB);
            siv = sr[ks];
            riv = sr[K + ks];
          } else {
            siv = __ldg(s + (g.row_base + myrow) * K + ks);
            riv = __ldg(r + (g.row_base + myrow) * K + ks);
          }
        }
        float dws = __fadd_rn(__fdiv_rn(cij, sjv), __fdiv_rn(cji, siv));
        dws = __fsub_rn(dws, riv);
        dws = __fsub_rn(dws, rjv);
        float basec = __fmul_rn(dws, wv);
        if (!unit_T) basec = __fdiv_rn(basec, T);
        basec = valid ? basec : 0.0f;
        float cfe[NP][EB];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const float a_own = __fmul_rn(ev[p], rsum);
          const float ind = (M::factor(lane, p) == ks) ? 1.0f : 0.0f;
          const float coef_own = __fmul_rn(basec, __fsub_rn(ind, a_own));
#pragma unroll
          for (int e = 0; e < EB; ++e) cfe[p][e] = __shfl_sync(DL_FULL, coef_own, gbase + M::lane_of_edge(e));
        }
        // accumulate in CSR order, switching rows where the row id changes
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
      rslot = (rslot + 1 == BW_RING) ? 0 : rslot + 1;
    }
    c = cn; cn = cnn;
    mA = mB; mB = mC;
  }
  if (cur_range >= 0) flush(true);
  dl_cp_async_wait<0>();
}

template <class M>
int launch_bwd_stream(const DlGraphDev& g, const float* Z, const float* G, const unsigned char* kstar,
                      const float* s, const float* r, float omb, float T, float* dZ, float* carry,
                      cudaStream_t st) {
  using C = BwdStreamCfg<M>;
  if (!C::OK) return -1000;
  int dev = 0, sms = 0;
  DL_CUDA_TRY(cudaGetDevice(&dev));
  DL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DL_CUDA_TRY(cudaFuncSetAttribute(k_bwd_edges_stream<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)C::SMEM));
  const long long n_chunks = (g.nnz + DL_CH - 1) / DL_CH;
  const long long n_ranges = (n_chunks + DL_RANGE - 1) / DL_RANGE;
  long long grid = (n_ranges + C::NW - 1) / C::NW;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  k_bwd_edges_stream<M><<<(int)grid, C::THREADS, C::SMEM, st>>>(g, Z, G, kstar, s, r, omb, T, dZ, carry);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

}  // namespace

// returns -1000 when (K, d) has no streaming instantiation; scratch as for the streaming gather
int dl_launch_bwd_edges_stream(const DlGraphDev& g, const float* Z, const float* G,
                               const unsigned char* kstar, const float* s, const float* r, int K, int d,
                               float omb, float T, float* dZ, float* scratch, cudaStream_t st) {
  if (!g.erow || g.nnz == 0 || !scratch) return -1000;
  int rc = -1000;
#define BODY_MACRO(M) rc = launch_bwd_stream<M>(g, Z, G, kstar, s, r, omb, T, dZ, scratch, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  if (rc != DL_OK) return rc;
  return dl_gather_chain_add(g, K, d, scratch, dZ, st);
}
