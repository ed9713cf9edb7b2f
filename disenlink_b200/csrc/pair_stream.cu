// pair_stream.cu -- backward of the factor-weighted link decoder, streaming version.
//
// [ref: autograd of model.py:109-113]  node-major over the incidence lists of a pair batch
// (dl_pair_incidence): for node n and every pair (n,o) it takes part in, with dS = dL/dlogit,
//   e_k = exp(<Z[n,k],Z[o,k]>/T),  dH[n,k] += dS e_k H[o,k],  dZ[n,k] += dS e_k <H[n,k],H[o,k]>/T Z[o,k]
// Same structure as the other streaming kernels (dl_stream.cuh): the incidences are cut into
// balanced 32-entry chunks, each warp keeps one 4-entry stage (Z[o], H[o] rows, plus Z[n], H[n] at
// the start of a node's run) in flight with cp.async while it computes on the other, accumulates
// dZ[n] / dH[n] in registers in incidence order and writes the row when the node id changes; nodes
// cut by a range boundary go through the carry / chain mechanism (mode 4), nodes without
// incidences get zeros.  No atomics: every row is written exactly once.
#include "dl_dispatch.cuh"
#include "dl_stream.cuh"

namespace {

constexpr int PS_RING = 2;
constexpr int PS_OWN = 2;

template <class M>
struct PairStreamCfg {
  static constexpr int ROWB = M::D * 4;
  static constexpr int STAGE_B = (DL_HS + PS_OWN) * 2 * ROWB;      // (Z, H) rows of 4 others + own slots
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int NW_RAW = BUDGET / (PS_RING * STAGE_B);
  static constexpr bool OK = NW_RAW >= 4 && M::EB == 4 && M::NP <= 4;   // NP = 5 spills badly: row-per-warp kernel
  // shapes whose row needs several passes over the warp hold 2x the registers: fewer, fatter warps
  static constexpr int NW_CAP = (M::NP >= 2) ? 8 : 16;
  static constexpr int NW = NW_RAW >= NW_CAP ? NW_CAP : (NW_RAW >= 4 ? NW_RAW : 4);
  static constexpr int THREADS = NW * 32;
  static constexpr size_t SMEM = (size_t)NW * PS_RING * STAGE_B;
};

template <int ROWB>
__device__ __forceinline__ void ps_stage_row(unsigned char* dst, const float* src, int lane) {
#pragma unroll
  for (int t = 0; t * 32 < ROWB / 16; ++t) {
    const int piece = t * 32 + lane;
    if (piece < ROWB / 16) dl_cp_async16(dst + piece * 16, src + piece * 4);
  }
}

struct PMeta {
  int row, col, pid;
  float ds;
};

template <class M>
__global__ void __launch_bounds__(PairStreamCfg<M>::THREADS, 1)
k_pair_bwd_stream(DlGraphDev g, const int* __restrict__ inc_pair, const float* __restrict__ Z,
                  const float* __restrict__ H, const float* __restrict__ dS, float T,
                  float* __restrict__ dZ, float* __restrict__ dH, float* __restrict__ carry) {
  using C = PairStreamCfg<M>;
  constexpr int D = M::D, NP = M::NP, EB = M::EB, LP = M::LP;
  constexpr int ROWB = C::ROWB, STAGE_B = C::STAGE_B;
  constexpr bool DENSE = (M::L == M::LP) && (M::K % M::FPP == 0);
  static_assert(DL_HS == 4, "one stage = one sub-block of 4 entries");
  extern __shared__ __align__(128) unsigned char dl_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = dl_smem_raw + (size_t)warp * PS_RING * STAGE_B;
  const long long gw = (long long)blockIdx.x * C::NW + warp;
  const long long RE = (long long)DL_CH * DL_RANGE;

  int off[NP];
  bool act[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { off[p] = M::offset(lane, p); act[p] = M::active(lane, p); }
  const int my_e = M::edge_of_lane(lane);
  const int gbase = lane & ~(LP - 1);
  const bool unit_T = (T == 1.0f);

  DlChunkStream cs;
  cs.init(g.nnz, (long long)gridDim.x * C::NW, g.range_shift);
  auto load_meta = [&](long long cc, PMeta& m) {
    m.row = -1; m.col = 0; m.pid = 0; m.ds = 0.0f;
    if (cc >= 0) {
      const long long e = cc * DL_CH + lane;
      if (e < g.nnz) { m.row = __ldg(g.erow + e); m.col = __ldg(g.col + e); m.pid = __ldg(inc_pair + e); }
    }
  };
  auto load_ds = [&](PMeta& m) {
    if (m.row >= 0) m.ds = __ldg(dS + m.pid);
  };
  // stage layout: [e][Z row | H row] for the 4 other endpoints, then [o][Z row | H row] own slots
  auto issue_stage = [&](unsigned char* st, const PMeta& m, int q) {
#pragma unroll
    for (int e = 0; e < DL_HS; ++e) {
      const int rr = __shfl_sync(DL_FULL, m.row, q * DL_HS + e);
      const long long cc = __shfl_sync(DL_FULL, m.col, q * DL_HS + e);
      if (rr >= 0) {
        ps_stage_row<ROWB>(st + e * 2 * ROWB, Z + cc * D, lane);
        ps_stage_row<ROWB>(st + e * 2 * ROWB + ROWB, H + cc * D, lane);
      }
    }
    const int prev = __shfl_up_sync(DL_FULL, m.row, 1);
    const bool start = (lane / DL_HS) == q && m.row >= 0 && ((lane % DL_HS) == 0 || prev != m.row);
    unsigned smask = __ballot_sync(DL_FULL, start);
#pragma unroll
    for (int o = 0; o < PS_OWN; ++o) {
      if (smask) {
        const int pos = __ffs(smask) - 1;
        smask &= smask - 1;
        const long long node = g.row_base + __shfl_sync(DL_FULL, m.row, pos);
        ps_stage_row<ROWB>(st + (DL_HS + o) * 2 * ROWB, Z + node * D, lane);
        ps_stage_row<ROWB>(st + (DL_HS + o) * 2 * ROWB + ROWB, H + node * D, lane);
      }
    }
  };

  float4 az[NP], ah[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) az[p] = ah[p] = dl_zero4();
  int cur_row = -1;
  bool first_run = true, head_open = false, tail_open = false;
  long long cur_range = -1;
  auto flush = [&](bool at_range_end) {
    if (cur_row >= 0) {
      const bool to_head = first_run && head_open;
      const bool to_tail = !to_head && at_range_end && tail_open;
      float* dz_dst;
      float* dh_dst;
      if (to_head || to_tail) {
        dz_dst = carry + (cur_range * 2 + (to_tail ? 1 : 0)) * 2 * D;
        dh_dst = dz_dst + D;
      } else {
        dz_dst = dZ + (g.row_base + cur_row) * D;
        dh_dst = dH + (g.row_base + cur_row) * D;
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (!act[p]) continue;
        *reinterpret_cast<float4*>(dz_dst + off[p]) = az[p];
        *reinterpret_cast<float4*>(dh_dst + off[p]) = ah[p];
        if (!(to_head || to_tail)) {
#pragma unroll 1
          for (int q = 0; q < g.n_peer_out; ++q)              // the all-gather of dH rides on the kernel
            *reinterpret_cast<float4*>(g.peer_out[q] + (g.row_base + cur_row) * D + off[p]) = ah[p];
        }
      }
      first_run = false;
    }
    cur_row = -1;
#pragma unroll
    for (int p = 0; p < NP; ++p) az[p] = ah[p] = dl_zero4();
  };

  long long c = cs.first(gw);
  PMeta mA, mB, mC;
  load_meta(c, mA);
  load_ds(mA);
  long long cn = cs.next(c);
  load_meta(cn, mB);
#pragma unroll
  for (int pq = 0; pq < PS_RING - 1; ++pq) {
    issue_stage(ring + pq * STAGE_B, mA, pq);
    dl_cp_async_commit();
  }
  int rslot = 0;

  while (c >= 0) {
    const long long cnn = cs.next(cn);
    load_meta(cnn, mC);
    load_ds(mB);
    const long long rg = c >> g.range_shift;
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
    const bool allownA = __all_sync(DL_FULL, rankA < PS_OWN);

#pragma unroll 1
    for (int q = 0; q < DL_QPC; ++q) {
      int islot = rslot + (PS_RING - 1);
      if (islot >= PS_RING) islot -= PS_RING;
      if (q < DL_QPC - (PS_RING - 1)) issue_stage(ring + islot * STAGE_B, mA, q + (PS_RING - 1));
      else issue_stage(ring + islot * STAGE_B, mB, q + (PS_RING - 1) - DL_QPC);
      dl_cp_async_commit();
      dl_cp_async_wait<PS_RING - 1>();
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
        const bool full = allownA && cnt == EB;
        const bool valid = my_e < cnt;
        float ds = __shfl_sync(DL_FULL, mA.ds, (q * DL_HS + my_e) & 31);
        ds = valid ? ds : 0.0f;
        float4 zo[EB][NP], ho[EB][NP];
        float che[NP][EB], cze[NP][EB];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          float pz[EB], ph[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            float4 zn = dl_zero4(), hn = dl_zero4();
            zo[e][p] = dl_zero4();
            ho[e][p] = dl_zero4();
            if (full ? (DENSE || act[p]) : (e < cnt && act[p])) {
              zo[e][p] = dl_lds4(st + e * 2 * ROWB + off[p] * 4);
              ho[e][p] = dl_lds4(st + e * 2 * ROWB + ROWB + off[p] * 4);
              if (full || rk[e] < PS_OWN) {
                zn = dl_lds4(st + (DL_HS + rk[e]) * 2 * ROWB + off[p] * 4);
                hn = dl_lds4(st + (DL_HS + rk[e]) * 2 * ROWB + ROWB + off[p] * 4);
              } else {
                zn = dl_ldg4(Z + (g.row_base + re[e]) * D + off[p]);
                hn = dl_ldg4(H + (g.row_base + re[e]) * D + off[p]);
              }
            }
            pz[e] = dl_chunk_dot(zn, zo[e][p]);
            ph[e] = dl_chunk_dot(hn, ho[e][p]);
          }
          float qv = dl_reduce_scatter<M>(pz, lane);
          if (!unit_T) qv = __fdiv_rn(qv, T);
          const float hh = dl_reduce_scatter<M>(ph, lane);
          const float ek = dl_expf(qv);
          const float ch = __fmul_rn(ds, ek);
          float cz = __fmul_rn(ch, hh);
          if (!unit_T) cz = __fdiv_rn(cz, T);
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            const int srcl = gbase + M::lane_of_edge(e);
            che[p][e] = __shfl_sync(DL_FULL, ch, srcl);
            cze[p][e] = __shfl_sync(DL_FULL, cz, srcl);
          }
        }
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          if (e < cnt) {
            if (re[e] != cur_row) { flush(false); cur_row = re[e]; }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              dl_fma4(ah[p], che[p][e], ho[e][p]);
              dl_fma4(az[p], cze[p][e], zo[e][p]);
            }
          }
        }
      }
      __syncwarp();
      rslot = (rslot + 1 == PS_RING) ? 0 : rslot + 1;
    }
    c = cn; cn = cnn;
    mA = mB; mB = mC;
  }
  if (cur_range >= 0) flush(true);
  dl_cp_async_wait<0>();
}

template <class M>
int launch_pair_bwd_stream(const DlGraphDev& g, const int* inc_pair, const float* Z, const float* H,
                           const float* dS, float T, float* dZ, float* dH, float* carry, cudaStream_t st) {
  using C = PairStreamCfg<M>;
  if (!C::OK) return -1000;
  int dev = 0, sms = 0;
  DL_CUDA_TRY(cudaGetDevice(&dev));
  DL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DL_CUDA_TRY(cudaFuncSetAttribute(k_pair_bwd_stream<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)C::SMEM));
  const long long n_chunks = (g.nnz + DL_CH - 1) / DL_CH;
  const long long n_ranges = (n_chunks + DL_RANGE - 1) / DL_RANGE;
  long long grid = (n_ranges + C::NW - 1) / C::NW;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  k_pair_bwd_stream<M><<<(int)grid, C::THREADS, C::SMEM, st>>>(g, inc_pair, Z, H, dS, T, dZ, dH, carry);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

}  // namespace

// returns -1000 when (K, d) has no streaming instantiation; scratch: 3 * n_ranges * 2*K*d floats
int dl_launch_pair_bwd_stream(const DlGraphDev& g, const int* inc_pair, const float* Z, const float* H,
                              const float* dS, int K, int d, float T, float* dZ, float* dH, float* scratch,
                              cudaStream_t st) {
  if (!g.erow || g.nnz == 0 || !scratch) return -1000;
  int rc = -1000;
#define BODY_MACRO(M) rc = launch_pair_bwd_stream<M>(g, inc_pair, Z, H, dS, T, dZ, dH, scratch, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  if (rc != DL_OK) return rc;
  return dl_gather_chain_pair(g, K, d, scratch, dZ, dH, st);
}
