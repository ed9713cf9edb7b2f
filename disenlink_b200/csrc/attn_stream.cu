// attn_stream.cu -- per-edge K-factor attention with hard routing, streaming version.
//
// [ref: model.py:56-70]  per CSR entry (i,j): q_k = z_i^k.z_j^k / T, a = softmax_k(q),
// kstar = first argmax, w = a[kstar].  Same canonical arithmetic (dl_common.cuh) as every other
// kernel, so kstar / w are bit-identical to the row-per-warp kernels and to the CPU oracle.
//
// Structure (dl_stream.cuh): the CSR entries are cut into 32-entry chunks regardless of row
// boundaries (balanced, no hub path); every warp walks its own stream of chunks and keeps two
// 8-entry stages of neighbour rows (plus the entries' own rows at the start of each row run) in
// flight with cp.async into a private shared-memory ring while it does the K dot products, the
// softmax over factors and the routing of the oldest stage from shared memory.
// The kernel has no per-row state: the routed row sums s[i,k] are a separate streaming pass over
// (kstar, w) (k_row_sums), 5 bytes per entry.
//
// HBM bytes per entry (D=128): 4 (col) + 4 (row id) + 512 (z_j) + 5 (kstar, w) + 512/deg (z_i).
#include <math_constants.h>

#include "dl_dispatch.cuh"
#include "dl_stream.cuh"

namespace {

constexpr int AT_RING = 2;   // stages per warp ring: one in flight while one is consumed (32 warps/SM)

template <class M>
struct AttnStreamCfg {
  static constexpr int ROWB = M::D * 4;                           // bytes of one node row
  static constexpr int STAGE_B = (DL_HS + DL_OWNQ) * ROWB;        // neighbour rows + own-row slots
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int NW_RAW = BUDGET / (AT_RING * STAGE_B);
  static constexpr bool OK = NW_RAW >= 4;                         // else: row-per-warp kernels
  static constexpr int NW = NW_RAW >= 32 ? 32 : (OK ? NW_RAW : 4);   // warps per CTA
  static constexpr int THREADS = NW * 32;
  static constexpr size_t SMEM = (size_t)NW * AT_RING * STAGE_B;
};

// cp.async one node row (ROWB bytes) into shared memory, all 32 lanes cooperating
template <int ROWB>
__device__ __forceinline__ void dl_stage_row(unsigned char* dst, const float* src, int lane) {
#pragma unroll
  for (int t = 0; t * 32 < ROWB / 16; ++t) {
    const int piece = t * 32 + lane;
    if (piece < ROWB / 16) dl_cp_async16(dst + piece * 16, src + piece * 4);
  }
}

// Issue the copies of one stage = quarter q of a chunk whose (row, col) ids are held one per lane:
// the DL_HS neighbour rows, and the own row of the first DL_OWNQ row runs of the quarter.
template <class M>
__device__ __forceinline__ void dl_issue_stage(unsigned char* st, const float* __restrict__ Z,
                                               long long row_base, int rowreg, int colreg, int q, int lane) {
  constexpr int ROWB = M::D * 4, D = M::D;
#pragma unroll
  for (int e = 0; e < DL_HS; ++e) {
    const int r = __shfl_sync(DL_FULL, rowreg, q * DL_HS + e);
    const int c = __shfl_sync(DL_FULL, colreg, q * DL_HS + e);
    if (r >= 0) dl_stage_row<ROWB>(st + e * ROWB, Z + (long long)c * D, lane);
  }
  const int prev = __shfl_up_sync(DL_FULL, rowreg, 1);
  const bool start = (lane / DL_HS) == q && rowreg >= 0 && ((lane % DL_HS) == 0 || prev != rowreg);
  unsigned smask = __ballot_sync(DL_FULL, start);
#pragma unroll
  for (int o = 0; o < DL_OWNQ; ++o) {
    if (smask) {                                   // warp-uniform
      const int pos = __ffs(smask) - 1;
      smask &= smask - 1;
      const int r = __shfl_sync(DL_FULL, rowreg, pos);
      dl_stage_row<ROWB>(st + (DL_HS + o) * ROWB, Z + (row_base + r) * D, lane);
    }
  }
}

template <class M>
__global__ void __launch_bounds__(AttnStreamCfg<M>::THREADS, 1)
k_attn_stream(DlGraphDev g, const int* __restrict__ erow, const float* __restrict__ Z, float T,
              unsigned char* __restrict__ kstar, float* __restrict__ w) {
  using C = AttnStreamCfg<M>;
  constexpr int K = M::K, D = M::D, NP = M::NP, EB = M::EB, LP = M::LP, FPP = M::FPP;
  constexpr int ROWB = C::ROWB, STAGE_B = C::STAGE_B;
  constexpr int SUBS = DL_HS / EB;   // sub-blocks per stage
  constexpr bool DENSE = (M::L == M::LP) && (M::K % M::FPP == 0);   // every lane active in every pass
  extern __shared__ __align__(128) unsigned char dl_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = dl_smem_raw + (size_t)warp * AT_RING * STAGE_B;
  const long long gw = (long long)blockIdx.x * C::NW + warp;

  int off[NP];
  bool act[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { off[p] = M::offset(lane, p); act[p] = M::active(lane, p); }
  const int my_e = M::edge_of_lane(lane);
  const int gsrc = lane & (EB - 1);
  const bool unit_T = (T == 1.0f);
  const unsigned lane_le = 0xffffffffu >> (31 - lane);

  DlChunkStream cs;
  cs.init(g.nnz, (long long)gridDim.x * C::NW, g.range_shift);
  auto load_meta = [&](long long cc, int& r, int& cl) {
    r = -1; cl = 0;
    if (cc >= 0) {
      const long long e = cc * DL_CH + lane;
      if (e < g.nnz) { r = __ldg(erow + e); cl = __ldg(g.col + e); }
    }
  };
  long long c = cs.first(gw);
  int rowA, colA, rowB, colB;
  load_meta(c, rowA, colA);
  long long cn = cs.next(c);
  load_meta(cn, rowB, colB);
#pragma unroll
  for (int pq = 0; pq < AT_RING - 1; ++pq) {
    dl_issue_stage<M>(ring + pq * STAGE_B, Z, g.row_base, rowA, colA, pq, lane);
    dl_cp_async_commit();
  }
  int slot = 0;   // ring slot of the stage about to be consumed

  while (c >= 0) {
    const long long cbase = c * DL_CH;
    int out_ks = 0;
    float out_w = 0.0f;
    // run index (own-row slot) of every entry inside its quarter
    const int prevA = __shfl_up_sync(DL_FULL, rowA, 1);
    const bool startA = rowA >= 0 && ((lane % DL_HS) == 0 || prevA != rowA);
    const unsigned smaskA = __ballot_sync(DL_FULL, startA);
    const unsigned qbits = ((1u << DL_HS) - 1u) << ((lane / DL_HS) * DL_HS);
    const int rankA = __popc(smaskA & qbits & lane_le) - 1;
    const unsigned vmaskA = __ballot_sync(DL_FULL, rowA >= 0);
    const bool allownA = __all_sync(DL_FULL, rankA < DL_OWNQ);
#pragma unroll
    for (int q = 0; q < DL_QPC; ++q) {
      // keep AT_RING-1 stages in flight: issue the stage that far ahead of the one consumed now
      int islot = slot + (AT_RING - 1);
      if (islot >= AT_RING) islot -= AT_RING;
      if (q < DL_QPC - (AT_RING - 1))
        dl_issue_stage<M>(ring + islot * STAGE_B, Z, g.row_base, rowA, colA, q + (AT_RING - 1), lane);
      else
        dl_issue_stage<M>(ring + islot * STAGE_B, Z, g.row_base, rowB, colB, q + (AT_RING - 1) - DL_QPC, lane);
      dl_cp_async_commit();
      dl_cp_async_wait<AT_RING - 1>();
      __syncwarp();
      const unsigned char* st = ring + slot * STAGE_B;
      const int cnt = __popc((vmaskA >> (q * DL_HS)) & ((1u << DL_HS) - 1u));
      for (int sb = 0; sb * EB < cnt; ++sb) {
        float ev[NP];
        // fast path: a full sub-block whose own rows are all staged (the common case) needs no
        // per-entry validity handling; DENSE shapes (every lane owns a chunk) need no lane predicate
        const bool full = allownA && (cnt - sb * EB >= EB);
        if (full) {
          int rk[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) rk[e] = __shfl_sync(DL_FULL, rankA, q * DL_HS + sb * EB + e);
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            float part[EB];
#pragma unroll
            for (int e = 0; e < EB; ++e) {
              float4 zi = dl_zero4(), zj = dl_zero4();
              if (DENSE || act[p]) {
                zj = dl_lds4(st + (sb * EB + e) * ROWB + off[p] * 4);
                zi = dl_lds4(st + (DL_HS + rk[e]) * ROWB + off[p] * 4);
              }
              part[e] = dl_chunk_dot(zi, zj);
            }
            float qv = dl_reduce_scatter<M>(part, lane);
            if (!unit_T) qv = __fdiv_rn(qv, T);
            ev[p] = dl_expf(qv);
          }
        } else {
          int rk[EB], rw[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            rk[e] = __shfl_sync(DL_FULL, rankA, q * DL_HS + sb * EB + e);
            rw[e] = __shfl_sync(DL_FULL, rowA, q * DL_HS + sb * EB + e);
          }
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            float part[EB];
#pragma unroll
            for (int e = 0; e < EB; ++e) {
              const int se = sb * EB + e;
              float4 zi = dl_zero4(), zj = dl_zero4();
              if (se < cnt && act[p]) {
                zj = dl_lds4(st + se * ROWB + off[p] * 4);
                if (rk[e] < DL_OWNQ) zi = dl_lds4(st + (DL_HS + rk[e]) * ROWB + off[p] * 4);
                else zi = dl_ldg4(Z + (g.row_base + rw[e]) * D + off[p]);
              }
              part[e] = dl_chunk_dot(zi, zj);
            }
            float qv = dl_reduce_scatter<M>(part, lane);
            if (!unit_T) qv = __fdiv_rn(qv, T);
            ev[p] = dl_expf(qv);
          }
        }
        float a[K];
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          a[k] = __shfl_sync(DL_FULL, ev[k / FPP], (k % FPP) * LP + gsrc);
          sum = (k == 0) ? a[0] : __fadd_rn(sum, a[k]);
        }
        // first argmax of a_k = e_k / sum; see k_edge_attn_fwd for why one division suffices
        // when the largest exponential is separated by more than 2^-22 relative
        int ks = 0;
        float emax = a[0];
#pragma unroll
        for (int k = 1; k < K; ++k)
          if (a[k] > emax) { emax = a[k]; ks = k; }
        const float thr = __fmul_rn(emax, 0.99999976158142089844f);
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
        if (q * SUBS + sb == lane / EB) { out_ks = ks; out_w = wv; }
      }
      __syncwarp();   // every lane is done with this stage before it is refilled
      slot = (slot + 1 == AT_RING) ? 0 : slot + 1;
    }
    {
      const int oi = (lane & ~(EB - 1)) + my_e;
      if (cbase + oi < g.nnz) {
        kstar[cbase + oi] = (unsigned char)out_ks;
        w[cbase + oi] = out_w;
      }
    }
    c = cn;
    rowA = rowB;
    colA = colB;
    cn = cs.next(c);
    load_meta(cn, rowB, colB);
  }
  dl_cp_async_wait<0>();
}

// s[row,k] = sum of w over the row's entries routed to k (zeros -> 1).  [ref: model.py:70-72]
// One warp per work item (row or hub segment), lanes stride over the entries, one pass per factor.
__global__ void __launch_bounds__(DL_CTA)
k_row_sums(DlGraphDev g, const unsigned char* __restrict__ kstar, const float* __restrict__ w, int K,
           float* __restrict__ s, float* __restrict__ hub_ws) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long n_items = dl_num_items(g);
  for (long long t = warp0; t < n_items; t += nwarps) {
    const DlItem it = dl_decode_item(g, t);
    float mine = 0.0f;   // lane k keeps the sum of factor k
    for (int k = 0; k < K; ++k) {
      float acc = 0.0f;
      for (long long e = it.e0 + lane; e < it.e1; e += 32)
        if (__ldg(kstar + e) == k) acc = __fadd_rn(acc, __ldg(w + e));
      for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(DL_FULL, acc, o));
      if (lane == k) mine = acc;
    }
    if (lane < K) {
      if (it.hub_slot >= 0) hub_ws[it.hub_slot * K + lane] = mine;
      else s[it.node * K + lane] = (mine == 0.0f) ? 1.0f : mine;
    }
  }
}

template <class M>
int launch_attn_stream(const DlGraphDev& g, const int* erow, const float* Z, float T, uint8_t* kstar,
                       float* w, cudaStream_t st) {
  using C = AttnStreamCfg<M>;
  if (!C::OK) return -1000;
  int dev = 0, sms = 0;
  DL_CUDA_TRY(cudaGetDevice(&dev));
  DL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DL_CUDA_TRY(cudaFuncSetAttribute(k_attn_stream<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  const long long n_chunks = (g.nnz + DL_CH - 1) / DL_CH;
  const long long n_ranges = (n_chunks + DL_RANGE - 1) / DL_RANGE;
  long long grid = (n_ranges + C::NW - 1) / C::NW;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  k_attn_stream<M><<<(int)grid, C::THREADS, C::SMEM, st>>>(g, erow, Z, T, kstar, w);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

}  // namespace

// returns -1000 when the shape has no streaming instantiation (caller falls back)
int dl_launch_attn_stream(const DlGraphDev& g, const int* erow, const float* Z, int K, int d, float T,
                          unsigned char* kstar, float* w, float* s, float* hub_ws, cudaStream_t st) {
  int rc = -1000;
#define BODY_MACRO(M) rc = launch_attn_stream<M>(g, erow, Z, T, kstar, w, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  if (rc != DL_OK) return rc;
  rc = dl_launch_gather_stream(2, g, nullptr, nullptr, kstar, w, nullptr, K, d, 0.0f, 0.0f, s, nullptr, hub_ws,
                               st);
  if (rc == -1000) {
    const long long n_items = g.n_hub_items + (g.N - g.n_hub);
    int grid = 1;
    rc = dl_grid_for(k_row_sums, n_items, &grid);
    if (rc) return rc;
    k_row_sums<<<grid, DL_CTA, 0, st>>>(g, kstar, w, K, s, hub_ws);
    DL_LAUNCH_CHECK();
    if (g.n_hub > 0) return -1001;   // caller runs the hub fix-up
    rc = DL_OK;
  }
  if (rc) return rc;
  return DL_OK;
}
