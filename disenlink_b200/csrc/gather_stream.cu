// gather_stream.cu -- streaming per-factor gather / segment-sum with per-row register accumulators.
//
//   MODE 0  aggregation forward  [ref: model.py:75]
//           H[i,k] = beta Z[i,k] + (1-beta) sum_{e in row i, kstar=k} (w[e] / s[col_e,k]) Z[col_e,k]
//   MODE 1  backward pass 1      [ref: autograd of model.py:70-75]
//           T_[i,k] = (1-beta)/s[i,k] sum_{e in row i, kstar=k} w[e] G[col_e,k]
//           r[i,k] = <Z[i,k],T_[i,k]>/s[i,k] ;  dZ[i] += beta G[i] + T_[i]
//   MODE 2  routed row sums      [ref: model.py:70-72]
//           s[i,k] = sum_{e in row i, kstar=k} w[e], zeros -> 1
//
// Same decomposition as attn_stream.cu: the CSR entries are cut into 32-entry chunks regardless of
// row boundaries and every warp walks its own stream of chunks (ranges of 64 chunks), so the load
// is balanced whatever the degree distribution.  Per chunk the warp
//   - has the (row, col, kstar, w) of the chunk two ahead in flight (coalesced loads),
//   - has the s[col,kstar] gather and the 32 routed slices (d floats each, cp.async into a private
//     double-buffered shared-memory tile, 8 entries per instruction at d = 16) of the next chunk
//     in flight,
//   - walks the entries of the current chunk IN CSR ORDER, adding coef * slice into the accumulator
//     of the current row (lane (k, g) owns chunk g of factor k, so only the routed factor's lanes
//     work) and flushing it when the row id changes.
// The accumulation order of a row is exactly its column order -- the same as the CPU oracle.
// A row cut by a range boundary leaves its partial sums in a carry buffer (head / tail per range);
// k_gather_chain sums the pieces of such rows in range order and applies the row epilogue, and
// k_gather_empty_rows writes the rows that have no entries.  No atomics anywhere.
// the gathers of this file are 64-byte slices and 4-byte scalars: ask L2 not to fetch beyond 64 B
// (measured: .L2::64B -2 %, .L2::256B +5 % against no qualifier)
#ifndef DL_CPA_L2
#define DL_CPA_L2 ".L2::64B"
#endif
#ifndef DL_LDG_L2
#define DL_LDG_L2 ".L2::64B"
#endif
#include "dl_dispatch.cuh"
#include "dl_stream.cuh"
#include "dl_fl.cuh"

namespace {

constexpr int GS_WARPS = 16;   // warps per CTA

// the 4-byte s[col,k] gather, with the L2 prefetch-size qualifier of this file
__device__ __forceinline__ float gs_ldg_s(const float* p) {
  float v;
  asm volatile("ld.global.nc" DL_LDG_L2 ".f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// flat width of a row's accumulator / carry record
__host__ __device__ constexpr int carry_width(int mode, int K, int d) {
  return mode == 2 ? K : (mode == 4 ? 2 * K * d : K * d);
}

// ---- row epilogues on a flat accumulator held by one thread-strided warp (rare rows) ----------
__device__ __forceinline__ void epilogue_flat(const DlGraphDev& g, int mode, int lane, long long node, int K, int d,
                                              const float* acc /* global or shared, flat */,
                                              const float* __restrict__ Z, const float* __restrict__ G,
                                              const float* __restrict__ s, float beta, float omb,
                                              float* __restrict__ OUT, float* __restrict__ r) {
  const long long D = (long long)K * d;
  if (mode == 4) {                       // decoder backward: record = [dZ row | dH row], overwrite
    for (long long x = lane; x < D; x += 32) {
      const float dh = acc ? acc[D + x] : 0.0f;
      OUT[node * D + x] = acc ? acc[x] : 0.0f;
      r[node * D + x] = dh;
      for (int q = 0; q < g.n_peer_out; ++q) g.peer_out[q][node * D + x] = dh;      // dH goes to the peers too
    }
  } else if (mode == 3) {                // plain accumulate (backward pass 2 partial sums)
    if (acc)
      for (long long x = lane; x < D; x += 32) OUT[node * D + x] = __fadd_rn(OUT[node * D + x], acc[x]);
  } else if (mode == 2) {
    for (int k = lane; k < K; k += 32) {
      const float v = acc ? acc[k] : 0.0f;
      const float sv = (v == 0.0f) ? 1.0f : v;
      OUT[node * K + k] = sv;
      for (int q = 0; q < g.n_peer_out; ++q) g.peer_out[q][node * K + k] = sv;     // s goes to the peers too
    }
  } else if (mode == 0) {
    for (long long x = lane; x < D; x += 32) {
      const float v = acc ? acc[x] : 0.0f;
      const float h = __fadd_rn(__fmul_rn(beta, Z[node * D + x]), __fmul_rn(omb, v));
      OUT[node * D + x] = h;
      for (int q = 0; q < g.n_peer_out; ++q) g.peer_out[q][node * D + x] = h;       // H goes to the peers too
    }
  } else {
    for (int k = 0; k < K; ++k) {
      const float sk = s[node * K + k];
      const float scale = __fdiv_rn(omb, sk);
      float part = 0.0f;
      for (int x = lane; x < d; x += 32) {
        const long long o = node * D + (long long)k * d + x;
        const float tv = __fmul_rn(scale, acc ? acc[(long long)k * d + x] : 0.0f);
        part = __fmaf_rn(Z[o], tv, part);
        OUT[o] = __fadd_rn(OUT[o], __fmaf_rn(beta, G[o], tv));
      }
      for (int o = 16; o > 0; o >>= 1) part = __fadd_rn(part, __shfl_xor_sync(DL_FULL, part, o));
      if (lane == 0) {
        const float rv = __fdiv_rn(part, sk);
        r[node * K + k] = rv;
        for (int q = 0; q < g.n_peer_out; ++q) g.peer_out[q][node * K + k] = rv;   // r goes to the peers too
      }
    }
  }
}

// rows without entries never show up in the entry stream
__global__ void __launch_bounds__(DL_CTA)
k_gather_empty_rows(DlGraphDev g, int mode, int K, int d, const float* __restrict__ Z,
                    const float* __restrict__ G, const float* __restrict__ s, float beta, float omb,
                    float* __restrict__ OUT, float* __restrict__ r) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  // a warp scans 32 rows at a time (coalesced rowptr reads) and handles the empty ones
  for (long long r0 = warp0 * 32; r0 < g.N; r0 += nwarps * 32) {
    const long long row = r0 + lane;
    const bool empty = row < g.N && __ldg(g.rowptr + row) == __ldg(g.rowptr + row + 1);
    unsigned m = __ballot_sync(DL_FULL, empty);
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      epilogue_flat(g, mode, lane, g.row_base + r0 + l, K, d, nullptr, Z, G, s, beta, omb, OUT, r);
    }
  }
}

// rows cut by range boundaries: sum the pieces in range order, then the epilogue.
// carry layout: [range][0 = head, 1 = tail][W]
__global__ void __launch_bounds__(DL_CTA)
k_gather_chain(DlGraphDev g, int mode, int K, int d, const float* __restrict__ carry,
               const float* __restrict__ Z, const float* __restrict__ G, const float* __restrict__ s,
               float beta, float omb, float* __restrict__ OUT, float* __restrict__ r,
               float* __restrict__ scratch /* [n_ranges][W] */) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const int W = carry_width(mode, K, d);
  const long long RE = (long long)DL_CH * DL_RANGE;            // entries per range
  const long long n_ranges = (g.nnz + RE - 1) / RE;
  for (long long b = warp0; b < n_ranges; b += nwarps) {
    const long long R1 = min((b + 1) * RE, g.nnz);
    if (R1 >= g.nnz) continue;                                  // nothing after the last range
    const int row = __ldg(g.erow + R1 - 1);
    if (__ldg(g.erow + R1) != row) continue;                    // no row crosses this boundary
    if (__ldg(g.rowptr + row) < b * RE) continue;               // the chain started in an earlier range
    const long long last = (__ldg(g.rowptr + row + 1) - 1) / RE;   // range holding the row's last entry
    float* acc = scratch + b * W;
    // every element is summed in range order (the same bits as a plain loop); eight elements per lane and four
    // ranges per step are in flight, because a hub row of the primary view chains hundreds of ranges and one
    // dependent load per step made this fix-up 4 % of the step
    for (int x0 = 0; x0 < W; x0 += 256) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int x = x0 + j * 32 + lane;
        v[j] = x < W ? carry[(b * 2 + 1) * W + x] : 0.0f;        // tail of the first range
      }
      long long bb = b + 1;
      for (; bb + 3 <= last; bb += 4) {
        float t[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int x = x0 + j * 32 + lane;
            t[u][j] = x < W ? carry[((bb + u) * 2) * W + x] : 0.0f;
          }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __fadd_rn(v[j], t[u][j]);
      }
      for (; bb <= last; ++bb)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int x = x0 + j * 32 + lane;
          if (x < W) v[j] = __fadd_rn(v[j], carry[(bb * 2) * W + x]);
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int x = x0 + j * 32 + lane;
        if (x < W) acc[x] = v[j];
      }
    }
    __syncwarp();
    epilogue_flat(g, mode, lane, g.row_base + row, K, d, acc, Z, G, s, beta, omb, OUT, r);
  }
}

// ---- the streaming kernel -----------------------------------------------------------------
template <class M, int MODE>
struct GatherStreamCfg {
  static constexpr int SLB = M::d * 4;                                  // bytes of one routed slice
  static constexpr int TILE_B = (MODE == 2) ? 0 : DL_CH * SLB;          // one chunk of slices
  static constexpr int CF_B = (MODE == 2) ? 0 : DL_CH * 4;              // the chunk's coefficients
  static constexpr int X_B = (MODE == 1) ? DL_CH * 4 : 0;               // the chunk's <G[j,k*], Z[i,k*]> dots
  static constexpr int WARP_B = 2 * TILE_B + CF_B + X_B;
  static constexpr size_t SMEM = (size_t)GS_WARPS * WARP_B;
};

// two CTAs per SM (32 warps): the register allocation must stay at 64 -- without the bound the pass-1
// instantiation drifted to 106 registers when two kernel parameters were added, i.e. to one CTA per SM and
// +40 % time (the kernel lives on the number of gathers in flight).  Shapes that need two passes over the
// factors (NP > 1: wide d) hold twice the accumulators and are left to the compiler.
template <class M, int MODE>
__global__ void __launch_bounds__(GS_WARPS * 32, (M::NP == 1) ? 2 : 1)
k_gather_stream(DlGraphDev g, const float* __restrict__ Z, const float* __restrict__ SRC,
                const unsigned char* __restrict__ kstar, const float* __restrict__ w,
                const float* __restrict__ s, float beta, float omb, float* __restrict__ OUT,
                float* __restrict__ r, float* __restrict__ carry, float* __restrict__ xout,
                const int* __restrict__ xidx, unsigned char* __restrict__ ku_out) {
  using C = GatherStreamCfg<M, MODE>;
  constexpr int K = M::K, d = M::d, D = M::D, NP = M::NP, L = M::L, LP = M::LP, FPP = M::FPP;
  constexpr int NG = 32 / LP, SLB = C::SLB, TILE_B = C::TILE_B;
  constexpr int W = (MODE == 2) ? K : D;
  extern __shared__ __align__(128) unsigned char dl_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* tile = dl_smem_raw + (size_t)warp * C::WARP_B;
  float* cfbuf = reinterpret_cast<float*>(tile + 2 * TILE_B);
  float* xbuf = cfbuf + DL_CH;                      // MODE 1 only
  const long long gw = (long long)blockIdx.x * GS_WARPS + warp;
  const int grp = lane / LP, gg = lane % LP, slot = M::slot(lane);
  const bool glane = gg < L;
  const long long RE = (long long)DL_CH * DL_RANGE;
  // lanes of one factor group (they walk the same entries, so they may shuffle among themselves
  // while the other groups are elsewhere in the divergent walk)
  const unsigned gmask = (LP >= 32) ? 0xffffffffu : (((1u << LP) - 1u) << (grp * LP));
  const bool want_x = (MODE == 1) && (L == LP) && xout != nullptr;

  DlChunkStream cs;
  cs.init(g.nnz, (long long)gridDim.x * GS_WARPS, g.range_shift);

  struct Meta { int row, col, ks; float wv; };
  auto load_meta = [&](long long cc, Meta& m) {
    m.row = -1; m.col = 0; m.ks = 255; m.wv = 0.0f;
    if (cc >= 0) {
      const long long e = cc * DL_CH + lane;
      if (e < g.nnz) {
        m.row = __ldg(g.erow + e);
        m.ks = __ldg(kstar + e);
        m.wv = __ldg(w + e);
        if (MODE != 2) m.col = __ldg(g.col + e);
      }
    }
  };
  auto issue_slices = [&](unsigned char* buf, const Meta& m) {
    if (MODE == 2) return;
#pragma unroll
    for (int rd = 0; rd < LP; ++rd) {
      const int idx = rd * NG + grp;
      const int cc = __shfl_sync(DL_FULL, m.col, idx);
      const int kk = __shfl_sync(DL_FULL, m.ks, idx);
      if (glane && kk < K)
        dl_cp_async16(buf + idx * SLB + gg * 16, SRC + (long long)cc * D + kk * d + gg * 4);
    }
  };

  // accumulator of the current row: lane (slot, g) owns chunk g of factor p*FPP+slot (MODE 0/1);
  // lane k owns factor k (MODE 2)
  float4 acc[NP];
  float acc2 = 0.0f;
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = dl_zero4();
  int cur_row = -1;
  // row vectors the epilogue needs, loaded when a run STARTS so their latency overlaps the run
  float4 zpre[NP], gpre[NP];
  float skpre[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { zpre[p] = gpre[p] = dl_zero4(); skpre[p] = 1.0f; }
  auto prefetch_row = [&](int row) {
    if (MODE == 2) return;
    const long long node = g.row_base + row;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (!M::active(lane, p)) continue;
      const int o = M::offset(lane, p);
      zpre[p] = dl_ldg4(Z + node * D + o);
      if (MODE == 1) gpre[p] = dl_ldg4(SRC + node * D + o);
    }
    if (MODE == 1) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int k = M::factor(lane, p);
        skpre[p] = (k < K) ? __ldg(s + node * K + k) : 1.0f;
      }
    }
  };
  bool first_run = true;          // no flush yet in the current range
  bool head_open = false, tail_open = false;
  long long cur_range = -1;

  auto flush = [&](bool at_range_end) {
    if (cur_row >= 0) {
      const bool to_head = first_run && head_open;
      const bool to_tail = !to_head && at_range_end && tail_open;
      if (to_head || to_tail) {
        float* dst = carry + (cur_range * 2 + (to_tail ? 1 : 0)) * W;
        if (MODE == 2) {
          if (lane < K) dst[lane] = acc2;
        } else {
#pragma unroll
          for (int p = 0; p < NP; ++p)
            if (M::active(lane, p)) *reinterpret_cast<float4*>(dst + M::offset(lane, p)) = acc[p];
        }
      } else {
        const long long node = g.row_base + cur_row;
        if (MODE == 2) {
          if (lane < K) {
            const float sv = (acc2 == 0.0f) ? 1.0f : acc2;
            OUT[node * K + lane] = sv;
            for (int q = 0; q < g.n_peer_out; ++q) g.peer_out[q][node * K + lane] = sv;
          }
        } else if (MODE == 0) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            if (!M::active(lane, p)) continue;
            const int o = M::offset(lane, p);
            const float4 zi = zpre[p];
            float4 h;
            h.x = __fadd_rn(__fmul_rn(beta, zi.x), __fmul_rn(omb, acc[p].x));
            h.y = __fadd_rn(__fmul_rn(beta, zi.y), __fmul_rn(omb, acc[p].y));
            h.z = __fadd_rn(__fmul_rn(beta, zi.z), __fmul_rn(omb, acc[p].z));
            h.w = __fadd_rn(__fmul_rn(beta, zi.w), __fmul_rn(omb, acc[p].w));
            *reinterpret_cast<float4*>(OUT + node * D + o) = h;
#pragma unroll 1
            for (int q = 0; q < g.n_peer_out; ++q)            // the all-gather of H rides on the kernel
              *reinterpret_cast<float4*>(g.peer_out[q] + node * D + o) = h;
          }
        } else {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const int k = M::factor(lane, p);
            const bool act = M::active(lane, p);
            const int o = M::offset(lane, p);
            const float sk = skpre[p];
            const float scale = __fdiv_rn(omb, sk);
            float4 tv;
            tv.x = __fmul_rn(scale, acc[p].x); tv.y = __fmul_rn(scale, acc[p].y);
            tv.z = __fmul_rn(scale, acc[p].z); tv.w = __fmul_rn(scale, acc[p].w);
            const float4 zi = act ? zpre[p] : dl_zero4();
            const float dotzt = dl_group_sum<M>(dl_chunk_dot(zi, tv));
            if (k < K && gg == 0) {
              const float rv = __fdiv_rn(dotzt, sk);
              r[node * K + k] = rv;
              for (int q = 0; q < g.n_peer_out; ++q) g.peer_out[q][node * K + k] = rv;
            }
            if (act) {
              // dZ[i] += beta G[i] + T_[i]: this warp is the row's only direct writer in this launch (rows cut by
              // a range boundary go through the carries), so a fire-and-forget reduction gives the value of
              // load-add-store in any schedule -- without holding the old dZ chunk in registers for the whole run
              const float4 gi = gpre[p];
              float4 add;
              add.x = __fmaf_rn(beta, gi.x, tv.x); add.y = __fmaf_rn(beta, gi.y, tv.y);
              add.z = __fmaf_rn(beta, gi.z, tv.z); add.w = __fmaf_rn(beta, gi.w, tv.w);
              fl_red_add4(OUT + node * D + o, add);
            }
          }
        }
      }
    }
    if (cur_row >= 0) first_run = false;     // only a real run consumes the "first run of the range" slot
    cur_row = -1;
    acc2 = 0.0f;
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] = dl_zero4();
  };

  long long c = cs.first(gw);
  Meta mA, mB, mC;
  load_meta(c, mA);
  long long cn = cs.next(c);
  load_meta(cn, mB);
  float sjA = 1.0f, sjB = 1.0f;
  // s == nullptr in MODE 0: SRC holds slices already divided by s (factor_fwd.cu k_scale_rows)
  if (MODE == 0 && s != nullptr && mA.ks != 255) sjA = gs_ldg_s(s + (long long)mA.col * K + mA.ks);
  int buf = 0;
  issue_slices(tile, mA);
  dl_cp_async_commit();

  while (c >= 0) {
    // pipeline: metadata two chunks ahead, s gather + slices one chunk ahead
    const long long cnn = cs.next(cn);
    load_meta(cnn, mC);
    if (MODE == 0 && s != nullptr && mB.ks != 255) sjB = gs_ldg_s(s + (long long)mB.col * K + mB.ks);
    issue_slices(tile + (buf ^ 1) * TILE_B, mB);
    dl_cp_async_commit();
    dl_cp_async_wait<1>();
    __syncwarp();

    // range bookkeeping
    const long long rg = c >> g.range_shift;
    if (rg != cur_range) {
      if (cur_range >= 0) flush(true);
      cur_range = rg;
      first_run = true;
      const long long R0 = rg * RE, R1 = min(R0 + RE, g.nnz);
      head_open = R0 > 0 && __ldg(g.erow + R0 - 1) == __ldg(g.erow + R0);
      tail_open = R1 < g.nnz && __ldg(g.erow + R1) == __ldg(g.erow + R1 - 1);
    }

    // MODE 0 has no r output: the pointer carries the optional per-entry copy of s[col, kstar]
    // (sj_out of dl_factor_spmm_fwd) that lets backward pass 2 skip this gather
    if (MODE == 0 && r != nullptr && s != nullptr && mA.row >= 0) r[c * DL_CH + lane] = sjA;
    const float coefA = (MODE == 0) ? __fdiv_rn(mA.wv, sjA) : mA.wv;
    const unsigned vmask = __ballot_sync(DL_FULL, mA.row >= 0);
    const int cnt = __popc(vmask);
    const unsigned char* sl = tile + buf * TILE_B;
    // the chunk is walked run by run (a run = consecutive entries of one row), so the hot loop has
    // no row test: bit i of `starts` is set when entry i opens a new run
    const int prow = __shfl_up_sync(DL_FULL, mA.row, 1);
    const bool st = mA.row >= 0 && (lane == 0 ? mA.row != cur_row : mA.row != prow);
    const unsigned starts = __ballot_sync(DL_FULL, st);
    int idx = 0;
    if (MODE == 2) {
      while (idx < cnt) {
        if ((starts >> idx) & 1u) {
          flush(false);
          cur_row = __shfl_sync(DL_FULL, mA.row, idx);
        }
        const unsigned rest = (idx < 31) ? (starts & ~((2u << idx) - 1u)) : 0u;
        const int end = rest ? (__ffs(rest) - 1) : cnt;
        for (; idx < end; ++idx) {
          const unsigned ke = (unsigned)__shfl_sync(DL_FULL, mA.ks, idx);
          const float cf = __shfl_sync(DL_FULL, coefA, idx);
          if ((unsigned)lane == ke) acc2 = __fadd_rn(acc2, cf);
        }
      }
    } else {
      // Every lane group walks only the entries routed to ITS factor: mine[p] has bit i set when
      // entry i of the chunk belongs to factor p*FPP + slot.  The walk is divergent between lane
      // groups (no shuffles inside: coefficients and slices come from shared memory) and costs
      // max-over-factors entries per run instead of all of them; each accumulator still receives
      // its entries in CSR order, so the sums are bit-identical to an entry-by-entry walk.
      cfbuf[lane] = coefA;
      unsigned mine[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) mine[p] = 0u;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const unsigned b = __ballot_sync(DL_FULL, mA.ks == k);
        if (glane && slot == (k % FPP)) mine[k / FPP] = b;
      }
      __syncwarp();
      while (idx < cnt) {
        if ((starts >> idx) & 1u) {
          flush(false);
          cur_row = __shfl_sync(DL_FULL, mA.row, idx);
          prefetch_row(cur_row);
        }
        const unsigned rest = (idx < 31) ? (starts & ~((2u << idx) - 1u)) : 0u;
        const int end = rest ? (__ffs(rest) - 1) : cnt;
        const unsigned runbits = ((end < 32) ? ((1u << end) - 1u) : 0xffffffffu) & ~((1u << idx) - 1u);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          unsigned m = mine[p] & runbits;
          while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            const float cf = cfbuf[i];
            const float4 v = dl_lds4(sl + i * SLB + gg * 16);
            dl_fma4(acc[p], cf, v);
            if (MODE == 1 && want_x) {
              // x[e] = <G[j,k*], Z[i,k*]> in the canonical order (chunk chains + balanced tree): the
              // row's own Z chunk is in zpre (loaded when the run started), the routed slice is v
              float pd = dl_chunk_dot(v, zpre[p]);
#pragma unroll
              for (int off = 1; off < LP; off <<= 1) pd = __fadd_rn(pd, __shfl_xor_sync(gmask, pd, off));
              if (gg == 0) xbuf[i] = pd;
            }
          }
        }
        idx = end;
      }
    }
    __syncwarp();
    if (MODE == 1 && want_x && mA.row >= 0) {
      // xidx (symmetric backward pass 2): only the primary entries (eidx >= 0) are needed, in primary-view order
      if (xidx == nullptr) {
        xout[c * DL_CH + lane] = xbuf[lane];
      } else {
        const int t = __ldg(xidx + c * DL_CH + lane);
        if (t >= 0) {
          xout[t] = xbuf[lane];
          if (ku_out) ku_out[t] = (unsigned char)mA.ks;     // kstar in primary-view order, for phase A
        }
      }
    }
    buf ^= 1;
    c = cn; cn = cnn;
    mA = mB; mB = mC;
    sjA = sjB;
  }
  if (cur_range >= 0) flush(true);
  dl_cp_async_wait<0>();
}

template <class M, int MODE>
int launch_gather_stream(const DlGraphDev& g, const float* Z, const float* SRC,
                         const unsigned char* kstar, const float* w, const float* s, float beta,
                         float omb, float* OUT, float* r, float* carry, float* xout, const int* xidx,
                         unsigned char* ku_out, cudaStream_t st) {
  using C = GatherStreamCfg<M, MODE>;
  if (C::SMEM > 200 * 1024) return -1000;
  if (C::SMEM > 48 * 1024)
    DL_CUDA_TRY(cudaFuncSetAttribute(k_gather_stream<M, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)C::SMEM));
  int dev = 0, sms = 0, per_sm = 0;
  DL_CUDA_TRY(cudaGetDevice(&dev));
  DL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gather_stream<M, MODE>, GS_WARPS * 32,
                                                            C::SMEM));
  if (per_sm < 1) per_sm = 1;
  const long long n_chunks = (g.nnz + DL_CH - 1) / DL_CH;
  const long long n_ranges = (n_chunks + DL_RANGE - 1) / DL_RANGE;
  long long grid = (n_ranges + GS_WARPS - 1) / GS_WARPS;
  const long long cap = (long long)sms * per_sm;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_gather_stream<M, MODE><<<(int)grid, GS_WARPS * 32, C::SMEM, st>>>(g, Z, SRC, kstar, w, s, beta, omb, OUT,
                                                                     r, carry, xout, xidx, ku_out);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

inline int small_grid(long long n_warps_wanted) {
  long long b = (n_warps_wanted + DL_WARPS_PER_CTA - 1) / DL_WARPS_PER_CTA;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

}  // namespace

// floats of scratch the streaming gather needs: carries [n_ranges][2][W] + chain scratch [n_ranges][W]
size_t dl_gather_stream_scratch_floats(long long nnz, int mode, int K, int d) {
  const long long RE = (long long)DL_CH << dl_range_shift(nnz);
  const long long n_ranges = (nnz + RE - 1) / RE;
  return (size_t)n_ranges * 3 * (size_t)carry_width(mode, K, d);
}

// mode 0 / 1 / 2 as above; returns -1000 when (K, d) has no streaming instantiation.
// scratch: dl_gather_stream_scratch_floats floats.
int dl_launch_gather_stream(int mode, const DlGraphDev& g, const float* Z, const float* SRC,
                            const unsigned char* kstar, const float* w, const float* s, int K, int d,
                            float beta, float omb, float* OUT, float* r, float* scratch,
                            cudaStream_t st, float* xout, const int* xidx, unsigned char* ku_out) {
  if (!g.erow || g.nnz == 0 || !scratch) return -1000;
  const long long RE = (long long)DL_CH * DL_RANGE;
  const long long n_ranges = (g.nnz + RE - 1) / RE;
  const int W = carry_width(mode, K, d);
  float* carry = scratch;
  float* chain = scratch + (size_t)n_ranges * 2 * W;
  int rc = -1000;
#define BODY_MACRO(M)                                                                                       \
  rc = (mode == 0)   ? launch_gather_stream<M, 0>(g, Z, SRC, kstar, w, s, beta, omb, OUT, r, carry, nullptr, nullptr, nullptr, st)  \
       : (mode == 1) ? launch_gather_stream<M, 1>(g, Z, SRC, kstar, w, s, beta, omb, OUT, r, carry, xout, xidx, ku_out, st)     \
                     : launch_gather_stream<M, 2>(g, Z, SRC, kstar, w, s, beta, omb, OUT, r, carry, nullptr, nullptr, nullptr, st);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  if (rc != DL_OK) return rc;
  float* r_node = (mode == 0) ? nullptr : r;       // in mode 0 `r` is the per-entry sj output
  k_gather_chain<<<small_grid(n_ranges), DL_CTA, 0, st>>>(g, mode, K, d, carry, Z, SRC, s, beta, omb, OUT,
                                                          r_node, chain);
  DL_LAUNCH_CHECK();
  k_gather_empty_rows<<<small_grid((g.N + 31) / 32), DL_CTA, 0, st>>>(g, mode, K, d, Z, SRC, s, beta, omb,
                                                                     OUT, r_node);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

bool dl_gather_stream_has_x(int K, int d) {
  bool has = false;
#define BODY_MACRO(M) has = (M::L == M::LP);
  DL_DISPATCH_SHAPES()
#undef BODY_MACRO
  return has;
}

// chain fix-up alone, adding the stitched partial sums into OUT (used by bwd_stream.cu)
int dl_gather_chain_add(const DlGraphDev& g, int K, int d, float* scratch, float* OUT, cudaStream_t st) {
  const long long RE = (long long)DL_CH * DL_RANGE;
  const long long n_ranges = (g.nnz + RE - 1) / RE;
  const int W = K * d;
  k_gather_chain<<<small_grid(n_ranges), DL_CTA, 0, st>>>(g, 3, K, d, scratch, nullptr, nullptr, nullptr, 0.0f,
                                                          0.0f, OUT, nullptr, scratch + (size_t)n_ranges * 2 * W);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

// chain fix-up + empty nodes for the streaming decoder backward (mode 4: OUT = dZ, second = dH)
int dl_gather_chain_pair(const DlGraphDev& g, int K, int d, float* scratch, float* dZ, float* dH,
                         cudaStream_t st) {
  const long long RE = (long long)DL_CH * DL_RANGE;
  const long long n_ranges = (g.nnz + RE - 1) / RE;
  const int W = 2 * K * d;
  k_gather_chain<<<small_grid(n_ranges), DL_CTA, 0, st>>>(g, 4, K, d, scratch, nullptr, nullptr, nullptr, 0.0f,
                                                          0.0f, dZ, dH, scratch + (size_t)n_ranges * 2 * W);
  DL_LAUNCH_CHECK();
  k_gather_empty_rows<<<small_grid((g.N + 31) / 32), DL_CTA, 0, st>>>(g, 4, K, d, nullptr, nullptr, nullptr,
                                                                     0.0f, 0.0f, dZ, dH);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

// chain fix-up + empty rows for row sums produced by attn_fl.cu (mode 2: W = K, 0 -> 1)
int dl_gather_chain_rowsum(const DlGraphDev& g, int K, float* scratch, float* s_out, cudaStream_t st) {
  const long long RE = (long long)DL_CH * DL_RANGE;
  const long long n_ranges = (g.nnz + RE - 1) / RE;
  k_gather_chain<<<small_grid(n_ranges), DL_CTA, 0, st>>>(g, 2, K, 1, scratch, nullptr, nullptr, nullptr, 0.0f,
                                                          0.0f, s_out, nullptr, scratch + (size_t)n_ranges * 2 * K);
  DL_LAUNCH_CHECK();
  k_gather_empty_rows<<<small_grid((g.N + 31) / 32), DL_CTA, 0, st>>>(g, 2, K, 1, nullptr, nullptr, nullptr,
                                                                     0.0f, 0.0f, s_out, nullptr);
  DL_LAUNCH_CHECK();
  return DL_OK;
}
