// sampling.cu -- structured negative sampling on the device (integer work, seedable).
//
// [ref: main_disentangled.py:160 -- torch_geometric.utils.structured_negative_sampling(edge_index),
// called m times per run on the full edge set.]  PyG's semantics (external dependency, version
// unpinned in the reference): for every edge column (i, j) draw k ~ U[0, num_nodes) and redraw
// while (i, k) is itself an edge column (k == i is allowed unless (i, i) is an edge); return k.
// PyG draws with torch.randint on the CPU, unseeded, so its stream cannot be reproduced; here the
// stream is a counter-based Philox4x32-10 keyed by `seed` with counter (edge id, attempt): every
// edge's draws are independent of launch geometry, reproducible, and restated bit-exactly by the
// numpy oracle.  Membership of (i, k) is a binary search in row i of the directed CSR.
#include "dl_common.cuh"

namespace {

struct Philox {
  unsigned c0, c1, c2, c3;
};

__device__ __forceinline__ Philox philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                                unsigned k0, unsigned k1) {
#pragma unroll
  for (int rd = 0; rd < 10; ++rd) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox p = {c0, c1, c2, c3};
  return p;
}

__device__ __forceinline__ bool row_contains(const DlGraphDev& g, long long row, int k) {
  long long lo = __ldg(g.rowptr + row), hi = __ldg(g.rowptr + row + 1);
  while (lo < hi) {
    const long long mid = lo + (hi - lo) / 2;
    const int c = __ldg(g.col + mid);
    if (c == k) return true;
    if (c < k) lo = mid + 1; else hi = mid;
  }
  return false;
}

__global__ void k_structured_neg(DlGraphDev g, const long long* __restrict__ src, long long E,
                                 long long num_nodes, unsigned long long seed, int max_tries,
                                 long long* __restrict__ k_out, int* __restrict__ n_failed) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const long long i = __ldg(src + e);
    long long k = -1;
    for (int t = 0; t < max_tries; ++t) {
      const Philox p = philox4x32_10((unsigned)e, (unsigned)((unsigned long long)e >> 32), (unsigned)t, 0u,
                                     (unsigned)seed, (unsigned)(seed >> 32));
      const unsigned long long r = ((unsigned long long)p.c1 << 32) | p.c0;
      const long long cand = (long long)__umul64hi(r, (unsigned long long)num_nodes);
      if (!row_contains(g, i, (int)cand)) { k = cand; break; }
    }
    if (k < 0) {
      // every draw hit a neighbour (a near-complete row): first non-neighbour at or after the last
      // candidate position 0, scanning the sorted row
      long long a = __ldg(g.rowptr + i), b = __ldg(g.rowptr + i + 1), want = 0;
      for (; a < b; ++a) {
        const int c = __ldg(g.col + a);
        if (c > want) break;
        if (c == want) ++want;
      }
      if (want < num_nodes) k = want;
      else atomicAdd(n_failed, 1);
    }
    k_out[e] = k;
  }
}

}  // namespace

extern "C" int dl_structured_negative_sampling(const dl_graph* g_host, const int64_t* src, int64_t E,
                                               int64_t num_nodes, uint64_t seed, int max_tries, int64_t* k_out,
                                               int* n_failed, dl_stream_t stream) {
  if (!g_host || E < 0 || num_nodes <= 0 || num_nodes > 0x7fffffffLL || max_tries < 1 || !n_failed) return DL_EINVAL;
  if (E > 0 && (!src || !k_out || !g_host->rowptr)) return DL_EINVAL;
  if (g_host->N < num_nodes) return DL_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  DL_CUDA_TRY(cudaMemsetAsync(n_failed, 0, sizeof(int), st));
  if (E == 0) return DL_OK;
  const DlGraphDev g = dl_graph_dev(g_host);
  long long grid = (E + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  k_structured_neg<<<(int)grid, 256, 0, st>>>(g, (const long long*)src, E, num_nodes, seed, max_tries,
                                              (long long*)k_out, n_failed);
  DL_LAUNCH_CHECK();
  return DL_OK;
}
