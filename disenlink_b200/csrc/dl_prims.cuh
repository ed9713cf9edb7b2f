// dl_prims.cuh -- the device-wide primitives of the integer pipeline, hand-written: a generic
// three-kernel scan (reduce per chunk, scan of the chunk totals, rescan with the carried prefix) and a
// stable least-significant-digit radix sort (8-bit digits: per-tile histograms, one scan of the
// digit-major counter table, stable scatter).  They replace the CUB calls of the first version of
// graph_build.cu / eval_metrics.cu; everything they produce is an exact function of the input (no
// floating point, no order-dependent atomics), so the CSR, bucket, incidence and AUC outputs stay
// bit-identical to the numpy oracle.
#pragma once
#include "dl_common.cuh"

namespace dlp {

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// ---------------------------------------------------------------------------------------------
// scan:  out(i, prefix) for i in [0, n), prefix = op-fold of in(0..i) (inclusive) or in(0..i-1)
// (exclusive, `ident` for i = 0).  In / Out are device functors, so callers fuse a transform on the
// way in (flags from neighbouring keys) and a scatter on the way out (unique, compaction).
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_IPT;

template <class T>
struct OpSum {
  __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
template <class T>
struct OpMax {
  __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; }
};

template <class T, class Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T ident, T* sh /* [SCAN_THREADS / 32] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(DL_FULL, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = ident;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) r = op(r, sh[w]);
  __syncthreads();
  return r;
}

template <class T, class Op, class In>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(In in, long long n, Op op, T ident, T* __restrict__ part) {
  __shared__ T sh[SCAN_THREADS / 32];
  const long long base = (long long)blockIdx.x * SCAN_CHUNK;
  T v = ident;
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) {
    const long long i = base + (long long)threadIdx.x * SCAN_IPT + j;
    if (i < n) v = op(v, in(i));
  }
  const T r = block_reduce(v, op, ident, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// exclusive scan of the chunk totals, in place, by one block (sequential over 1024-wide strips)
template <class T, class Op>
__global__ void __launch_bounds__(1024) k_scan_partials(T* __restrict__ part, long long nb, Op op, T ident) {
  __shared__ T sh[32];
  __shared__ T carry_s;
  if (threadIdx.x == 0) carry_s = ident;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (long long s0 = 0; s0 < nb; s0 += 1024) {
    const long long i = s0 + threadIdx.x;
    const T x = i < nb ? part[i] : ident;
    T v = x;                                           // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const T y = __shfl_up_sync(DL_FULL, v, o);
      if (lane >= o) v = op(y, v);
    }
    if (lane == 31) sh[w] = v;
    __syncthreads();
    T wp = ident;                                      // totals of the warps before mine
    for (int k = 0; k < w; ++k) wp = op(wp, sh[k]);
    const T carry = carry_s;
    const T incl = op(carry, op(wp, v));
    // exclusive value = carry (+) warps before (+) lanes before
    T prev = __shfl_up_sync(DL_FULL, v, 1);
    const T excl = op(carry, lane == 0 ? wp : op(wp, prev));
    if (i < nb) part[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
}

template <class T, class Op, class In, class Out, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(In in, Out out, long long n, Op op, T ident, const T* __restrict__ part) {
  __shared__ T sh[SCAN_THREADS / 32];
  const long long base = (long long)blockIdx.x * SCAN_CHUNK;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T x[SCAN_IPT];
  T tot = ident;
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) {
    const long long i = base + (long long)threadIdx.x * SCAN_IPT + j;
    x[j] = i < n ? in(i) : ident;
    tot = op(tot, x[j]);
  }
  T v = tot;                                            // inclusive scan of the thread totals in the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T y = __shfl_up_sync(DL_FULL, v, o);
    if (lane >= o) v = op(y, v);
  }
  if (lane == 31) sh[w] = v;
  __syncthreads();
  T pre = part[blockIdx.x];
  for (int k = 0; k < w; ++k) pre = op(pre, sh[k]);
  const T prev = __shfl_up_sync(DL_FULL, v, 1);
  if (lane > 0) pre = op(pre, prev);                    // exclusive prefix of this thread's first item
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) {
    const long long i = base + (long long)threadIdx.x * SCAN_IPT + j;
    const T incl = op(pre, x[j]);
    if (i < n) out(i, INCLUSIVE ? incl : pre);
    pre = incl;
  }
}

inline size_t scan_ws_bytes(long long n, size_t elem) {
  const long long nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  return align256((size_t)(nb > 0 ? nb : 1) * elem);
}

template <bool INCLUSIVE, class T, class Op, class In, class Out>
int scan(In in, Out out, long long n, Op op, T ident, void* ws, cudaStream_t st) {
  if (n <= 0) return DL_OK;
  const long long nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  if (nb > 0x7fffffffLL) return DL_EINVAL;
  T* part = (T*)ws;
  k_scan_reduce<T, Op, In><<<(int)nb, SCAN_THREADS, 0, st>>>(in, n, op, ident, part);
  DL_LAUNCH_CHECK();
  k_scan_partials<T, Op><<<1, 1024, 0, st>>>(part, nb, op, ident);
  DL_LAUNCH_CHECK();
  k_scan_apply<T, Op, In, Out, INCLUSIVE><<<(int)nb, SCAN_THREADS, 0, st>>>(in, out, n, op, ident, part);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

// plain array in / out functors
template <class T>
struct LoadArr {
  const T* p;
  __device__ __forceinline__ T operator()(long long i) const { return p[i]; }
};
template <class T>
struct StoreArr {
  T* p;
  __device__ __forceinline__ void operator()(long long i, T v) const { p[i] = v; }
};

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits.  (kin, vin) are clobbered; the result is in (kout, vout).
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_IPT = 32;
constexpr int RS_TILE = RS_THREADS * RS_IPT;

struct NoVal {};

template <class K>
__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const K* __restrict__ kin, long long n, int shift, unsigned* __restrict__ hist, int nb) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_IPT; ++r) {
    const long long i = base + (long long)r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(unsigned)(kin[i] >> shift) & 255u], 1u);   // integer counts: order-independent
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];        // digit-major
}

template <class K, class V, bool HASV>
__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const K* __restrict__ kin, K* __restrict__ kout, const V* __restrict__ vin, V* __restrict__ vout,
             long long n, int shift, const unsigned* __restrict__ offs, int nb) {
  __shared__ unsigned base[256];                  // next output position of every digit for this tile
  __shared__ unsigned short wcnt[RS_THREADS / 32][256];
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  base[t] = offs[(size_t)t * nb + blockIdx.x];
  const long long tile = (long long)blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_IPT; ++r) {
    const long long i = tile + (long long)r * RS_THREADS + t;
    if (tile + (long long)r * RS_THREADS >= n) break;                   // uniform: nothing left in the tile
    const bool valid = i < n;
    K key = K();
    if (valid) key = kin[i];
    const unsigned d = valid ? ((unsigned)(key >> shift) & 255u) : 256u;
#pragma unroll
    for (int ww = 0; ww < RS_THREADS / 32; ++ww) wcnt[ww][t] = 0;
    __syncthreads();
    const unsigned peers = __match_any_sync(DL_FULL, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));                // lanes before me with my digit
    if (valid && rank == 0) wcnt[w][d] = (unsigned short)__popc(peers);
    __syncthreads();
    unsigned run = 0;                                                   // thread t owns digit t
#pragma unroll
    for (int ww = 0; ww < RS_THREADS / 32; ++ww) {
      const unsigned c = wcnt[ww][t];
      wcnt[ww][t] = (unsigned short)run;
      run += c;
    }
    __syncthreads();
    if (valid) {
      const unsigned pos = base[d] + wcnt[w][d] + (unsigned)rank;
      kout[pos] = key;
      if (HASV) vout[pos] = vin[i];
    }
    __syncthreads();
    base[t] += run;
  }
}

inline size_t sort_ws_bytes(long long n) {
  const long long nb = (n + RS_TILE - 1) / RS_TILE;
  const long long cnt = 256 * (nb > 0 ? nb : 1);
  return align256((size_t)cnt * 4) + scan_ws_bytes(cnt, 4);
}

template <class K, class V, bool HASV>
int radix_sort_impl(K* kin, K* kout, V* vin, V* vout, long long n, int begin_bit, int end_bit, void* ws,
                    cudaStream_t st) {
  if (n <= 0) return DL_OK;
  if (n >= (1LL << 32)) return DL_EINVAL;                               // positions are 32-bit
  const long long nbl = (n + RS_TILE - 1) / RS_TILE;
  const int nb = (int)nbl;
  const long long cnt = 256LL * nb;
  unsigned* hist = (unsigned*)ws;
  void* scan_ws = (char*)ws + align256((size_t)cnt * 4);
  int passes = (end_bit - begin_bit + 7) / 8;
  if (passes < 1) passes = 1;
  K* a = kin; K* b = kout;
  V* va = vin; V* vb = vout;
  for (int p = 0; p < passes; ++p) {
    const int shift = begin_bit + 8 * p;
    k_rs_hist<K><<<nb, RS_THREADS, 0, st>>>(a, n, shift, hist, nb);
    DL_LAUNCH_CHECK();
    int rc = scan<false, unsigned>(LoadArr<unsigned>{hist}, StoreArr<unsigned>{hist}, cnt, OpSum<unsigned>(), 0u,
                                   scan_ws, st);
    if (rc) return rc;
    k_rs_scatter<K, V, HASV><<<nb, RS_THREADS, 0, st>>>(a, b, va, vb, n, shift, hist, nb);
    DL_LAUNCH_CHECK();
    K* tk = a; a = b; b = tk;
    V* tv = va; va = vb; vb = tv;
  }
  if (a != kout) {                                                      // even number of passes
    DL_CUDA_TRY(cudaMemcpyAsync(kout, a, (size_t)n * sizeof(K), cudaMemcpyDeviceToDevice, st));
    if (HASV) DL_CUDA_TRY(cudaMemcpyAsync(vout, va, (size_t)n * sizeof(V), cudaMemcpyDeviceToDevice, st));
  }
  return DL_OK;
}

template <class K>
int sort_keys(K* kin, K* kout, long long n, int begin_bit, int end_bit, void* ws, cudaStream_t st) {
  return radix_sort_impl<K, NoVal, false>(kin, kout, (NoVal*)nullptr, (NoVal*)nullptr, n, begin_bit, end_bit, ws, st);
}
template <class K, class V>
int sort_pairs(K* kin, K* kout, V* vin, V* vout, long long n, int begin_bit, int end_bit, void* ws,
               cudaStream_t st) {
  return radix_sort_impl<K, V, true>(kin, kout, vin, vout, n, begin_bit, end_bit, ws, st);
}

// ---------------------------------------------------------------------------------------------
// unique of a sorted array: out = the first element of every run, *count_out = number of runs
// ---------------------------------------------------------------------------------------------
template <class K>
struct FirstOfRun {
  const K* k;
  __device__ __forceinline__ unsigned operator()(long long i) const { return (i == 0 || k[i] != k[i - 1]) ? 1u : 0u; }
};
template <class K>
struct UniqueOut {
  const K* k;
  K* out;
  long long n;
  long long* count;
  __device__ __forceinline__ void operator()(long long i, unsigned excl) const {
    const bool first = (i == 0 || k[i] != k[i - 1]);
    if (first) out[excl] = k[i];
    if (i == n - 1) *count = (long long)excl + (first ? 1 : 0);
  }
};

inline size_t unique_ws_bytes(long long n) { return scan_ws_bytes(n, 4); }

template <class K>
int unique_sorted(const K* kin, K* kout, long long* count_out, long long n, void* ws, cudaStream_t st) {
  if (n <= 0) {
    DL_CUDA_TRY(cudaMemsetAsync(count_out, 0, sizeof(long long), st));
    return DL_OK;
  }
  if (n >= (1LL << 32)) return DL_EINVAL;
  return scan<false, unsigned>(FirstOfRun<K>{kin}, UniqueOut<K>{kin, kout, n, count_out}, n, OpSum<unsigned>(), 0u,
                               ws, st);
}

}  // namespace dlp
