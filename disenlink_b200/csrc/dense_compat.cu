// dense_compat.cu -- dense [N,N] / [K,N,N] views for the small-N drop-in contract.
//
// The unmodified training script consumes Disentangle.forward's second return as a dense [N,N]
// tensor through boolean masks (main_disentangled.py:195,202,217) and Disentangle_layer.forward
// returns dense alpha0 / att (model.py:77).  These kernels materialise exactly those tensors from
// the sparse state so the module can stand in for model.py unchanged when N is small.  They are
// fp32 CUDA-core kernels on purpose: the factor projection is the only tensor-core contraction
// (BASELINE.json north_star).
#include "dl_common.cuh"

namespace {

constexpr int TP = 16;   // pair tile edge
constexpr int DC = 64;   // per-factor chunk staged in shared memory

// prob[u,v] = sigmoid( sum_k exp(<Z[u,k],Z[v,k]>/T) * <H[u,k],H[v,k]> )   [ref: model.py:109-113]
__global__ void __launch_bounds__(TP * TP)
k_allpairs_fwd(const float* __restrict__ Z, const float* __restrict__ H, long long N, int K, int d,
               float T, float* __restrict__ prob) {
  __shared__ float zu[TP][DC + 1], zv[TP][DC + 1], hu[TP][DC + 1], hv[TP][DC + 1];
  const int tu = threadIdx.y, tv = threadIdx.x;
  const long long u0 = (long long)blockIdx.y * TP, v0 = (long long)blockIdx.x * TP;
  const long long D = (long long)K * d;
  const int tid = tu * TP + tv;
  float S = 0.0f;
  for (int k = 0; k < K; ++k) {
    float q = 0.0f, gh = 0.0f;
    for (int c0 = 0; c0 < d; c0 += DC) {
      const int dc = min(DC, d - c0);
      __syncthreads();
      for (int x = tid; x < TP * dc; x += TP * TP) {
        const int rr = x / dc, cc = x % dc;
        const long long ur = u0 + rr, vr = v0 + rr;
        zu[rr][cc] = ur < N ? Z[ur * D + (long long)k * d + c0 + cc] : 0.0f;
        hu[rr][cc] = ur < N ? H[ur * D + (long long)k * d + c0 + cc] : 0.0f;
        zv[rr][cc] = vr < N ? Z[vr * D + (long long)k * d + c0 + cc] : 0.0f;
        hv[rr][cc] = vr < N ? H[vr * D + (long long)k * d + c0 + cc] : 0.0f;
      }
      __syncthreads();
      for (int x = 0; x < dc; ++x) {
        q = __fmaf_rn(zu[tu][x], zv[tv][x], q);
        gh = __fmaf_rn(hu[tu][x], hv[tv][x], gh);
      }
    }
    const float tvv = __fmul_rn(dl_expf(__fdiv_rn(q, T)), gh);
    S = (k == 0) ? tvv : __fadd_rn(S, tvv);
  }
  const long long u = u0 + tu, v = v0 + tv;
  if (u < N && v < N) prob[u * N + v] = dl_sigmoid(S);
}

// dSsym[n,o] = dS[n,o] + dS[o,n] with dS = dL/dlogit.  One warp per (node, factor); lanes over the
// factor's d elements; sequential over o (deterministic).
//   dH[n,k] = sum_o dSsym e_k H[o,k] ;  dZ[n,k] = sum_o dSsym e_k <H[n,k],H[o,k]>/T Z[o,k]
__global__ void __launch_bounds__(DL_CTA)
k_allpairs_bwd(const float* __restrict__ Z, const float* __restrict__ H,
               const float* __restrict__ dSsym, long long N, int K, int d, float T,
               float* __restrict__ dZ, float* __restrict__ dH) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * DL_WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * DL_WARPS_PER_CTA;
  const long long D = (long long)K * d;
  constexpr int R = DL_MAX_D / 32;
  for (long long x = warp0; x < N * K; x += nwarps) {
    const long long n = x / K;
    const int k = (int)(x % K);
    float zn[R], hn[R], az[R], ah[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      const int e = lane + 32 * t;
      zn[t] = e < d ? Z[n * D + (long long)k * d + e] : 0.0f;
      hn[t] = e < d ? H[n * D + (long long)k * d + e] : 0.0f;
      az[t] = 0.0f;
      ah[t] = 0.0f;
    }
    for (long long o = 0; o < N; ++o) {
      const float ds = __ldg(dSsym + n * N + o);
      if (ds == 0.0f) continue;  // warp-uniform: masked-out pairs carry exactly zero gradient
      float zo[R], ho[R];
      float pq = 0.0f, ph = 0.0f;
#pragma unroll
      for (int t = 0; t < R; ++t) {
        const int e = lane + 32 * t;
        zo[t] = e < d ? __ldg(Z + o * D + (long long)k * d + e) : 0.0f;
        ho[t] = e < d ? __ldg(H + o * D + (long long)k * d + e) : 0.0f;
        pq = __fmaf_rn(zn[t], zo[t], pq);
        ph = __fmaf_rn(hn[t], ho[t], ph);
      }
      for (int off = 16; off > 0; off >>= 1) {
        pq = __fadd_rn(pq, __shfl_xor_sync(DL_FULL, pq, off));
        ph = __fadd_rn(ph, __shfl_xor_sync(DL_FULL, ph, off));
      }
      const float ek = dl_expf(__fdiv_rn(pq, T));
      const float ch = __fmul_rn(ds, ek);
      const float cz = __fdiv_rn(__fmul_rn(ch, ph), T);
#pragma unroll
      for (int t = 0; t < R; ++t) {
        ah[t] = __fmaf_rn(ch, ho[t], ah[t]);
        az[t] = __fmaf_rn(cz, zo[t], az[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < R; ++t) {
      const int e = lane + 32 * t;
      if (e < d) {
        dZ[n * D + (long long)k * d + e] = az[t];
        dH[n * D + (long long)k * d + e] = ah[t];
      }
    }
  }
}

// alpha0[k,i,j] = exp(<Z[i,k],Z[j,k]>/T)   [ref: model.py:56-57]
__global__ void k_dense_alpha0(const float* __restrict__ Z, long long N, int K, int d, float T,
                               float* __restrict__ alpha0) {
  const long long D = (long long)K * d;
  const long long total = (long long)K * N * N;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
    const int k = (int)(x / (N * N));
    const long long i = (x / N) % N, j = x % N;
    const float* zi = Z + i * D + (long long)k * d;
    const float* zj = Z + j * D + (long long)k * d;
    float q = 0.0f;
    for (int t = 0; t < d; ++t) q = __fmaf_rn(zi[t], zj[t], q);
    alpha0[x] = dl_expf(__fdiv_rn(q, T));
  }
}

// att[kstar[e], i, j] = w[e] / s[j, kstar[e]] (att zero-filled by the caller)   [ref: model.py:70-74]
__global__ void k_dense_att(DlGraphDev g, const unsigned char* __restrict__ kstar,
                            const float* __restrict__ w, const float* __restrict__ s, int K,
                            float* __restrict__ att) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp0; i < g.N; i += nwarps) {
    for (long long p = g.rowptr[i] + lane; p < g.rowptr[i + 1]; p += 32) {
      const long long j = g.col[p];
      const int k = kstar[p];
      att[((long long)k * g.N + i) * g.N + j] = __fdiv_rn(w[p], s[j * K + k]);
    }
  }
}

}  // namespace

extern "C" {

int dl_allpairs_score_fwd(const float* Z, const float* H, int64_t N, int K, int d, float T,
                          float* prob, dl_stream_t stream) {
  if (N < 0 || K < 1 || K > DL_MAX_K || d < 1 || d > DL_MAX_D) return DL_EINVAL;
  if (N == 0) return DL_OK;
  if (!Z || !H || !prob || !(T == T) || T == 0.0f) return DL_EINVAL;
  long long tiles = (N + TP - 1) / TP;
  if (tiles > 65535) return DL_EUNSUPPORTED;  // dense [N,N] output: small N only
  dim3 grid((unsigned)tiles, (unsigned)tiles), block(TP, TP);
  k_allpairs_fwd<<<grid, block, 0, (cudaStream_t)stream>>>(Z, H, N, K, d, T, prob);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

int dl_allpairs_score_bwd(const float* Z, const float* H, const float* dSsym, int64_t N, int K, int d,
                          float T, float* dZ, float* dH, dl_stream_t stream) {
  if (N < 0 || K < 1 || K > DL_MAX_K || d < 1 || d > DL_MAX_D) return DL_EINVAL;
  if (N == 0) return DL_OK;
  if (!Z || !H || !dSsym || !dZ || !dH || !(T == T) || T == 0.0f) return DL_EINVAL;
  int grid = 1;
  int rc = dl_grid_for(k_allpairs_bwd, N * K, &grid);
  if (rc) return rc;
  k_allpairs_bwd<<<grid, DL_CTA, 0, (cudaStream_t)stream>>>(Z, H, dSsym, N, K, d, T, dZ, dH);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

int dl_dense_alpha0(const float* Z, int64_t N, int K, int d, float T, float* alpha0,
                    dl_stream_t stream) {
  if (N < 0 || K < 1 || K > DL_MAX_K || d < 1 || d > DL_MAX_D) return DL_EINVAL;
  if (N == 0) return DL_OK;
  if (!Z || !alpha0 || !(T == T) || T == 0.0f) return DL_EINVAL;
  long long total = (long long)K * N * N;
  long long b = (total + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  k_dense_alpha0<<<(int)b, 256, 0, (cudaStream_t)stream>>>(Z, N, K, d, T, alpha0);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

int dl_dense_att(const dl_graph* g_host, const uint8_t* kstar, const float* w, const float* s, int K,
                 float* att, dl_stream_t stream) {
  if (!dl_graph_ok(g_host) || K < 1 || K > DL_MAX_K) return DL_EINVAL;
  if (g_host->N == 0 || g_host->nnz == 0) return DL_OK;
  if (!kstar || !w || !s || !att) return DL_EINVAL;
  const DlGraphDev g = dl_graph_dev(g_host);
  long long b = (g.N * 32 + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  k_dense_att<<<(int)b, 256, 0, (cudaStream_t)stream>>>(g, kstar, w, s, K, att);
  DL_LAUNCH_CHECK();
  return DL_OK;
}

int dl_abi_version(void) { return 2; }

const char* dl_error_string(int code) {
  switch (code) {
    case DL_OK: return "ok";
    case DL_EINVAL: return "invalid argument";
    case DL_EWORKSPACE: return "workspace too small";
    case DL_ERANGE: return "index out of range";
    case DL_EASYM: return "adjacency pattern is not symmetric";
    case DL_EUNSUPPORTED: return "unsupported size";
    case DL_EINTERNAL: return "internal invariant violated (debug build)";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

}  // extern "C"
