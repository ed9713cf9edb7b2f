// gather_probe.cu -- what does a B200 deliver for RANDOM gathers of small objects out of a large array?
//
// The slice gathers of the hot path (aggregation, backward pass 1: one 64-byte routed slice per CSR
// entry out of a 25.6 GB [N,K,d] array) sit at 0.38-0.47 of the HBM roofline.  This probe measures the
// ceiling the hardware sets for that access pattern independently of our kernels: uniformly random
// objects of 32 / 64 / 128 / 256 / 512 bytes, three ways of issuing the loads
//   ldg      ld.global.nc.v4 into registers, U independent loads in flight per thread
//   cpasync  cp.async.cg 16-byte pieces into a per-warp shared-memory ring (what the kernels do)
//   gather4  cp.async.bulk.tensor.2d ... tile::gather4 (TMA: 4 rows per instruction, mbarrier completion)
// and prints one JSON line per (method, object size, array size).  Standalone: nvcc only, no torch.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_bin/gather_probe tools/gather_probe.cu
//   tools/_bin/gather_probe [array_GB ...]        (default 25.6 and 1.0)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ uint32_t pick(uint64_t id, uint32_t n_objs) {
  return (uint32_t)(((uint64_t)mix32(id) * n_objs) >> 32);
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- ldg: LPO = lanes per object (S / 16); every warp instruction fetches 32/LPO objects ------------
template <int S, int U>
__global__ void __launch_bounds__(256) k_ldg(const float4* __restrict__ base, uint32_t n_objs, long long n_iter,
                                             float* sink) {
  constexpr int LPO = S / 16 < 1 ? 1 : S / 16, OPI = 32 / LPO;
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (long long it = gw; it < n_iter; it += nw) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t id = ((uint64_t)it * U + u) * OPI + lane / LPO;
      const uint32_t o = pick(id, n_objs);
      const float4* p = base + (size_t)o * (S / 16) + (lane % LPO);
      if (S == 32) {          // 2 lanes x 16 B
        asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
      } else {
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 1.2345678f) *sink = acc;
}

// ---- cp.async: per-warp ring of R stages, each stage = 32 lanes x 16 B (OPI objects) x U instructions --
template <int S, int U, int R>
__global__ void __launch_bounds__(512) k_cpasync(const float4* __restrict__ base, uint32_t n_objs, long long n_iter,
                                                 float* sink) {
  constexpr int LPO = S / 16 < 1 ? 1 : S / 16, OPI = 32 / LPO;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const unsigned ring = smem_u32(smem) + warp * (R * U * 512);
  float acc = 0.f;
  auto issue = [&](long long it, int slot) {
    if (it < n_iter) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint64_t id = ((uint64_t)it * U + u) * OPI + lane / LPO;
        const uint32_t o = pick(id, n_objs);
        const float4* p = base + (size_t)o * (S / 16) + (lane % LPO);
        const unsigned dst = ring + (slot * U + u) * 512 + lane * 16;
        if (S <= 64) asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
        else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  long long it = gw;
  for (int r = 0; r < R - 1; ++r) issue(it + (long long)r * nw, r);
  int slot = 0;
  for (; it < n_iter; it += nw) {
    int is = slot + R - 1; if (is >= R) is -= R;
    issue(it + (long long)(R - 1) * nw, is);
    asm volatile("cp.async.wait_group %0;" ::"n"(R - 1) : "memory");
    __syncwarp();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ring + (slot * U + u) * 512 + lane * 16));
      acc += v.x + v.y + v.z + v.w;
    }
    __syncwarp();
    slot = slot + 1 == R ? 0 : slot + 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (acc == 1.2345678f) *sink = acc;
}

// ---- TMA tile::gather4: lane 0 of every warp issues U gather4 (4 rows of S bytes each) per stage --------
template <int S, int U, int R>
__global__ void __launch_bounds__(512) k_gather4(const __grid_constant__ CUtensorMap tmap, uint32_t n_objs,
                                                 long long n_iter, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  constexpr int STAGE_B = U * 4 * S;
  const unsigned ring = smem_u32(smem) + warp * (R * STAGE_B);
  const unsigned bars = smem_u32(smem) + nwarp * (R * STAGE_B) + warp * (R * 8);
  if (lane == 0) {
    for (int r = 0; r < R; ++r) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + r * 8));
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  float acc = 0.f;
  auto issue = [&](long long it, int slot) {
    if (it < n_iter && lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + slot * 8), "r"(STAGE_B) : "memory");
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint64_t id = ((uint64_t)it * U + u) * 4;
        const int r0 = (int)pick(id, n_objs), r1 = (int)pick(id + 1, n_objs), r2 = (int)pick(id + 2, n_objs), r3 = (int)pick(id + 3, n_objs);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(ring + slot * STAGE_B + u * 4 * S),
            "l"(&tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bars + slot * 8)
            : "memory");
      }
    }
  };
  long long it = gw;
  for (int r = 0; r < R - 1; ++r) issue(it + (long long)r * nw, r);
  int slot = 0;
  unsigned phase = 0;
  for (; it < n_iter; it += nw) {
    int is = slot + R - 1; if (is >= R) is -= R;
    issue(it + (long long)(R - 1) * nw, is);
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bars + slot * 8),
        "r"((phase >> slot) & 1u)
        : "memory");
    phase ^= 1u << slot;
    for (int off = lane * 16; off < STAGE_B; off += 512) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ring + slot * STAGE_B + off));
      acc += v.x + v.y + v.z + v.w;
    }
    __syncwarp();
    slot = slot + 1 == R ? 0 : slot + 1;
  }
  if (acc == 1.2345678f) *sink = acc;
}

__global__ void k_fill(float4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (EncodeTiled)fn;
}

template <class F>
static double time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

static void report(const char* method, int S, double gb, long long n_obj_accessed, double ms, const char* note) {
  const double rate = n_obj_accessed / (ms * 1e-3);
  printf("{\"probe\": \"random_gather\", \"method\": \"%s\", \"object_bytes\": %d, \"array_gb\": %.2f, \"objects\": %lld, "
         "\"ms\": %.4f, \"objects_per_s\": %.4g, \"useful_gbs\": %.1f, \"note\": \"%s\"}\n",
         method, S, gb, n_obj_accessed, ms, rate, rate * S * 1e-9, note);
  fflush(stdout);
}

template <int S>
static void run_size(float4* base, size_t bytes, float* sink, EncodeTiled enc, int sms) {
  const double gb = bytes * 1e-9;
  const uint32_t n_objs = (uint32_t)(bytes / S);
  constexpr int LPO = S / 16, OPI = 32 / LPO;
  const long long target = 1LL << 27;                 // objects per launch (2^27 x 64 B = 8.6 GB)
  {
    constexpr int U = 8;
    const long long n_iter = target / (OPI * U);
    double ms = time_ms([&] { k_ldg<S, U><<<sms * 8, 256>>>(base, n_objs, n_iter, sink); }, 3);
    report("ldg", S, gb, n_iter * OPI * U, ms, "8 loads in flight per thread, 64 warps/SM");
  }
  {
    constexpr int U = 4, R = 4;                       // 16 warps x 4 stages x 2 KB = 128 KB, 96 KB in flight per SM
    const size_t smem = 16 * R * U * 512;
    CK(cudaFuncSetAttribute(k_cpasync<S, U, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_iter = target / (OPI * U);
    double ms = time_ms([&] { k_cpasync<S, U, R><<<sms, 512, smem>>>(base, n_objs, n_iter, sink); }, 3);
    report("cpasync", S, gb, n_iter * OPI * U, ms, "16 warps/SM, ring 4 x 2 KB per warp (96 KB in flight per SM)");
  }
  {
    constexpr int U = 2, R = 3;                       // a TMA-like depth: 32 KB in flight per SM
    const size_t smem = 16 * R * U * 512;
    CK(cudaFuncSetAttribute(k_cpasync<S, U, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_iter = target / (OPI * U);
    double ms = time_ms([&] { k_cpasync<S, U, R><<<sms, 512, smem>>>(base, n_objs, n_iter, sink); }, 3);
    report("cpasync_shallow", S, gb, n_iter * OPI * U, ms, "16 warps/SM, ring 3 x 1 KB per warp (32 KB in flight per SM)");
  }
  if (enc) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)(S / 4), (cuuint64_t)n_objs};
    cuuint64_t strides[1] = {(cuuint64_t)S};
    cuuint32_t box[2] = {(cuuint32_t)(S / 4), 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
      printf("{\"probe\": \"random_gather\", \"method\": \"gather4\", \"object_bytes\": %d, \"error\": \"cuTensorMapEncodeTiled rc=%d\"}\n", S, (int)rc);
    } else {
      constexpr int U = (S >= 256) ? 1 : 2, R = 4, NW = 16;
      constexpr int STAGE_B = U * 4 * S;
      const size_t smem = (size_t)NW * R * STAGE_B + NW * R * 8;
      CK(cudaFuncSetAttribute(k_gather4<S, U, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const long long n_iter = (target / 4) / (4 * U);          // TMA issue rate is the limit: fewer objects
      double ms = time_ms([&] { k_gather4<S, U, R><<<sms, NW * 32, smem>>>(tm, n_objs, n_iter, sink); }, 3);
      report("gather4", S, gb, n_iter * 4 * U, ms, "TMA tile::gather4, 16 warps/SM each issuing its own, ring of 4 stages");
    }
  }
}

int main(int argc, char** argv) {
  std::vector<double> sizes;
  for (int i = 1; i < argc; ++i) sizes.push_back(atof(argv[i]));
  if (sizes.empty()) { sizes.push_back(25.6); sizes.push_back(1.0); }
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  EncodeTiled enc = get_encode();
  float* sink;
  CK(cudaMalloc(&sink, 4));
  for (double gb : sizes) {
    const size_t bytes = ((size_t)(gb * 1e9) / 4096) * 4096;
    float4* base;
    CK(cudaMalloc(&base, bytes));
    k_fill<<<sms * 8, 256>>>(base, bytes / 16);
    CK(cudaDeviceSynchronize());
    {   // streaming reference on the same array (sequential objects: id instead of hash)
      cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      float4* dst; const size_t cp = bytes / 4 < (size_t)4e9 ? bytes / 4 : (size_t)4e9;
      CK(cudaMalloc(&dst, cp));
      CK(cudaMemcpy(dst, base, cp, cudaMemcpyDeviceToDevice));
      CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(dst, base, cp, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      printf("{\"probe\": \"copy\", \"array_gb\": %.2f, \"bytes\": %zu, \"ms\": %.4f, \"read_plus_write_gbs\": %.1f}\n", gb, cp, ms, 2.0 * cp / (ms * 1e-3) * 1e-9);
      CK(cudaFree(dst));
    }
    run_size<32>(base, bytes, sink, enc, sms);
    run_size<64>(base, bytes, sink, enc, sms);
    run_size<128>(base, bytes, sink, enc, sms);
    run_size<256>(base, bytes, sink, enc, sms);
    run_size<512>(base, bytes, sink, enc, sms);
    CK(cudaFree(base));
  }
  return 0;
}
