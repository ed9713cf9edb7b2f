// dl_fl.cuh -- small helpers shared by the factor-per-lane kernels (bwd_fl.cu, attn_fl.cu):
// shared-memory accesses through 32-bit shared addresses (the ring base is converted once), 16-byte
// and 4-byte cp.async, approximate reciprocal, fire-and-forget vector reduction.
#pragma once
#include "dl_common.cuh"

__device__ __forceinline__ float4 fl_lds4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float fl_lds1(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float fl_rcp(float x) {
  float v;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(x));
  return v;
}
__device__ __forceinline__ void fl_red_add4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void fl_cp16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void fl_cp4(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}


// small gathers (a 64-byte routed slice, a 4-byte per-node scalar): optional L2 prefetch-size
// qualifier, set per build for experiments (-DFL_SMALL_L2='".L2::64B"')
#ifndef FL_SMALL_L2
#define FL_SMALL_L2 ""
#endif
__device__ __forceinline__ void fl_cp16_small(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global" FL_SMALL_L2 " [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float fl_ldg_small(const float* p) {
  float v;
  asm volatile("ld.global.nc" FL_SMALL_L2 ".f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
