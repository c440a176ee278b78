// Device helpers shared by the tensor-core kernels (dense_tc.cu, tucker_tc.cu): streaming loads,
// MUFU-based exp/log selection, fp32 -> (tf32 hi, lo) splitting of vectors, lane transposes.
#pragma once
#include "common.cuh"
#include "sm100.cuh"

namespace ckb {
namespace sm100 {

__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <bool FAST>
__device__ __forceinline__ float exp_(float x) {
  return FAST ? fast_exp(x) : expf(x);
}
// exp(d) for d <= 0, d possibly -inf: one clamp keeps the MUFU path finite (exp(-104) flushes to 0)
template <bool FAST>
__device__ __forceinline__ float exp_nonpos(float d) {
  return FAST ? fast_exp_finite(fmaxf(d, -104.f)) : expf(d);
}
// exp(min(d, 88)): finite, so that a zero gradient times it stays zero
template <bool FAST>
__device__ __forceinline__ float exp_capped(float d) {
  return FAST ? fast_exp_finite(fminf(d, 88.f)) : expf(fminf(d, 88.f));
}
// 16-byte streaming load that leaves zeros when `ok` is false (no divergent branch)
__device__ __forceinline__ float4 ldg_stream_if(const float* p, bool ok) {
  float4 v;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
      "mov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
      "@q ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=&f"(v.x), "=&f"(v.y), "=&f"(v.z), "=&f"(v.w)
      : "l"(p), "r"((int)ok));
  return v;
}
template <bool FAST>
__device__ __forceinline__ float log_(float x) {
  return FAST ? fast_log(x) : logf(x);
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
  split_tf32(v.x, hi.x, lo.x);
  split_tf32(v.y, hi.y, lo.y);
  split_tf32(v.z, hi.z, lo.z);
  split_tf32(v.w, hi.w, lo.w);
}

// 4x4 transpose across the 4 lanes that differ in their two low lane bits: on entry lane j holds
// (row j, cols 0..3); on exit it holds (rows 0..3, col j).
__device__ __forceinline__ float4 transpose4(float4 v, int j) {
  const bool p = j & 1, q = j & 2;
  float s0 = p ? v.x : v.y, s1 = p ? v.z : v.w;
  float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
  if (p) { v.x = r0; v.z = r1; } else { v.y = r0; v.w = r1; }
  s0 = q ? v.x : v.z;
  s1 = q ? v.y : v.w;
  r0 = __shfl_xor_sync(0xffffffffu, s0, 2);
  r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
  if (q) { v.x = r0; v.y = r1; } else { v.z = r0; v.w = r1; }
  return v;
}


}  // namespace sm100

// process-wide switches (dense_tc.cu)
bool tc_disabled();
int tc_flags();

}  // namespace ckb
