// Shared device helpers and the internal launcher interface of libcirkit_b200.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "cirkit_b200.h"

namespace ckb {

constexpr int kNumSMs = 148;  // B200

// ------------------------------------------------------------------ error plumbing
void set_error(const char* fmt, ...);
#define CKB_CUDA_CHECK(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::ckb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                       \
      return CKB_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)
#define CKB_LAUNCH_CHECK() CKB_CUDA_CHECK(cudaGetLastError())

// Kernel attributes (the opt-in to > 48 KB of dynamic shared memory) belong to a device, not to
// the process: one mask bit per device ordinal, set the first time a launcher runs there.  A
// race between two host threads only repeats an idempotent cudaFuncSetAttribute call.
struct PerDeviceOnce {
  unsigned long long done = 0;
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done & bit) return false;
    done |= bit;
    return true;
  }
};

// ------------------------------------------------------------------ programmatic dependent launch
// The step is a chain of ~40 dependent launches, many of them 10-20 us long: with plain stream
// order every boundary costs the drain of one grid plus the launch latency and prologue of the
// next.  Kernels that opt in do the part of their prologue that touches no global memory written
// by kernels of the same step first (barrier set-up, TMEM allocation, weight staging), then
// pdl_wait() -- it returns when the preceding grid has completed and its writes are visible --
// and only THEN pdl_launch_dependents(): the next grid of the stream may be scheduled as soon as
// every CTA of this one has passed its own wait (or exited).  The order matters: a grid that
// triggers before it has waited lets its successor start while its PREDECESSOR is still running,
// and the successor's prologue (which reads effective weights) would then race with a parameter-op
// kernel two launches back.  With wait-then-trigger, "grid N+1 is running" implies "grid N-1 is
// complete" for every N.  Launch them with launch_pdl().
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();  // CKB_PDL=0 switches the launch attribute off (plain stream order)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ------------------------------------------------------------------ per-call context
struct Ctx {
  int64_t B;
  const void* xT;        // (D, B) int32 or float32
  int x_is_float;
  const uint8_t* maskT;  // (D, B) or (D, 1) bytes, or nullptr
  int64_t mask_ld;       // number of mask rows: B or 1
  float* const* tensors; // host array of device pointers
  float* const* grads;   // host array of device pointers (backward only)
  float* arena;
  float* garena;
  char* ws;
  size_t ws_bytes;
  cudaStream_t stream;
  int64_t launches;
};

// Each launcher enqueues the kernels of one step and returns a ckb_status.
int table_fwd(const ckb_step_desc_t& d, Ctx& c);
int table_bwd(const ckb_step_desc_t& d, Ctx& c);
int gaussian_fwd(const ckb_step_desc_t& d, Ctx& c);
int gaussian_bwd(const ckb_step_desc_t& d, Ctx& c);
int constant_fwd(const ckb_step_desc_t& d, Ctx& c);
int constant_bwd(const ckb_step_desc_t& d, Ctx& c);
int external_fwd(const ckb_step_desc_t& d, Ctx& c);
int external_bwd(const ckb_step_desc_t& d, Ctx& c);
int hadamard_fwd(const ckb_step_desc_t& d, Ctx& c);
int hadamard_bwd(const ckb_step_desc_t& d, Ctx& c);
int kronecker_fwd(const ckb_step_desc_t& d, Ctx& c);
int kronecker_bwd(const ckb_step_desc_t& d, Ctx& c);
int mixing_fwd(const ckb_step_desc_t& d, Ctx& c);
int mixing_bwd(const ckb_step_desc_t& d, Ctx& c);
int dense_fwd(const ckb_step_desc_t& d, Ctx& c);
int dense_bwd(const ckb_step_desc_t& d, Ctx& c);
int tucker_fwd(const ckb_step_desc_t& d, Ctx& c);
int tucker_bwd(const ckb_step_desc_t& d, Ctx& c);
int table_pair_gather(const ckb_step_desc_t& d, Ctx& c, float* u);
int table_dense_fwd(const ckb_step_desc_t& d, Ctx& c);
int table_dense_bwd(const ckb_step_desc_t& d, Ctx& c);
size_t table_dense_ws(const ckb_step_desc_t& d, int64_t B);
int param_op_fwd(const ckb_param_op_t& op, Ctx& c);
int param_op_bwd(const ckb_param_op_t& op, Ctx& c);
int multi_softmax(const ckb_param_op_t* ops, int n_ops, bool bwd, Ctx& c);

size_t table_bwd_ws(const ckb_step_desc_t& d, int64_t B);
size_t mixing_bwd_ws(const ckb_step_desc_t& d, int64_t B);
size_t dense_bwd_ws(const ckb_step_desc_t& d, int64_t B);
size_t tucker_ws(const ckb_step_desc_t& d, int64_t B);

// out[i] = sum_s partial[s*n + i]
int reduce_partials(const float* partial, float* out, int64_t n, int splits, Ctx& c);

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LSESumSemiring.apply_reduce clamps the row max to the finite range
// (cirkit/backend/torch/semiring.py:392-399) so that an all -inf row gives exp(-inf)=0.
__device__ __forceinline__ float clamp_max(float m) { return fminf(fmaxf(m, -FLT_MAX), FLT_MAX); }

// Gradient of one output element: the sum of the rows its consumers wrote (pull-based
// accumulation, see PlanLayout).  `gdirect` short-cuts the CSR for scratch-backed steps.
struct GradSrc {
  const float* garena;
  const int32_t* cons_ptr;
  const int64_t* cons_rows;
  int64_t B;
};
__device__ __forceinline__ float pull_grad(const GradSrc& g, int f, int64_t b, int K, int k) {
  if (g.cons_ptr == nullptr) return g.garena[((int64_t)f * g.B + b) * K + k];  // direct (F,B,K)
  float acc = 0.f;
  const int c1 = g.cons_ptr[f + 1];
  for (int c = g.cons_ptr[f]; c < c1; ++c) acc += g.garena[g.B * g.cons_rows[c] + b * K + k];
  return acc;
}

__device__ __forceinline__ int read_state(const void* xT, int x_is_float, int64_t idx) {
  // `.long()` of a float input truncates toward zero (layers/input.py:400-403)
  return x_is_float ? (int)((const float*)xT)[idx] : ((const int32_t*)xT)[idx];
}
__device__ __forceinline__ float read_value(const void* xT, int x_is_float, int64_t idx) {
  return x_is_float ? ((const float*)xT)[idx] : (float)((const int32_t*)xT)[idx];
}
// maskT is (D, mask_rows) with mask_rows == batch, or 1 to broadcast one row over the batch
__device__ __forceinline__ bool read_mask(const uint8_t* maskT, int64_t mask_rows, int var, int64_t b) {
  return maskT != nullptr && maskT[(int64_t)var * mask_rows + (mask_rows > 1 ? b : 0)] != 0;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
__host__ __device__ inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

}  // namespace ckb
