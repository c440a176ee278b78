// In-switch sum of the replicas' gradient buffers over NVLink / NVSwitch multicast ("two-shot"
// all-reduce): the one exchange of the data-parallel path that carries real bytes (SURVEY 8(e):
// 77 MB at the north-star shape).
//
// Every rank holds the same buffer at the same offset of a multicast object (symmetric memory: the
// host side allocates it once and exchanges the handles, cirkit_b200/distributed.py).  Rank r owns
// elements [r n / N, (r + 1) n / N): for each 16-byte piece of its share it issues
//   multimem.ld_reduce.add.v4.f32  -- the switch fetches the piece from all N replicas, adds them
//                                     in fp32 and returns the sum,
//   multimem.st.v4.f32             -- the switch writes the sum into all N replicas,
// so a byte crosses a GPU's links once in each direction and no SM adds anything.  Every element is
// summed by exactly one rank and broadcast: all replicas end up with the same bits.
//
// Ordering is the caller's: a cross-GPU barrier on the stream before the launch (every replica's
// backward pass has written its gradients) and one after it (every share has been broadcast).
// NCCL's own all-reduce of the same 77 MB takes 312 us on 8 B200s (247 GB/s per rank,
// scripts/allreduce_probe.py); see DESIGN.md section 6 for what this kernel measures.
#include <stdlib.h>

#include "common.cuh"

namespace ckb {
namespace {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

template <int UNROLL>
__global__ void __launch_bounds__(512) nvls_allreduce_kernel(float* __restrict__ mc, int64_t first4,
                                                             int64_t count4) {
  // `mc`: multicast address of the buffer; this rank's share is float4 pieces [first4, first4 + count4)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < count4; i += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = multimem_ld_reduce_add(mc + 4 * (first4 + i + u * stride));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) multimem_st(mc + 4 * (first4 + i + u * stride), v[u]);
  }
  for (; i < count4; i += stride) multimem_st(mc + 4 * (first4 + i), multimem_ld_reduce_add(mc + 4 * (first4 + i)));
}

// ---- the same exchange as ONE kernel: cross-GPU ordering and the copy into the caller's buffer
// included.  The two stream barriers around the kernel above cost ~27 us each on 8 GPUs (a launch
// plus a flag round trip), a separate copy-out another launch; here
//   (1) block 0 tells every peer "my gradients are written" (st.release.sys into the peer's signal
//       pad -- peer-mapped memory the symmetric-memory rendezvous hands out), every block waits
//       until all peers have said so (ld.acquire.sys on its own pad);
//   (2) the share of this rank is reduced in the switch and broadcast, as above;
//   (3) the last block to finish tells every peer "my share is broadcast"; every block waits for
//       all peers' flags and then
//   (4) copies its part of the now complete buffer into `out` (the tensors autograd hands out must
//       not alias a buffer the next backward pass overwrites).
// Flags carry the call's epoch (monotonic), so they never need resetting.  All blocks spin, so the
// grid must be co-resident: at most one CTA per SM.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
constexpr int kPadReady = 256, kPadDone = 320;  // word offsets inside a signal pad (64 ranks each)

template <int UNROLL>
__global__ void __launch_bounds__(512) nvls_allreduce_fused_kernel(float* __restrict__ mc,
                                                                   const float* __restrict__ local,
                                                                   float* __restrict__ out, int64_t n4,
                                                                   int64_t first4, int64_t count4,
                                                                   uint32_t* const* __restrict__ pads, int rank,
                                                                   int world, uint32_t epoch,
                                                                   unsigned int* __restrict__ counter) {
  uint32_t* my_pad = pads[rank];
  // (1) ready: this rank's gradients were written by earlier kernels of this stream
  if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(pads[threadIdx.x] + kPadReady + rank, epoch);
  if (threadIdx.x < world)
    while ((int32_t)(ld_acquire_sys(my_pad + kPadReady + threadIdx.x) - epoch) < 0) {
    }
  __syncthreads();
  // (2) reduce + broadcast this rank's share
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < count4; i += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = multimem_ld_reduce_add(mc + 4 * (first4 + i + u * stride));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) multimem_st(mc + 4 * (first4 + i + u * stride), v[u]);
  }
  for (; i < count4; i += stride) multimem_st(mc + 4 * (first4 + i), multimem_ld_reduce_add(mc + 4 * (first4 + i)));
  // (3) done: every block of this rank has issued its stores -> the last one tells the peers
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    last = prev + 1 == gridDim.x;
    if (last) *counter = 0;  // ready for the next call (nobody reads it before that call's kernel)
  }
  __syncthreads();
  if (last && threadIdx.x < world) st_release_sys(pads[threadIdx.x] + kPadDone + rank, epoch);
  if (threadIdx.x < world)
    while ((int32_t)(ld_acquire_sys(my_pad + kPadDone + threadIdx.x) - epoch) < 0) {
    }
  __syncthreads();
  // (4) the whole buffer is reduced: copy it out
  const float4* src = reinterpret_cast<const float4*>(local);
  float4* dst = reinterpret_cast<float4*>(out);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += stride) dst[j] = src[j];
}

}  // namespace
}  // namespace ckb

using namespace ckb;

extern "C" int ckb_nvls_allreduce(void* multicast_ptr, int64_t num_floats, int32_t rank, int32_t world,
                                      int32_t num_ctas, void* stream) {
  if (multicast_ptr == nullptr || num_floats <= 0 || (num_floats & 3) || ((uintptr_t)multicast_ptr & 15) ||
      world <= 0 || rank < 0 || rank >= world) {
    set_error("ckb_nvls_allreduce: needs a 16-byte aligned multicast address, a multiple of 4 floats and a valid rank");
    return CKB_ERR_INVALID;
  }
  const int64_t n4 = num_floats / 4;
  const int64_t first4 = n4 * rank / world, last4 = n4 * (rank + 1) / world;
  if (last4 == first4) return CKB_OK;
  const int ctas = num_ctas > 0 ? num_ctas : kNumSMs;
  static int unroll = 0;  // CKB_NVLS_UNROLL: loads in flight per thread (tuning probe), default 4
  if (unroll == 0) {
    const char* e = getenv("CKB_NVLS_UNROLL");
    unroll = e ? atoi(e) : 4;
  }
  float* mc = (float*)multicast_ptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (unroll >= 8) nvls_allreduce_kernel<8><<<ctas, 512, 0, st>>>(mc, first4, last4 - first4);
  else if (unroll <= 2) nvls_allreduce_kernel<2><<<ctas, 512, 0, st>>>(mc, first4, last4 - first4);
  else nvls_allreduce_kernel<4><<<ctas, 512, 0, st>>>(mc, first4, last4 - first4);
  CKB_LAUNCH_CHECK();
  return CKB_OK;
}

extern "C" int ckb_nvls_allreduce_fused(void* multicast_ptr, const float* local_ptr, float* out,
                                        int64_t num_floats, int32_t rank, int32_t world,
                                        void* signal_pads_dev, uint32_t epoch, void* counter, int32_t num_ctas,
                                        void* stream) {
  if (multicast_ptr == nullptr || local_ptr == nullptr || out == nullptr || signal_pads_dev == nullptr ||
      counter == nullptr || num_floats <= 0 || (num_floats & 3) || ((uintptr_t)multicast_ptr & 15) ||
      ((uintptr_t)local_ptr & 15) || ((uintptr_t)out & 15) || world <= 0 || world > 64 || rank < 0 ||
      rank >= world || epoch == 0) {
    set_error("ckb_nvls_allreduce_fused: bad arguments (16-byte aligned buffers, a multiple of 4 floats, "
              "world <= 64, epoch >= 1)");
    return CKB_ERR_INVALID;
  }
  const int64_t n4 = num_floats / 4;
  const int64_t first4 = n4 * rank / world, last4 = n4 * (rank + 1) / world;
  // every block spins on flags: the grid has to be resident as a whole
  const int ctas = num_ctas > 0 ? (num_ctas < kNumSMs ? num_ctas : kNumSMs) : 64;
  nvls_allreduce_fused_kernel<4><<<ctas, 512, 0, (cudaStream_t)stream>>>(
      (float*)multicast_ptr, local_ptr, out, n4, first4, last4 - first4, (uint32_t* const*)signal_pads_dev, rank,
      world, epoch, (unsigned int*)counter);
  CKB_LAUNCH_CHECK();
  return CKB_OK;
}
