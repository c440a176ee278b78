// Micro-benchmark: cycles per tcgen05.mma (kind::tf32, M=128, K=8) issued back to back by one
// thread, for N = 64/128/256 with the A operand in shared memory (SS) or in TMEM (TS), plus the
// cost of tcgen05.commit and of an already-satisfied mbarrier try_wait.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cirkit_b200/csrc -I include -o mma_rate scripts/micro/mma_rate.cu
#include <cstdio>
#include "sm100.cuh"
using namespace ckb::sm100;

__device__ __forceinline__ void mma_ts_elect(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_elect(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int ts>
__global__ void __launch_bounds__(128, 1) k(long long* out, int n_cols, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  float* f = (float*)sm;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) f[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tbase;
  if (ts >= 2) {
    if (threadIdx.x < 32) {   // whole warp, converged; the MMA is predicated on the elected lane
      const uint32_t idesc = make_idesc_tf32(128, n_cols, 0, 0);
      const uint64_t da = make_desc(smem_u32(sm), 16, 1024);
      const uint64_t db = make_desc(smem_u32(sm) + 16 * 1024, 16, 1024);
      long long t0 = clock64();
#pragma unroll 4
      for (int r = 0; r < reps; ++r) {
        if (ts == 4) mma_ts_elect(tb, tb + 256, db, idesc, 1u);
        else if (ts == 5) mma_ss_elect(tb, da, db, idesc, 1u);
        else if (ts == 3) mma_ts_elect(tb, tb + 256 + (r & 3) * 8, desc_at(db, (r & 3) * 32), idesc, 1u);
        else mma_ss_elect(tb, desc_at(da, (r & 3) * 32), desc_at(db, (r & 3) * 32), idesc, 1u);
      }
      long long t1 = clock64();
      if (threadIdx.x == 0) mma_commit(&bar);
      __syncwarp();
      long long t2 = clock64();
      mbar_wait(&bar, 0);
      long long t3 = clock64();
      if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = 0; out[4] = 0; }
    }
  } else if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(128, n_cols, 0, 0);
    const uint64_t da = make_desc(smem_u32(sm), 16, 1024);
    const uint64_t db = make_desc(smem_u32(sm) + 16 * 1024, 16, 1024);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (ts) mma_tf32_ts(tb, tb + 256 + (r & 3) * 8, desc_at(db, (r & 3) * 32), idesc, 1u);
      else mma_tf32(tb, desc_at(da, (r & 3) * 32), desc_at(db, (r & 3) * 32), idesc, 1u);
    }
    long long t1 = clock64();
    mma_commit(&bar);
    long long t2 = clock64();
    mbar_wait(&bar, 0);
    long long t3 = clock64();
    // satisfied try_wait cost
    for (int r = 0; r < 16; ++r) mbar_wait(&bar, 0);
    long long t4 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = (t4 - t3) / 16;
    fence_proxy_async_smem();
    long long t5 = clock64();
    out[4] = t5 - t4;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int reps = 512;
  void (*kern[6])(long long*, int, int) = {k<0>, k<1>, k<2>, k<3>, k<4>, k<5>};
  for (int ts = 0; ts < 6; ++ts)
    for (int n : {64, 128, 256}) {
      long long h[8];
      cudaFuncSetAttribute(kern[ts], cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      for (int it = 0; it < 2; ++it) {
        kern[ts]<<<1, 128, 64 * 1024>>>(d, n, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      printf("%s N=%3d: issue %.1f clk/mma, issue+drain %.1f clk/mma, commit %lld, drain %lld, try_wait(satisfied) %lld, fence.proxy.async %lld\n",
             ts == 0 ? "SS-lane0" : ts == 1 ? "TS-lane0" : ts == 2 ? "SS-elect" : ts == 3 ? "TS-elect" : ts == 4 ? "TS-elect-const" : "SS-elect-const", n, (double)h[0] / reps, (double)(h[0] + h[1] + h[2]) / reps, h[1], h[2], h[3], h[4]);
    }
  return 0;
}
