// Probe 3 (round 2): register <-> TMEM cell mapping of the 16x256b shape of tcgen05.ld / tcgen05.st,
// and whether a warp can address the upper 16 lanes of its 32-lane quadrant with it.
//
// Why: the rewritten dense_tc backward wants ONE thread <-> element ownership for the global
// loads (8-byte pieces: 8 rows x 32 bytes per warp instruction = whole sectors), the r operand
// written to TMEM (TS-form GEMM 1), the MN-major shared-memory images of r and e (GEMM 2), the
// accumulator read-back and the du stores -- the mma-fragment-like 16x256b shape gives exactly that.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cirkit_b200/csrc -I include \
//             -o scripts/micro/tmem_shape_probe scripts/micro/tmem_shape_probe.cu
#include <cstdio>
#include <vector>
#include "sm100.cuh"
using namespace ckb::sm100;

__device__ __forceinline__ void ld_16x256b_x4(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st_16x256b_x4(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
      "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}

// mode 0: fill with 32x32b, read with 16x256b.x4 at lane offset `half*16`, column offset 32*chalf
// mode 1: write with 16x256b.x4 (value = tid*16 + reg), read everything back with 32x32b
__global__ void __launch_bounds__(128, 1) probe(float* out, int mode, int half, int chalf) {
  __shared__ uint32_t tbase;
  if (threadIdx.x < 32) tmem_alloc(&tbase, 64);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t quad = tb + ((uint32_t)(warp * 32) << 16);
  float v[16];
  if (mode == 0) {
    for (int cc = 0; cc < 64; cc += 16) {
      for (int j = 0; j < 16; ++j) v[j] = (float)((warp * 32 + lane) * 100 + cc + j);
      tmem_st16(quad + cc, v);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    ld_16x256b_x4(quad + ((uint32_t)(half * 16) << 16) + chalf * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = v[j];
  } else {
    for (int cc = 0; cc < 64; cc += 16) {
      for (int j = 0; j < 16; ++j) v[j] = -1.f;
      tmem_st16(quad + cc, v);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    for (int j = 0; j < 16; ++j) v[j] = (float)(threadIdx.x * 16 + j);
    st_16x256b_x4(quad + ((uint32_t)(half * 16) << 16) + chalf * 32, v);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    for (int cc = 0; cc < 64; cc += 16) {
      tmem_ld16(quad + cc, v);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + cc + j] = v[j];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tb, 64); }
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  std::vector<float> h(128 * 64);
  for (int half = 0; half < 2; ++half)
    for (int chalf = 0; chalf < 2; ++chalf) {
      cudaMemset(d, 0, 128 * 64 * 4);
      probe<<<1, 128>>>(d, 0, half, chalf);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ld half=%d chalf=%d: CUDA error %s\n", half, chalf, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), d, 128 * 16 * 4, cudaMemcpyDeviceToHost);
      // expected: thread t (of warp w), reg 4n+2a+c  <->  lane w*32 + half*16 + t/4 + 8a, col chalf*32 + 8n + 2(t%4) + c
      int ok = 1;
      for (int t = 0; t < 128; ++t)
        for (int r = 0; r < 16; ++r) {
          const int w = t / 32, l = t % 32, n = r / 4, a = (r / 2) % 2, c = r % 2;
          const int lane = w * 32 + half * 16 + l / 4 + 8 * a, col = chalf * 32 + 8 * n + 2 * (l % 4) + c;
          ok &= h[t * 16 + r] == (float)(lane * 100 + col);
        }
      printf("ld.16x256b.x4 lane offset %2d col offset %2d: expected fragment mapping %s\n", half * 16, chalf * 32, ok ? "OK" : "NO");
      if (!ok) {
        for (int t = 0; t < 8; ++t) {
          printf("   thread %d:", t);
          for (int r = 0; r < 16; ++r) printf(" (%d,%d)", (int)h[t * 16 + r] / 100, (int)h[t * 16 + r] % 100);
          printf("\n");
        }
      }
    }
  for (int half = 0; half < 2; ++half) {
    cudaMemset(d, 0, 128 * 64 * 4);
    probe<<<1, 128>>>(d, 1, half, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("st half=%d: CUDA error %s\n", half, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    int ok = 1;
    for (int lane = 0; lane < 128; ++lane)
      for (int col = 0; col < 64; ++col) {
        float want = -1.f;
        const int w = lane / 32, lr = lane % 32;
        if (lr / 16 == half && col >= 32) {
          const int rr = lr % 16, cc = col - 32;
          const int t = w * 32 + (rr % 8) * 4 + (cc % 8) / 2, reg = (cc / 8) * 4 + (rr / 8) * 2 + cc % 2;
          want = (float)(t * 16 + reg);
        }
        ok &= h[lane * 64 + col] == want;
      }
    printf("st.16x256b.x4 lane offset %2d col offset 32: expected fragment mapping %s\n", half * 16, ok ? "OK" : "NO");
  }
  return 0;
}
