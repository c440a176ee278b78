// Probe: which shared-memory image / descriptor does tcgen05.mma kind::tf32 accept for an
// MN-MAJOR A operand (A given as [k][m], m contiguous -- i.e. the transpose of a row-major tile)?
//
// Why: the backward GEMM 2 of the sum-product block (dW = r^T e, contraction over the samples)
// needs r and e with the SAMPLE index as K, while they arrive as [sample][unit] rows.  Round 1
// transposes them in registers (two 4x4 lane transposes + swizzled stores per float4,
// dense_tc.cu `bwd_transform`), which is a large share of the 8 400-clock operand transform.  If
// the tensor core can read [sample][unit] rows as an MN-major operand, the transposes go away.
//
// The instruction descriptor has the bits (a_major / b_major, 15 / 16) and CUTLASS's header says
// they are valid for TF32, but its layout selector only ever pairs 4-byte MN-major operands with
// the SWIZZLE_128B_BASE32B layout (32-byte swizzle granule, atom = 4 k-rows x 128 bytes).  This
// program tries the candidate images and descriptor encodings against a CPU product with
// integer-valued (tf32-exact) operands and prints the max error of each.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cirkit_b200/csrc -I include \
//             -o mn_major_probe scripts/micro/mn_major_probe.cu
// run  : ./mn_major_probe            (one B200; prints one line per variant, "OK" = exact)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "sm100.cuh"
using namespace ckb::sm100;

constexpr int M = 128, N = 64, K = 32;      // 4 k-steps of 8
constexpr int A_BYTES = M * K * 4;          // 16 KB
constexpr int B_BYTES = N * K * 4;          // 8 KB: one 128-byte row per n, K-major SWIZZLE_128B

struct Variant {
  const char* name;
  int layout_type;   // descriptor bits 61..63: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  int image;         // 0: 16-byte chunks xor (k & 7), atom 8 k-rows; 1: 32-byte chunks xor (k & 3), atom 4 k-rows
  int lbo, sbo;      // descriptor fields, bytes
  int kstep_bytes;   // start-address advance per k-step (8 k-rows)
};

// Byte offset of A[m][k] inside the image: the 32 m-values of an m-atom form one 128-byte row per
// k; all K rows of an m-atom are contiguous (so the next m-atom is K * 128 bytes further on).
static int a_offset(int image, int m, int k) {
  const int atom_m = m / 32, mm = m % 32;
  int row_off;
  if (image == 0) row_off = ((((mm / 4) ^ (k & 7)) & 7) << 4) + (mm & 3) * 4;
  else row_off = ((((mm / 8) ^ (k & 3)) & 3) << 5) + (mm & 7) * 4;
  return atom_m * (K * 128) + k * 128 + row_off;
}

__device__ __forceinline__ uint64_t make_desc_lt(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)lt << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe(const float* a_img, const float* b_img, float* d_out, int lt, int lbo, int sbo, int kstep) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  float* fa = (float*)sm;
  float* fb = (float*)(sm + A_BYTES);
  for (int i = threadIdx.x; i < A_BYTES / 4; i += blockDim.x) fa[i] = a_img[i];
  for (int i = threadIdx.x; i < B_BYTES / 4; i += blockDim.x) fb[i] = b_img[i];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 64);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tbase;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N, /*a_mn=*/1, /*b_mn=*/0);
    const uint64_t da = make_desc_lt(smem_u32(sm), lbo, sbo, lt);
    const uint64_t db = make_desc(smem_u32(sm) + A_BYTES, 16, 1024);
    for (int ks = 0; ks < K / 8; ++ks)
      mma_tf32(tb, desc_at(da, ks * kstep), desc_at(db, ks * 32), idesc, ks ? 1u : 0u);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * N + c + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tb, 64); }
}

int main() {
  std::vector<float> A(M * K), B(N * K), ref(M * N);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) A[m * K + k] = (float)(((m * 7 + k * 13) % 17) - 8);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) B[n * K + k] = (float)(((n * 5 + k * 3) % 11) - 5);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      ref[m * N + n] = s;
    }
  // B: K-major, 128-byte swizzle (the layout every kernel of this repo uses)
  std::vector<float> b_img(B_BYTES / 4, 0.f);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) b_img[(n * 128 + ((((k / 4) ^ n) & 7) << 4) + (k & 3) * 4) / 4] = B[n * K + k];

  const int atom = K * 128;  // bytes between m-atoms
  const Variant variants[] = {
      {"SW128        image16 LBO=atom SBO=1024", 2, 0, atom, 1024, 1024},
      {"SW128        image16 LBO=1024 SBO=atom", 2, 0, 1024, atom, 1024},
      {"SW128_BASE32 image32 LBO=atom SBO=512 ", 1, 1, atom, 512, 1024},
      {"SW128_BASE32 image32 LBO=512  SBO=atom", 1, 1, 512, atom, 1024},
      {"SW128_BASE32 image16 LBO=atom SBO=512 ", 1, 0, atom, 512, 1024},
      {"SW128        image32 LBO=atom SBO=1024", 2, 1, atom, 1024, 1024},
  };
  float *da, *db, *dd;
  cudaMalloc(&da, A_BYTES);
  cudaMalloc(&db, B_BYTES);
  cudaMalloc(&dd, M * N * 4);
  cudaMemcpy(db, b_img.data(), B_BYTES, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  for (const Variant& v : variants) {
    std::vector<float> a_img(A_BYTES / 4, 0.f);
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) a_img[a_offset(v.image, m, k) / 4] = A[m * K + k];
    cudaMemcpy(da, a_img.data(), A_BYTES, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, M * N * 4);
    probe<<<1, 128, 32 * 1024>>>(da, db, dd, v.layout_type, v.lbo, v.sbo, v.kstep_bytes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      // an illegal descriptor poisons the context: report and stop, re-run with the variant removed
      printf("%s : CUDA error %s\n", v.name, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> got(M * N);
    cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
    float err = 0.f;
    int bad = 0;
    for (int i = 0; i < M * N; ++i) {
      const float d = fabsf(got[i] - ref[i]);
      if (d > err) err = d;
      bad += d != 0.f;
    }
    printf("%s : max |err| %.1f, %d of %d entries differ  %s\n", v.name, err, bad, M * N, bad ? "" : "OK");
  }
  return 0;
}
