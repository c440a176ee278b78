// Probe 2 (round 2): questions that decide the layout of the rewritten dense_tc backward kernel.
//   (1) does a K-MAJOR tf32 operand work with the SWIZZLE_128B_BASE32B layout type (32-byte
//       swizzle granule)?  If so ONE [sample][unit] image of r serves GEMM 1 (K-major A, K = unit)
//       and GEMM 2 (MN-major A, K = sample) of the backward pass;
//   (2) does an MN-MAJOR B operand work with the same image as the MN-major A that probe 1 found;
//   (3) where do the rows of an M = 64 accumulator live in TMEM (expected: row m -> lane
//       (m % 16) + 32 * (m / 16));
//   (4) are tf32 subnormal operands / fp32 subnormal products preserved by the tensor core
//       (reference edge case tests/backend/torch/test_semiring.py:41-61: weight 1e-38).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cirkit_b200/csrc -I include \
//             -o scripts/micro/umma_probe2 scripts/micro/umma_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "sm100.cuh"
using namespace ckb::sm100;

constexpr int MAXM = 128, N = 64, K = 32;
constexpr int A_BYTES = MAXM * K * 4;  // 16 KB
constexpr int B_BYTES = N * K * 4;     // 8 KB

__device__ __forceinline__ uint64_t make_desc_lt(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)lt << 61;
  return d;
}

struct Cfg {
  int M;
  int a_mn, b_mn;
  int a_lt, b_lt;
  int a_lbo, a_sbo, a_kstep;
  int b_lbo, b_sbo, b_kstep;
};

__global__ void __launch_bounds__(128, 1) probe(const float* a_img, const float* b_img, float* d_out, Cfg c) {
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {  // poison the accumulator so that untouched lanes are recognisable
    float v[16];
    for (int j = 0; j < 16; ++j) v[j] = -12345.f;
    for (int cc = 0; cc < N; cc += 16) tmem_st16(tb + ((uint32_t)(warp * 32) << 16) + cc, v);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(c.M, N, c.a_mn, c.b_mn);
    const uint64_t da = make_desc_lt(smem_u32(sm), c.a_lbo, c.a_sbo, c.a_lt);
    const uint64_t db = make_desc_lt(smem_u32(sm) + A_BYTES, c.b_lbo, c.b_sbo, c.b_lt);
    for (int ks = 0; ks < K / 8; ++ks)
      mma_tf32(tb, desc_at(da, ks * c.a_kstep), desc_at(db, ks * c.b_kstep), idesc, ks ? 1u : 0u);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int cc = 0; cc < N; cc += 16) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + cc, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * N + cc + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tb, 64); }
}

// images -------------------------------------------------------------------------------------
// K-major, rows of 128 bytes (32 k's); swz16: 16-byte chunk ^ (row & 7); swz32: 32-byte chunk ^ (row & 3)
static int kmajor_off(int swz32, int row, int k) {
  if (!swz32) return row * 128 + ((((k / 4) ^ row) & 7) << 4) + (k & 3) * 4;
  return row * 128 + ((((k / 8) ^ row) & 3) << 5) + (k & 7) * 4;
}
// MN-major: one 128-byte row per k holding 32 consecutive m's; m-atoms K*128 bytes apart;
// 32-byte chunk ^ (k & 3)
static int mnmajor_off(int m, int k) {
  const int atom = m / 32, mm = m % 32;
  return atom * (K * 128) + k * 128 + ((((mm / 8) ^ (k & 3)) & 3) << 5) + (mm & 7) * 4;
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;  // run one test per process: a bad descriptor poisons the context
  std::vector<float> A(MAXM * K), B(N * K);
  for (int m = 0; m < MAXM; ++m)
    for (int k = 0; k < K; ++k) A[m * K + k] = (float)(((m * 7 + k * 13) % 17) - 8);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) B[n * K + k] = (float)(((n * 5 + k * 3) % 11) - 5);
  float *da, *db, *dd;
  cudaMalloc(&da, A_BYTES);
  cudaMalloc(&db, B_BYTES);
  cudaMalloc(&dd, MAXM * N * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);

  struct Test { const char* name; Cfg c; int a_img, b_img; };  // img: 0 K-major swz16, 1 K-major swz32, 2 MN-major
  const int atom = K * 128;
  const Test tests[] = {
      {"baseline  A K-major SW128          x B K-major SW128        ", {128, 0, 0, 2, 2, 16, 1024, 32, 16, 1024, 32}, 0, 0},
      {"(1)       A K-major SW128_BASE32   x B K-major SW128        ", {128, 0, 0, 1, 2, 16, 1024, 32, 16, 1024, 32}, 1, 0},
      {"(1b)      A K-major SW128_BASE32 SBO=512 x B K-major SW128  ", {128, 0, 0, 1, 2, 16, 512, 32, 16, 1024, 32}, 1, 0},
      {"(1c)      A K-major SW128 x B K-major SW128_BASE32          ", {128, 0, 0, 2, 1, 16, 1024, 32, 16, 1024, 32}, 0, 1},
      {"(2)       A K-major SW128 x B MN-major SW128_BASE32         ", {128, 0, 1, 2, 1, 16, 1024, 32, atom, 512, 1024}, 0, 2},
      {"(2b)      A MN-major BASE32 x B MN-major BASE32             ", {128, 1, 1, 1, 1, atom, 512, 1024, atom, 512, 1024}, 2, 2},
      {"(3)       M=64: A K-major SW128 x B K-major SW128           ", {64, 0, 0, 2, 2, 16, 1024, 32, 16, 1024, 32}, 0, 0},
  };
  int ti = -1;
  for (const Test& t : tests) {
    ++ti;
    if (only >= 0 && ti != only) continue;
    std::vector<float> a_img(A_BYTES / 4, 0.f), b_img(B_BYTES / 4, 0.f);
    for (int m = 0; m < MAXM; ++m)
      for (int k = 0; k < K; ++k) {
        const int off = t.a_img == 2 ? mnmajor_off(m, k) : kmajor_off(t.a_img, m, k);
        a_img[off / 4] = A[m * K + k];
      }
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        const int off = t.b_img == 2 ? mnmajor_off(n, k) : kmajor_off(t.b_img, n, k);
        b_img[off / 4] = B[n * K + k];
      }
    cudaMemcpy(da, a_img.data(), A_BYTES, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), B_BYTES, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, MAXM * N * 4);
    probe<<<1, 128, 32 * 1024>>>(da, db, dd, t.c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s : CUDA error %s\n", t.name, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> got(MAXM * N);
    cudaMemcpy(got.data(), dd, MAXM * N * 4, cudaMemcpyDeviceToHost);
    if (t.c.M == 128) {
      float err = 0.f;
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          float s = 0.f;
          for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
          const float d = fabsf(got[m * N + n] - s);
          if (d > err) err = d;
          bad += d != 0.f;
        }
      printf("%s : max |err| %.1f, %d of %d differ  %s\n", t.name, err, bad, 128 * N, bad ? "" : "OK");
    } else {
      // which TMEM lane holds which row?
      printf("%s :\n   lane -> row: ", t.name);
      int ok_expected = 1;
      for (int lane = 0; lane < 128; ++lane) {
        int row = -1;
        if (got[lane * N] != -12345.f) {
          for (int m = 0; m < 64 && row < 0; ++m) {
            int match = 1;
            for (int n = 0; n < N && match; ++n) {
              float s = 0.f;
              for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
              match = got[lane * N + n] == s;
            }
            if (match) row = m;
          }
          if (row < 0) row = -2;  // written, but no row matches
        }
        const int expect = (lane % 32) < 16 ? (lane / 32) * 16 + lane % 32 : -1;
        ok_expected &= row == expect;
        if (lane % 16 == 0) printf("| %d:", lane);
        printf("%d ", row);
      }
      printf("\n   expected (m%%16)+32*(m/16): %s\n", ok_expected ? "OK" : "NO");
    }
  }

  // (4) subnormals: A = 1 in column 0, B[n][0] = 1e-38 * (n+1); and A = 1e-38 with B = 1
  if (only < 0 || only == 7) {
    std::vector<float> a_img(A_BYTES / 4, 0.f), b_img(B_BYTES / 4, 0.f);
    for (int m = 0; m < 128; ++m) a_img[kmajor_off(0, m, 0) / 4] = m < 64 ? 1.0f : 1e-38f;
    for (int n = 0; n < N; ++n) b_img[kmajor_off(0, n, 0) / 4] = n < 32 ? 1e-38f * (float)(n + 1) : 1.0f;
    cudaMemcpy(da, a_img.data(), A_BYTES, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), B_BYTES, cudaMemcpyHostToDevice);
    Cfg c{128, 0, 0, 2, 2, 16, 1024, 32, 16, 1024, 32};
    probe<<<1, 128, 32 * 1024>>>(da, db, dd, c);
    cudaDeviceSynchronize();
    std::vector<float> got(MAXM * N);
    cudaMemcpy(got.data(), dd, MAXM * N * 4, cudaMemcpyDeviceToHost);
    printf("(4) subnormals: 1 x 1e-38 -> %g (want 1e-38), 1 x 3e-38 -> %g, 1e-38 x 1 -> %g, 1e-38 x 1e-38 -> %g\n",
           got[0 * N + 0], got[0 * N + 2], got[64 * N + 32], got[64 * N + 0]);
  }
  return 0;
}
