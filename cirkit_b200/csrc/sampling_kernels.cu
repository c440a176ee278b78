// Ancestral (top-down) sampling from a smooth, decomposable circuit -- SamplingQuery,
// cirkit/backend/torch/queries.py:187-275 with the per-layer rules of layers/inner.py:129-133
// (Hadamard), :189-197 (Kronecker), :275-300 (Sum), layers/optimized.py:180-202 (CP-T) and
// layers/input.py:423-434 (Categorical), :680-685 (Gaussian).
//
// The reference samples bottom-up: EVERY unit of every layer draws num_samples values of every
// variable ((F, K, N, D) tensors), and a sum unit then picks, per sample, which input unit's
// values to keep.  Only the picks on the path from the root unit survive, so this file walks the
// circuit the other way: a "selection arena" holds, for every (layer, fold) row and every sample,
// the unit the sample's path goes through (-1: the path does not visit the row).  Layers are
// visited root first; a sum row draws its mixture component from the row-wise CDF of its weights
// and writes the selected unit into the rows of its inputs; product rows pass the selection on;
// input rows draw the variable.  Work and memory are O(sum_layers F * N), not O(F * K * N * D).
// The distribution of the returned samples is the reference's: sample n of the root unit uses
// independent draws along its own induced tree in both schemes.
//
// Randomness: Philox4x32-10, counter = (sample index, arena row), key = seed -- every (row,
// sample) pair owns one 128-bit block, so results do not depend on the launch geometry and the
// CPU oracle (oracle/sampling.py) reproduces the stream.
#include "common.cuh"

namespace ckb {
namespace {

__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// 24-bit uniform in [0, 1)
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// First index j with cdf[j] > u * cdf[n-1] (inverse-CDF draw from the unnormalised row); entries
// of probability zero (cdf[j] == cdf[j-1]) are never returned.
__device__ __forceinline__ int draw(const float* __restrict__ cdf, int n, float u) {
  const float t = u * cdf[n - 1];
  int lo = 0, hi = n - 1;  // invariant: the answer is in [lo, hi]
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] > t) hi = mid;
    else lo = mid + 1;
  }
  while (lo > 0 && cdf[lo] == cdf[lo - 1]) --lo;  // t rounded up to the total: step off a flat tail
  return lo;
}

// Row-wise inclusive prefix sums, one thread per row, summed left to right in fp32 (the order the
// oracle's numpy cumsum uses).  mode 0: src (rows, cols) non-negative weights.  mode 1: src is a
// log-probability table laid out (F, V, K) (the layout of the table kernels); dst is (F, K, V).
__global__ void cdf_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows,
                                int cols, int mode, int K) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float acc = 0.f;
  if (mode == 0) {
    for (int c = 0; c < cols; ++c) {
      acc += src[r * cols + c];
      dst[r * cols + c] = acc;
    }
  } else {
    const int64_t f = r / K;
    const int k = (int)(r % K);
    for (int v = 0; v < cols; ++v) {
      acc += expf(src[(f * cols + v) * K + k]);
      dst[r * cols + v] = acc;
    }
  }
}

struct SampleArgs {
  ckb_sample_step_t s;
  int64_t N;
  int64_t base;  // global index of sample 0 of this call (Philox counter)
  int32_t* sel;  // (rows, N) selected unit per arena row and sample, -1 = not on the path
  int32_t* mix;  // (rows, N) mixture component drawn by sum rows (-1 elsewhere), may be NULL
  void* x;       // (N, D) int64 or float32
  int D, x_is_float;
  uint32_t k0, k1;
};

__global__ void sample_step_kernel(const SampleArgs a) {
  const ckb_sample_step_t& s = a.s;
  const int f = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const int64_t row = s.sel_row + f;
  const int o = a.sel[row * a.N + n];
  if (o < 0) return;
  const uint64_t gn = (uint64_t)(a.base + n);
  uint32_t c[4] = {(uint32_t)gn, (uint32_t)(gn >> 32), (uint32_t)row, 0u};
  philox4x32_10(c, a.k0, a.k1);
  const int H = s.arity;
  const int32_t* in = s.in_sel_rows ? s.in_sel_rows + (int64_t)f * H : nullptr;
  switch (s.kind) {
    case CKB_STEP_DENSE: {
      const int Kred = (s.flags & CKB_DENSE_CONCAT) ? H * s.k_in : s.k_in;
      const int j = draw(s.cdf + ((int64_t)f * s.k_out + o) * Kred, Kred, u01(c[0]));
      if (a.mix) a.mix[row * a.N + n] = j;
      if (s.flags & CKB_DENSE_CONCAT) {
        const int h = j / s.k_in;
        a.sel[(int64_t)in[h] * a.N + n] = j - h * s.k_in;
      } else {
        for (int h = 0; h < H; ++h) a.sel[(int64_t)in[h] * a.N + n] = j;
      }
      break;
    }
    case CKB_STEP_TUCKER: {
      const int Kred = s.k_in * s.k_in;
      const int j = draw(s.cdf + ((int64_t)f * s.k_out + o) * Kred, Kred, u01(c[0]));
      if (a.mix) a.mix[row * a.N + n] = j;
      a.sel[(int64_t)in[0] * a.N + n] = j / s.k_in;
      a.sel[(int64_t)in[1] * a.N + n] = j % s.k_in;
      break;
    }
    case CKB_STEP_MIXING: {
      const int h = draw(s.cdf + ((int64_t)f * s.k_out + o) * H, H, u01(c[0]));
      if (a.mix) a.mix[row * a.N + n] = h;
      a.sel[(int64_t)in[h] * a.N + n] = o;
      break;
    }
    case CKB_STEP_HADAMARD:
      for (int h = 0; h < H; ++h) a.sel[(int64_t)in[h] * a.N + n] = o;
      break;
    case CKB_STEP_KRONECKER:
      a.sel[(int64_t)in[0] * a.N + n] = o / s.k_in;
      a.sel[(int64_t)in[1] * a.N + n] = o % s.k_in;
      break;
    case CKB_STEP_TABLE: {
      const int V = s.num_states;
      const int v = draw(s.cdf + ((int64_t)f * s.k_out + o) * V, V, u01(c[0]));
      const int64_t at = n * a.D + s.scope_var[f];
      if (a.x_is_float) ((float*)a.x)[at] = (float)v;
      else ((int64_t*)a.x)[at] = v;
      break;
    }
    case CKB_STEP_GAUSSIAN: {
      // Box-Muller on two uniforms of the block; u in (0, 1] keeps the log finite
      const float u1 = 1.0f - u01(c[0]), u2 = u01(c[1]);
      const float z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
      const int64_t p = (int64_t)f * s.k_out + o;
      const float val = s.p0[p] + s.p1[p] * z;
      const int64_t at = n * a.D + s.scope_var[f];
      if (a.x_is_float) ((float*)a.x)[at] = val;
      else ((int64_t*)a.x)[at] = (int64_t)val;
      break;
    }
    default:
      break;
  }
}

__global__ void fill_row_kernel(int32_t* row, int64_t N, int32_t value) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) row[n] = value;
}

}  // namespace
}  // namespace ckb

using namespace ckb;

extern "C" {

int ckb_sample_cdf_rows(const float* src, float* dst, int64_t rows, int32_t cols, int32_t mode,
                        int32_t units, void* stream) {
  if (!src || !dst || rows <= 0 || cols <= 0 || (mode != 0 && mode != 1) || (mode == 1 && units <= 0)) {
    set_error("ckb_sample_cdf_rows: bad arguments");
    return CKB_ERR_INVALID;
  }
  cdf_rows_kernel<<<ceil_div(rows, 128), 128, 0, (cudaStream_t)stream>>>(src, dst, rows, cols, mode, units);
  CKB_LAUNCH_CHECK();
  return CKB_OK;
}

int ckb_plan_sample(const ckb_sample_step_t* steps, int32_t n_steps, int64_t num_samples,
                    int64_t sample_base, uint64_t seed, int64_t num_rows, int32_t root_row, int32_t root_unit,
                    int32_t* sel, int32_t* mix, void* x, int32_t num_vars, int32_t x_is_float,
                    void* stream) {
  if (!steps || n_steps <= 0 || num_samples <= 0 || sample_base < 0 || num_rows <= 0 || !sel || !x || num_vars <= 0 ||
      root_row < 0 || root_row >= num_rows || root_unit < 0) {
    set_error("ckb_plan_sample: bad arguments");
    return CKB_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n_steps; ++i) {
    const ckb_sample_step_t& s = steps[i];
    const bool inner = s.kind == CKB_STEP_DENSE || s.kind == CKB_STEP_TUCKER || s.kind == CKB_STEP_MIXING ||
                       s.kind == CKB_STEP_HADAMARD || s.kind == CKB_STEP_KRONECKER;
    const bool draws = s.kind == CKB_STEP_DENSE || s.kind == CKB_STEP_TUCKER || s.kind == CKB_STEP_MIXING ||
                       s.kind == CKB_STEP_TABLE;
    if (!inner && s.kind != CKB_STEP_TABLE && s.kind != CKB_STEP_GAUSSIAN) {
      set_error("step %d: sampling is not supported for layers of kind %d", i, s.kind);
      return CKB_ERR_UNSUPPORTED;
    }
    if (s.num_folds <= 0 || s.k_out <= 0 || s.sel_row < 0 || s.sel_row + s.num_folds > num_rows ||
        (inner && (!s.in_sel_rows || s.arity <= 0)) || (draws && !s.cdf) ||
        (!inner && !s.scope_var) || (s.kind == CKB_STEP_GAUSSIAN && (!s.p0 || !s.p1)) ||
        (s.kind == CKB_STEP_TABLE && s.num_states <= 0) ||
        ((s.kind == CKB_STEP_TUCKER || s.kind == CKB_STEP_KRONECKER) && s.arity != 2)) {
      set_error("step %d: bad sampling descriptor", i);
      return CKB_ERR_INVALID;
    }
  }
  CKB_CUDA_CHECK(cudaMemsetAsync(sel, 0xFF, (size_t)num_rows * num_samples * 4, st));
  if (mix) CKB_CUDA_CHECK(cudaMemsetAsync(mix, 0xFF, (size_t)num_rows * num_samples * 4, st));
  const int blocks = ceil_div(num_samples, 256);
  fill_row_kernel<<<blocks, 256, 0, st>>>(sel + (int64_t)root_row * num_samples, num_samples, root_unit);
  CKB_LAUNCH_CHECK();
  SampleArgs a;
  a.N = num_samples;
  a.base = sample_base;
  a.sel = sel;
  a.mix = mix;
  a.x = x;
  a.D = num_vars;
  a.x_is_float = x_is_float;
  a.k0 = (uint32_t)seed;
  a.k1 = (uint32_t)(seed >> 32);
  for (int i = n_steps - 1; i >= 0; --i) {  // root first: consumers write their inputs' rows
    a.s = steps[i];
    dim3 grid(blocks, steps[i].num_folds);
    sample_step_kernel<<<grid, 256, 0, st>>>(a);
    CKB_LAUNCH_CHECK();
  }
  return CKB_OK;
}

}  // extern "C"
