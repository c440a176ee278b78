// Argument block shared by the SIMT (dense_kernels.cu) and tensor-core (dense_tc.cu) versions of
// the fused sum-product block.
#pragma once
#include "common.cuh"

namespace ckb {

struct DenseArgs {
  const float* W;          // (F, Ko, Kred)
  const int64_t* in_rows;  // (F*H) per-sample offsets, or nullptr: rows are x_base + f*B*Ki (H==1)
  const float* arena;      // base the in_rows offsets refer to (or x_base when in_rows == nullptr)
  float* y;                // (F, B, Ko) output block (forward: written; backward: read)
  int64_t B;
  int H, Ki, Ko, Kred, concat;
  // backward only
  GradSrc gs;
  float* gin;   // (F, gin_h, B, Ki)
  float* dWp;   // [splits][F][Ko][Kred] or nullptr
  int64_t chunk;
  int max_cons;  // largest number of consumer rows any fold sums (1 in a tree)
  int rows64;    // every gathered / consumer row offset is a multiple of 64 floats (CKB_STEP_ROWS64)
};

__device__ __forceinline__ const float* in_row(const DenseArgs& a, int f, int h) {
  return a.in_rows ? a.arena + a.B * a.in_rows[f * a.H + h] : a.arena + (int64_t)f * a.B * a.Ki;
}

// Writes u (pre-activation, log space) for one sample into `dst[0..Kred)`, returns the row max.
__device__ __forceinline__ float load_u(const DenseArgs& a, const float* const* rows, int64_t b,
                                        int lane, float* dst) {
  float m = -INFINITY;
  if (!a.concat) {
    for (int k = lane; k < a.Kred; k += 32) {
      float u = 0.f;
      for (int h = 0; h < a.H; ++h) u += rows[h][b * a.Ki + k];
      dst[k] = u;
      m = fmaxf(m, u);
    }
  } else {
    for (int h = 0; h < a.H; ++h)
      for (int k = lane; k < a.Ki; k += 32) {
        const float u = rows[h][b * a.Ki + k];
        dst[h * a.Ki + k] = u;
        m = fmaxf(m, u);
      }
  }
  return clamp_max(warp_max(m));
}

// Tensor-core (tcgen05) versions; return CKB_OK when they ran, 1 when the shape is not theirs.
int dense_tc_fwd(const DenseArgs& a, int F, Ctx& c);
int dense_tc_bwd(const DenseArgs& a, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes);
size_t dense_tc_bwd_ws(int F, int H, int Ko, int Kred, int64_t B);

// TMA-fed backward for Ki = Ko = 64 (dense_tc_bwd3.cu)
bool dense_tc_bwd3_ok(const DenseArgs& a, int rows64);
size_t dense_tc_bwd3_ws(int F, int64_t B);
int dense_tc_bwd3(const DenseArgs& a, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes);

// Ki = Ko = 128 on tcgen05 (dense128_tc.cu), bit 9 of CKB_OPT_TC_FAST_MATH (set in the default value)
int tc_flags();  // CKB_OPT_TC_FAST_MATH bits (dense_tc.cu)
bool dense128_tc_ok(const DenseArgs& a);
int dense128_tc_fwd(const DenseArgs& a, int F, Ctx& c);
size_t dense128_tc_bwd_ws(int F, int64_t B);
int dense128_tc_bwd(const DenseArgs& a, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes);

// Ki = Ko = 32 (dense32_kernels.cu)
bool dense32_ok(const DenseArgs& a);
int dense32_fwd(const DenseArgs& a, int F, Ctx& c);
int dense32_bwd(const DenseArgs& a, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes);
size_t dense32_bwd_ws(int F, int64_t B);

}  // namespace ckb
