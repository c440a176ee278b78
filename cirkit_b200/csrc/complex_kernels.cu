// 'complex-lse-sum' building blocks, FP32 SIMT (first correct version: one warp per (fold, sample)).
//
// Per-layer entry points (ckb_complex_*), validated on the B200 in round 2 against the oracle's
// complex path (oracle/reference_eval.py, fixtures tests/golden/*_complex*.npz):
// tests/test_gpu_zzz_complex_kernels.py.  The plan executor runs complex plans through its own
// step kernels (complex_plan.cu, deterministic weight gradients); these stay as unit-testable
// building blocks (their weight gradients use atomics).
//
// Reference semantics (cirkit/backend/torch):
//   * activations are complex logarithms, stored interleaved (re, im) as float2, layout
//     (fold, batch, unit) like the real-valued arena;
//   * ComplexLSESumSemiring.apply_reduce, semiring.py:440-476:
//       m = clamp(max_i Re u_i),  e_i = exp(u_i - m),  S_o = sum_i W[o,i] e_i,
//       y_o = csafelog(S_o) + m;
//   * csafelog, utils.py:32-50: forward log(z) = (log|z|, arg z); backward
//     nan_to_num(g / conj(z)) (NaN -> 0, +-inf -> +-FLT_MAX, per component);
//   * gradients follow PyTorch's convention for complex tensors (the stored gradient is
//     dL/d conj(z)): through a holomorphic f it is g_in = g_out * conj(f'(z)), through a product
//     z = w e it is g_w = g_z conj(e), g_e = g_z conj(w);
//   * TorchEmbeddingLayer.forward, layers/input.py:258-266: y[f,b,k] = csafelog(W[f,k,x[b,var_f]]).
// The shift m has a zero analytic derivative (d y / d m = 1 - S/S) and is treated as a constant,
// as in the real-valued kernels.
#include "common.cuh"

namespace ckb {
namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cexp_shifted(float2 u, float m) {
  float s, c;
  sincosf(u.y, &s, &c);
  const float r = expf(u.x - m);
  return make_float2(r * c, r * s);
}
__device__ __forceinline__ float2 clog(float2 z) {
  return make_float2(logf(hypotf(z.x, z.y)), atan2f(z.y, z.x));
}
__device__ __forceinline__ float nan_to_num(float v) {
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? FLT_MAX : -FLT_MAX;
  return v;
}
// nan_to_num(g / conj(z)): the quotient follows c10::complex's operator/ (scaled by the larger
// component of the divisor; a zero divisor yields component-wise g / 0).
__device__ __forceinline__ float2 safe_div_conj(float2 g, float2 z) {
  const float c = z.x, d = -z.y;
  const float ac = fabsf(c), ad = fabsf(d);
  float re, im;
  if (ac >= ad) {
    if (ac == 0.f && ad == 0.f) {
      re = g.x / ac;
      im = g.y / ad;
    } else {
      const float rat = d / c, scl = 1.f / (c + d * rat);
      re = (g.x + g.y * rat) * scl;
      im = (g.y - g.x * rat) * scl;
    }
  } else {
    const float rat = c / d, scl = 1.f / (c * rat + d);
    re = (g.x * rat + g.y) * scl;
    im = (g.y * rat - g.x) * scl;
  }
  return make_float2(nan_to_num(re), nan_to_num(im));
}

// u = x0 (+ x1) of one (fold, sample) row -> e (shared, per warp), returns the shift m
__device__ __forceinline__ float load_shifted_exp(const float2* x0, const float2* x1, int Ki,
                                                  int lane, float2* e) {
  float m = -INFINITY;
  for (int i = lane; i < Ki; i += 32) {
    float2 u = x0[i];
    if (x1) {
      const float2 v = x1[i];
      u.x += v.x;
      u.y += v.y;
    }
    e[i] = u;
    m = fmaxf(m, u.x);
  }
  m = clamp_max(warp_max(m));
  __syncwarp();
  for (int i = lane; i < Ki; i += 32) e[i] = cexp_shifted(e[i], m);
  __syncwarp();
  return m;
}

// ------------------------------------------------------------------------------------------
// CP-T / dense sum block: x0, x1 (F,B,Ki) [x1 may be null: arity 1], w (F,Ko,Ki), y (F,B,Ko)
// ------------------------------------------------------------------------------------------
__global__ void complex_cpt_fwd_kernel(const float2* __restrict__ x0, const float2* __restrict__ x1,
                                       const float2* __restrict__ w, float2* __restrict__ y,
                                       int64_t B, int Ki, int Ko) {
  extern __shared__ float2 smem_c[];
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float2* e = smem_c + warp * Ki;
  const float2* wf = w + (int64_t)f * Ko * Ki;
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < B; b += (int64_t)gridDim.x * nwarps) {
    const int64_t row = ((int64_t)f * B + b) * Ki;
    const float m = load_shifted_exp(x0 + row, x1 ? x1 + row : nullptr, Ki, lane, e);
    for (int o = lane; o < Ko; o += 32) {
      const float2* wr = wf + (int64_t)o * Ki;
      float2 s = make_float2(0.f, 0.f);
      for (int i = 0; i < Ki; ++i) {
        const float2 p = cmul(wr[i], e[i]);
        s.x += p.x;
        s.y += p.y;
      }
      float2 l = clog(s);
      l.x += m;
      y[((int64_t)f * B + b) * Ko + o] = l;
    }
    __syncwarp();
  }
}

// gy (F,B,Ko) -> gu (F,B,Ki) [gradient of u = x0 + x1: both inputs receive it], gw (F,Ko,Ki)
// accumulated with atomics (the caller zeroes it; gw may be null).
__global__ void complex_cpt_bwd_kernel(const float2* __restrict__ x0, const float2* __restrict__ x1,
                                       const float2* __restrict__ w, const float2* __restrict__ y,
                                       const float2* __restrict__ gy, float2* __restrict__ gu,
                                       float* __restrict__ gw, int64_t B, int Ki, int Ko) {
  extern __shared__ float2 smem_c[];
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float2* e = smem_c + warp * (Ki + Ko);
  float2* r = e + Ki;
  const float2* wf = w + (int64_t)f * Ko * Ki;
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < B; b += (int64_t)gridDim.x * nwarps) {
    const int64_t row = ((int64_t)f * B + b) * Ki;
    const float m = load_shifted_exp(x0 + row, x1 ? x1 + row : nullptr, Ki, lane, e);
    for (int o = lane; o < Ko; o += 32) {
      // S_o = exp(y_o - m): the sum the forward took the logarithm of
      const float2 s = cexp_shifted(y[((int64_t)f * B + b) * Ko + o], m);
      r[o] = safe_div_conj(gy[((int64_t)f * B + b) * Ko + o], s);
    }
    __syncwarp();
    for (int i = lane; i < Ki; i += 32) {
      float2 ge = make_float2(0.f, 0.f);
      for (int o = 0; o < Ko; ++o) {
        const float2 p = cmul_conj(r[o], wf[(int64_t)o * Ki + i]);
        ge.x += p.x;
        ge.y += p.y;
      }
      gu[row + i] = cmul_conj(ge, e[i]);
      if (gw) {
        for (int o = 0; o < Ko; ++o) {
          const float2 p = cmul_conj(r[o], e[i]);
          float* dst = gw + 2 * (((int64_t)f * Ko + o) * Ki + i);
          if (p.x != 0.f) atomicAdd(dst, p.x);
          if (p.y != 0.f) atomicAdd(dst + 1, p.y);
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Embedding: x (B, ld) int64 row-major, var (F,) int32, w (F,K,V), y (F,B,K)
// ------------------------------------------------------------------------------------------
__global__ void complex_embedding_fwd_kernel(const int64_t* __restrict__ x, int64_t ld,
                                             const int32_t* __restrict__ var,
                                             const float2* __restrict__ w, float2* __restrict__ y,
                                             int64_t B, int K, int V) {
  const int f = blockIdx.y;
  const int v_col = var[f];
  const int64_t total = B * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / K;
    const int k = (int)(idx - b * K);
    int64_t v = x[b * ld + v_col];
    v = v < 0 ? 0 : (v >= V ? V - 1 : v);
    y[(int64_t)f * total + idx] = clog(w[((int64_t)f * K + k) * V + v]);
  }
}

__global__ void complex_embedding_bwd_kernel(const int64_t* __restrict__ x, int64_t ld,
                                             const int32_t* __restrict__ var,
                                             const float2* __restrict__ w,
                                             const float2* __restrict__ gy, float* __restrict__ gw,
                                             int64_t B, int K, int V) {
  const int f = blockIdx.y;
  const int v_col = var[f];
  const int64_t total = B * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / K;
    const int k = (int)(idx - b * K);
    int64_t v = x[b * ld + v_col];
    v = v < 0 ? 0 : (v >= V ? V - 1 : v);
    const int64_t at = ((int64_t)f * K + k) * V + v;
    const float2 g = safe_div_conj(gy[(int64_t)f * total + idx], w[at]);
    if (g.x != 0.f) atomicAdd(gw + 2 * at, g.x);
    if (g.y != 0.f) atomicAdd(gw + 2 * at + 1, g.y);
  }
}

int check_shape(const char* what, int64_t F, int64_t B, int Ki, int Ko) {
  if (F <= 0 || B < 0 || Ki <= 0 || Ko <= 0 || F > 65535) {
    set_error("%s: bad shape F=%lld B=%lld Ki=%d Ko=%d", what, (long long)F, (long long)B, Ki, Ko);
    return CKB_ERR_INVALID;
  }
  return CKB_OK;
}

}  // namespace
}  // namespace ckb

using namespace ckb;

extern "C" {

int ckb_complex_cpt_fwd(const float* x0, const float* x1, const float* w, float* y, int32_t F,
                        int64_t B, int32_t Ki, int32_t Ko, void* stream) {
  if (!x0 || !w || !y) {
    set_error("ckb_complex_cpt_fwd: null pointer");
    return CKB_ERR_INVALID;
  }
  if (int rc = check_shape("ckb_complex_cpt_fwd", F, B, Ki, Ko)) return rc;
  if (B == 0) return CKB_OK;
  const size_t smem = (size_t)8 * Ki * sizeof(float2);
  if (smem > 48 * 1024) {
    set_error("ckb_complex_cpt_fwd: Ki = %d exceeds the per-warp row buffer", Ki);
    return CKB_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)min64(ceil_div(B, 8), 4 * kNumSMs), (unsigned)F);
  complex_cpt_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
      (const float2*)x0, (const float2*)x1, (const float2*)w, (float2*)y, B, Ki, Ko);
  CKB_LAUNCH_CHECK();
  return CKB_OK;
}

int ckb_complex_cpt_bwd(const float* x0, const float* x1, const float* w, const float* y,
                        const float* gy, float* gu, float* gw, int32_t F, int64_t B, int32_t Ki,
                        int32_t Ko, void* stream) {
  if (!x0 || !w || !y || !gy || !gu) {
    set_error("ckb_complex_cpt_bwd: null pointer");
    return CKB_ERR_INVALID;
  }
  if (int rc = check_shape("ckb_complex_cpt_bwd", F, B, Ki, Ko)) return rc;
  if (B == 0) return CKB_OK;
  const size_t smem = (size_t)8 * (Ki + Ko) * sizeof(float2);
  if (smem > 48 * 1024) {
    set_error("ckb_complex_cpt_bwd: Ki + Ko = %d exceeds the per-warp row buffers", Ki + Ko);
    return CKB_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)min64(ceil_div(B, 8), 4 * kNumSMs), (unsigned)F);
  complex_cpt_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
      (const float2*)x0, (const float2*)x1, (const float2*)w, (const float2*)y, (const float2*)gy,
      (float2*)gu, gw, B, Ki, Ko);
  CKB_LAUNCH_CHECK();
  return CKB_OK;
}

int ckb_complex_embedding_fwd(const int64_t* x, int64_t ld, const int32_t* var, const float* w,
                              float* y, int32_t F, int64_t B, int32_t K, int32_t V, void* stream) {
  if (!x || !var || !w || !y || V <= 0) {
    set_error("ckb_complex_embedding_fwd: bad arguments");
    return CKB_ERR_INVALID;
  }
  if (int rc = check_shape("ckb_complex_embedding_fwd", F, B, K, K)) return rc;
  if (B == 0) return CKB_OK;
  dim3 grid((unsigned)min64(ceil_div(B * K, 256), 4 * kNumSMs), (unsigned)F);
  complex_embedding_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      x, ld, var, (const float2*)w, (float2*)y, B, K, V);
  CKB_LAUNCH_CHECK();
  return CKB_OK;
}

int ckb_complex_embedding_bwd(const int64_t* x, int64_t ld, const int32_t* var, const float* w,
                              const float* gy, float* gw, int32_t F, int64_t B, int32_t K,
                              int32_t V, void* stream) {
  if (!x || !var || !w || !gy || !gw || V <= 0) {
    set_error("ckb_complex_embedding_bwd: bad arguments");
    return CKB_ERR_INVALID;
  }
  if (int rc = check_shape("ckb_complex_embedding_bwd", F, B, K, K)) return rc;
  if (B == 0) return CKB_OK;
  dim3 grid((unsigned)min64(ceil_div(B * K, 256), 4 * kNumSMs), (unsigned)F);
  complex_embedding_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      x, ld, var, (const float2*)w, (const float2*)gy, gw, B, K, V);
  CKB_LAUNCH_CHECK();
  return CKB_OK;
}

}  // extern "C"
