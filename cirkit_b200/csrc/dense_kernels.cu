// The fused sum-product block:  gather -> (Hadamard | concat) -> max -> exp -> K x K product
// against the fold's weights -> log -> + max, and its backward.  FP32 SIMT version.
//
// Reference path replaced (per folded layer, >= 6 HBM round trips of the (F,B,K) tensor):
//   circuits.py:42-47 (cat + index gather)  ->  layers/optimized.py:171-178 / inner.py:266-273
//   ->  semiring.py:382-408 (amax, clamp, sub, exp, einsum "fbi,foi->fbo", log, add).
// Here a CTA owns one fold and a tile of samples: the fold's (Ko,Kred) weight slice is staged in
// shared memory once, every warp turns SW samples into exp-shifted rows in shared memory and then
// runs a register-tiled matrix-vector product against the staged weights, so an activation row
// is read from HBM once and written once.
#include "dense.cuh"

namespace ckb {

constexpr int kMaxH = 64;  // inputs per fold the small kernels keep row pointers for

// Pre-activations of SW consecutive samples, lane owning the reduction indices lane + 32 j
// (-inf beyond Kred or beyond the batch).  All loads of one input row pointer are issued before
// any of them is consumed.
template <int SW, int NJ>
__device__ __forceinline__ void gather_u(const DenseArgs& a, const float* const* rows, int64_t b0,
                                         int lane, float (&uu)[SW][NJ]) {
  {  // summed inputs only (Hadamard-style); concatenating layers take the per-sample path
#pragma unroll
    for (int s = 0; s < SW; ++s)
#pragma unroll
      for (int j = 0; j < NJ; ++j) uu[s][j] = (b0 + s < a.B && lane + 32 * j < a.Kred) ? 0.f : -INFINITY;
    for (int h = 0; h < a.H; ++h) {
      const float* r = rows[h] + b0 * a.Ki + lane;
      float t[SW][NJ];
#pragma unroll
      for (int s = 0; s < SW; ++s)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          t[s][j] = (b0 + s < a.B && lane + 32 * j < a.Kred) ? __ldg(r + (int64_t)s * a.Ki + 32 * j) : 0.f;
#pragma unroll
      for (int s = 0; s < SW; ++s)
#pragma unroll
        for (int j = 0; j < NJ; ++j) uu[s][j] += t[s][j];
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward, Kred <= 128 and Ko <= 128 ("small" = the whole weight slice fits in shared memory)
// ------------------------------------------------------------------------------------------
template <int NO, int SW>
__global__ void __launch_bounds__(256) dense_fwd_small(DenseArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int KoP = 32 * NO + 1;
  const int KredP = (a.Kred + 3) & ~3;
  float* Wt = smem;                       // [KredP][KoP]  (transposed: lane o reads Wt[k][o])
  float* e_all = Wt + KredP * KoP;        // [8][SW][KredP]
  e_all = (float*)(((uintptr_t)e_all + 15) & ~(uintptr_t)15);
  __shared__ const float* rows[kMaxH];
  const int f = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < KredP * KoP; i += 256) Wt[i] = 0.f;
  if (tid < a.H) rows[tid] = in_row(a, f, tid);
  __syncthreads();
  const float* Wf = a.W + (int64_t)f * a.Ko * a.Kred;
  for (int i = tid; i < a.Ko * a.Kred; i += 256) {
    const int o = i / a.Kred, k = i - o * a.Kred;
    Wt[k * KoP + o] = Wf[i];
  }
  __syncthreads();

  float* e_w = e_all + warp * SW * KredP;
  for (int64_t b0 = ((int64_t)blockIdx.x * 8 + warp) * SW; b0 < a.B;
       b0 += (int64_t)gridDim.x * 8 * SW) {
    float m_reg[SW];
    if (a.concat || a.Kred > 32 * NO) {
      // concatenating sums, and reduction lengths beyond the register tile of the batched gather
      // below (the tile is sized by Ko: 32 * NO indices) -- rare and small: one sample at a time
#pragma unroll 1
      for (int s = 0; s < SW; ++s) {
        const int64_t b = b0 + s;
        float* es = e_w + s * KredP;
        float m = 0.f;
        if (b < a.B) m = load_u(a, rows, b, lane, es);
        __syncwarp();
        for (int k = lane; k < KredP; k += 32) es[k] = (b < a.B && k < a.Kred) ? expf(es[k] - m) : 0.f;
        m_reg[s] = m;
      }
    } else {
      // the loads of all SW samples go out together (lane owns reduction indices lane + 32 j):
      // one memory round trip per input instead of one per sample
      float uu[SW][NO];
      gather_u<SW, NO>(a, rows, b0, lane, uu);
#pragma unroll
      for (int s = 0; s < SW; ++s) {
        const bool valid = b0 + s < a.B;
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < NO; ++j) m = fmaxf(m, uu[s][j]);
        m = valid ? clamp_max(warp_max(m)) : 0.f;
        float* es = e_w + s * KredP;
#pragma unroll
        for (int j = 0; j < NO; ++j) {
          const int k = lane + 32 * j;
          if (k < KredP) es[k] = (valid && k < a.Kred) ? expf(uu[s][j] - m) : 0.f;
        }
        m_reg[s] = m;
      }
    }
    __syncwarp();
    float acc[SW][NO];
#pragma unroll
    for (int s = 0; s < SW; ++s)
#pragma unroll
      for (int j = 0; j < NO; ++j) acc[s][j] = 0.f;
    for (int k = 0; k < KredP; k += 4) {
      float w[4][NO];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int j = 0; j < NO; ++j) w[kk][j] = Wt[(k + kk) * KoP + lane + 32 * j];
#pragma unroll
      for (int s = 0; s < SW; ++s) {
        const float4 e4 = *reinterpret_cast<const float4*>(e_w + s * KredP + k);
#pragma unroll
        for (int j = 0; j < NO; ++j) {
          acc[s][j] = fmaf(e4.x, w[0][j], acc[s][j]);
          acc[s][j] = fmaf(e4.y, w[1][j], acc[s][j]);
          acc[s][j] = fmaf(e4.z, w[2][j], acc[s][j]);
          acc[s][j] = fmaf(e4.w, w[3][j], acc[s][j]);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < SW; ++s) {
      const int64_t b = b0 + s;
      if (b < a.B) {
#pragma unroll
        for (int j = 0; j < NO; ++j) {
          const int o = lane + 32 * j;
          if (o < a.Ko) a.y[((int64_t)f * a.B + b) * a.Ko + o] = logf(acc[s][j]) + m_reg[s];
        }
      }
    }
    __syncwarp();
  }
}

template <int NO, int SW>
static int launch_dense_fwd_small(const DenseArgs& a, int F, Ctx& c) {
  const int KredP = (a.Kred + 3) & ~3;
  const size_t smem = ((size_t)KredP * (32 * NO + 1) + 8 * SW * KredP) * 4 + 16;
  static PerDeviceOnce attr;
  if (attr.first()) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_fwd_small<NO, SW>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  dim3 grid(max(1, ceil_div(a.B, 8 * SW)), F);
  dense_fwd_small<NO, SW><<<grid, 256, smem, c.stream>>>(a);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// ------------------------------------------------------------------------------------------
// forward, any shape: one warp per (fold, sample), weights read through L1/L2.
// ------------------------------------------------------------------------------------------
__global__ void dense_fwd_generic(DenseArgs a, int e_stride) {
  extern __shared__ __align__(16) float smem[];
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* e = smem + warp * e_stride;
  const float* Wf = a.W + (int64_t)f * a.Ko * a.Kred;
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < a.B; b += (int64_t)gridDim.x * nwarps) {
    float m = -INFINITY;
    for (int h = 0; h < a.H; ++h) {
      const float* r = in_row(a, f, h) + b * a.Ki;
      for (int k = lane; k < a.Ki; k += 32) {
        const int kk = a.concat ? h * a.Ki + k : k;
        const float u = (a.concat || h == 0) ? r[k] : e[kk] + r[k];
        e[kk] = u;
      }
    }
    __syncwarp();
    for (int k = lane; k < a.Kred; k += 32) m = fmaxf(m, e[k]);
    m = clamp_max(warp_max(m));
    for (int k = lane; k < a.Kred; k += 32) e[k] = expf(e[k] - m);
    __syncwarp();
    for (int o = lane; o < a.Ko; o += 32) {
      const float* wr = Wf + (int64_t)o * a.Kred;
      float s = 0.f;
      for (int k = 0; k < a.Kred; ++k) s = fmaf(e[k], wr[k], s);
      a.y[((int64_t)f * a.B + b) * a.Ko + o] = logf(s) + m;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Any shape, tiled (sum layers over concatenated inputs with Kred > 128: PoonDomingos at K = 128
// has Kred up to 1792).  The warp-per-sample kernel above streams the whole weight matrix of a
// fold (Ko * Kred * 4 bytes, ~1 MB) through L2 once PER SAMPLE; here a CTA owns a 64-sample x
// 64-output tile and walks the reduction in steps of 32 through shared memory, so weights and
// inputs are read once per tile (2.2 ms -> 0.2 ms for the F = 4, Kred = 1792 layer of config 4).
//   pass 1  row shifts m[f,b] = max_i u[b,i]                       (dense_rowmax_generic)
//   pass 2  y[b,o] = log sum_i W[o,i] exp(u[b,i] - m[b]) + m[b]     (dense_fwd_tiled_generic)
// and for the backward, with r = g exp(m - y):
//   du[b,i] = exp(u[b,i] - m[b]) * sum_o r[b,o] W[o,i]              (dense_du_tiled_generic)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float gather_u(const DenseArgs& a, int f, int64_t b, int i) {
  if (a.concat) {
    const int h = i / a.Ki;
    return in_row(a, f, h)[b * a.Ki + (i - h * a.Ki)];
  }
  float u = 0.f;
  for (int h = 0; h < a.H; ++h) u += in_row(a, f, h)[b * a.Ki + i];
  return u;
}

__global__ void dense_rowmax_generic(DenseArgs a, float* __restrict__ mrow) {
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < a.B; b += (int64_t)gridDim.x * nwarps) {
    float m = -INFINITY;
    for (int i = lane; i < a.Kred; i += 32) m = fmaxf(m, gather_u(a, f, b, i));
    m = clamp_max(warp_max(m));
    if (lane == 0) mrow[(int64_t)f * a.B + b] = m;
  }
}

constexpr int kGtB = 64, kGtN = 64, kGtK = 32, kGtLd = 68;  // tile sizes; padded row stride
// thread (ty, tx) of 16 x 16 owns samples 4 ty .. 4 ty + 3 and columns 4 tx .. 4 tx + 3 of the tile
__global__ void __launch_bounds__(256) dense_fwd_tiled_generic(DenseArgs a, const float* __restrict__ mrow) {
  __shared__ __align__(16) float es[kGtK][kGtLd];  // [reduction index][sample]
  __shared__ __align__(16) float ws[kGtK][kGtLd];  // [reduction index][output]
  const int f = blockIdx.z;
  const int64_t b0 = (int64_t)blockIdx.x * kGtB;
  const int o0 = blockIdx.y * kGtN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const float* Wf = a.W + (int64_t)f * a.Ko * a.Kred;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a.Kred; k0 += kGtK) {
    for (int q = tid; q < kGtB * kGtK; q += 256) {
      const int bb = q / kGtK, k = q - bb * kGtK;  // consecutive threads: consecutive reduction indices
      const int64_t b = b0 + bb;
      float e = 0.f;
      if (b < a.B && k0 + k < a.Kred) e = expf(gather_u(a, f, b, k0 + k) - mrow[(int64_t)f * a.B + b]);
      es[k][bb] = e;
    }
    for (int q = tid; q < kGtN * kGtK; q += 256) {
      const int oo = q / kGtK, k = q - oo * kGtK;
      ws[k][oo] = (o0 + oo < a.Ko && k0 + k < a.Kred) ? __ldg(Wf + (int64_t)(o0 + oo) * a.Kred + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kGtK; ++k) {
      const float4 ev = *reinterpret_cast<const float4*>(&es[k][4 * ty]);
      const float4 wv = *reinterpret_cast<const float4*>(&ws[k][4 * tx]);
      const float e4[4] = {ev.x, ev.y, ev.z, ev.w}, w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(e4[p], w4[q], acc[p][q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int64_t b = b0 + 4 * ty + p;
    if (b >= a.B) continue;
    const float m = mrow[(int64_t)f * a.B + b];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int o = o0 + 4 * tx + q;
      if (o < a.Ko) a.y[((int64_t)f * a.B + b) * a.Ko + o] = logf(acc[p][q]) + m;
    }
  }
}

__global__ void __launch_bounds__(256) dense_du_tiled_generic(DenseArgs a, const float* __restrict__ mrow) {
  __shared__ __align__(16) float rs[kGtK][kGtLd];  // [output][sample]
  __shared__ __align__(16) float ws[kGtK][kGtLd];  // [output][reduction index]
  const int f = blockIdx.z;
  const int64_t b0 = (int64_t)blockIdx.x * kGtB;
  const int i0 = blockIdx.y * kGtN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const float* Wf = a.W + (int64_t)f * a.Ko * a.Kred;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a.Ko; k0 += kGtK) {
    for (int q = tid; q < kGtB * kGtK; q += 256) {
      const int bb = q / kGtK, k = q - bb * kGtK;  // consecutive threads: consecutive outputs of a sample
      const int64_t b = b0 + bb;
      float r = 0.f;
      if (b < a.B && k0 + k < a.Ko) {
        const float g = pull_grad(a.gs, f, b, a.Ko, k0 + k);
        if (g != 0.f) r = g * expf(mrow[(int64_t)f * a.B + b] - a.y[((int64_t)f * a.B + b) * a.Ko + k0 + k]);
      }
      rs[k][bb] = r;
    }
    for (int q = tid; q < kGtK * kGtN; q += 256) {
      const int k = q / kGtN, ii = q - k * kGtN;  // consecutive threads: consecutive reduction indices
      ws[k][ii] = (k0 + k < a.Ko && i0 + ii < a.Kred) ? __ldg(Wf + (int64_t)(k0 + k) * a.Kred + i0 + ii) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kGtK; ++k) {
      const float4 rv = *reinterpret_cast<const float4*>(&rs[k][4 * ty]);
      const float4 wv = *reinterpret_cast<const float4*>(&ws[k][4 * tx]);
      const float r4[4] = {rv.x, rv.y, rv.z, rv.w}, w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(r4[p], w4[q], acc[p][q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int64_t b = b0 + 4 * ty + p;
    if (b >= a.B) continue;
    const float m = mrow[(int64_t)f * a.B + b];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + 4 * tx + q;
      if (i >= a.Kred) continue;
      const float du = expf(gather_u(a, f, b, i) - m) * acc[p][q];
      if (!a.concat) {
        a.gin[((int64_t)f * a.B + b) * a.Ki + i] = du;
      } else {
        const int h = i / a.Ki;
        a.gin[(((int64_t)f * a.H + h) * a.B + b) * a.Ki + (i - h * a.Ki)] = du;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Single-output layers (Ko = 1: the root sum of a circuit).  One warp per sample, the weight
// row in registers; the batch reduction of dW goes through per-block partials.
// ------------------------------------------------------------------------------------------
constexpr int kKo1MaxPerLane = 4;  // Kred <= 128

__device__ __forceinline__ float ko1_load(const DenseArgs& a, int f, int64_t b, int lane, float* u) {
  float m = -INFINITY;
#pragma unroll
  for (int t = 0; t < kKo1MaxPerLane; ++t) {
    const int k = lane + 32 * t;
    float v = -INFINITY;
    if (k < a.Kred) {
      if (!a.concat) {
        v = 0.f;
        for (int h = 0; h < a.H; ++h) v += in_row(a, f, h)[b * a.Ki + k];
      } else {
        const int h = k / a.Ki;
        v = in_row(a, f, h)[b * a.Ki + (k - h * a.Ki)];
      }
    }
    u[t] = v;
    m = fmaxf(m, v);
  }
  return clamp_max(warp_max(m));
}

__global__ void dense_ko1_fwd_kernel(DenseArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float w[kKo1MaxPerLane];
#pragma unroll
  for (int t = 0; t < kKo1MaxPerLane; ++t)
    w[t] = lane + 32 * t < a.Kred ? a.W[(int64_t)f * a.Kred + lane + 32 * t] : 0.f;
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < a.B; b += (int64_t)gridDim.x * nwarps) {
    float u[kKo1MaxPerLane];
    const float m = ko1_load(a, f, b, lane, u);
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < kKo1MaxPerLane; ++t)
      if (lane + 32 * t < a.Kred) s = fmaf(w[t], expf(u[t] - m), s);
    s = warp_sum(s);
    if (lane == 0) a.y[(int64_t)f * a.B + b] = logf(s) + m;
  }
}

__global__ void dense_ko1_bwd_kernel(DenseArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[8][32 * kKo1MaxPerLane];
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float w[kKo1MaxPerLane], dw[kKo1MaxPerLane];
#pragma unroll
  for (int t = 0; t < kKo1MaxPerLane; ++t) {
    w[t] = lane + 32 * t < a.Kred ? a.W[(int64_t)f * a.Kred + lane + 32 * t] : 0.f;
    dw[t] = 0.f;
  }
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < a.B; b += (int64_t)gridDim.x * nwarps) {
    float u[kKo1MaxPerLane];
    const float m = ko1_load(a, f, b, lane, u);
    const float g = pull_grad(a.gs, f, b, 1, 0);
    const float r = (g == 0.f) ? 0.f : g * expf(m - a.y[(int64_t)f * a.B + b]);
#pragma unroll
    for (int t = 0; t < kKo1MaxPerLane; ++t) {
      const int i = lane + 32 * t;
      if (i < a.Kred) {
        const float e = expf(u[t] - m);
        dw[t] = fmaf(r, e, dw[t]);
        const float du = e * r * w[t];
        if (!a.concat) {
          a.gin[((int64_t)f * a.B + b) * a.Ki + i] = du;
        } else {
          const int h = i / a.Ki;
          a.gin[(((int64_t)f * a.H + h) * a.B + b) * a.Ki + (i - h * a.Ki)] = du;
        }
      }
    }
  }
  if (a.dWp == nullptr) return;
#pragma unroll
  for (int t = 0; t < kKo1MaxPerLane; ++t) red[warp][lane + 32 * t] = dw[t];
  __syncthreads();
  for (int i = threadIdx.x; i < a.Kred; i += blockDim.x) {
    float s2 = 0.f;
    for (int wv = 0; wv < nwarps; ++wv) s2 += red[wv][i];
    a.dWp[((int64_t)blockIdx.x * gridDim.y + f) * a.Kred + i] = s2;
  }
}

static bool dense_ko1_ok(const DenseArgs& a) { return a.Ko == 1 && a.Kred <= 32 * kKo1MaxPerLane; }
static int dense_ko1_blocks(int F, int64_t B) {
  return (int)max64(1, min64(ceil_div(B, 8), ceil_div(4 * kNumSMs, F)));
}

static int run_dense_fwd(const DenseArgs& a, int F, Ctx& c) {
  {
    const int rc = dense_tc_fwd(a, F, c);  // tensor-core path for the hot shape
    if (rc <= 0) return rc;
  }
  if (dense32_ok(a)) return dense32_fwd(a, F, c);
  if (dense128_tc_ok(a)) return dense128_tc_fwd(a, F, c);  // Ki = Ko = 128 on tcgen05
  if (dense_ko1_ok(a)) {
    dim3 grid(dense_ko1_blocks(F, a.B), F);
    CKB_CUDA_CHECK(launch_pdl(dense_ko1_fwd_kernel, grid, dim3(256), 0, c.stream, a));
    c.launches++;
    return CKB_OK;
  }
  if (a.Kred <= 128 && a.Ko <= 128 && a.H <= kMaxH) {
    const int kmax = max(a.Ko, 1);
    if (kmax <= 32) return launch_dense_fwd_small<1, 16>(a, F, c);
    if (kmax <= 64) return launch_dense_fwd_small<2, 16>(a, F, c);
    return launch_dense_fwd_small<4, 8>(a, F, c);
  }
  if (c.ws != nullptr && c.ws_bytes >= (size_t)F * a.B * 4) {  // tiled: row shifts, then 64 x 64 tiles
    float* mrow = (float*)c.ws;
    dim3 g1((int)min64(ceil_div(a.B, 8), 8 * kNumSMs), F);
    dense_rowmax_generic<<<g1, 256, 0, c.stream>>>(a, mrow);
    CKB_LAUNCH_CHECK();
    dim3 g2(ceil_div(a.B, kGtB), ceil_div(a.Ko, kGtN), F);
    dense_fwd_tiled_generic<<<g2, 256, 0, c.stream>>>(a, mrow);
    CKB_LAUNCH_CHECK();
    c.launches += 2;
    return CKB_OK;
  }
  const int e_stride = (a.Kred + 3) & ~3;
  int nwarps = 8;
  while ((size_t)nwarps * e_stride * 4 > 96 * 1024 && nwarps > 1) nwarps >>= 1;
  const size_t smem = (size_t)nwarps * e_stride * 4;
  if (smem > 200 * 1024) {
    set_error("dense_fwd: reduction length %d too large", a.Kred);
    return CKB_ERR_UNSUPPORTED;
  }
  static PerDeviceOnce attr;
  if (attr.first()) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_fwd_generic,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  dim3 grid((int)min64(ceil_div(a.B, nwarps), 8 * kNumSMs), F);
  dense_fwd_generic<<<grid, nwarps * 32, smem, c.stream>>>(a, e_stride);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// ------------------------------------------------------------------------------------------
// backward, small shapes.  With r[o] = g[o] / S[o] (S[o] = exp(y[o] - m), so the forward
// product is not recomputed) and e = exp(u - m):
//   d/du[i]   = e[i] * sum_o r[o] W[o,i]          (phase B, per warp)
//   d/dW[o,i] = sum_b r[b,o] e[b,i]               (phase C, per CTA, registers across tiles)
// ------------------------------------------------------------------------------------------
template <int NI, int TT, int SW>
__global__ void __launch_bounds__(256) dense_bwd_small(DenseArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int KS = 32 * NI + 1;  // row stride of the staged weights
  constexpr int AL = TT > 4 ? TT : 4;  // row strides keep float4 / TT-wide reads in bounds
  const int LR = max(4, (a.Ko + AL - 1) / AL * AL);
  const int LE = (a.Kred + AL - 1) / AL * AL;
  float* Wn = smem;                  // [LR][KS]   (natural: lane i reads Wn[o][i])
  float* r_all = Wn + LR * KS;       // [8*SW][LR]
  r_all = (float*)(((uintptr_t)r_all + 15) & ~(uintptr_t)15);
  float* e_all = r_all + 8 * SW * LR;  // [8*SW][LE]
  __shared__ const float* rows[kMaxH];
  const int f = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < LR * KS; i += 256) Wn[i] = 0.f;
  if (tid < a.H) rows[tid] = in_row(a, f, tid);
  __syncthreads();
  const float* Wf = a.W + (int64_t)f * a.Ko * a.Kred;
  for (int i = tid; i < a.Ko * a.Kred; i += 256) {
    const int o = i / a.Kred, k = i - o * a.Kred;
    Wn[o * KS + k] = Wf[i];
  }
  __syncthreads();

  // phase C mapping: thread owns the (TT x TT) block of dW at (o0, i0)
  const int n_ti = (a.Kred + TT - 1) / TT;
  const int n_to = (a.Ko + TT - 1) / TT;
  const int ti = tid % n_ti, to = tid / n_ti;
  const bool c_active = a.dWp != nullptr && to < n_to;
  const int o0 = to * TT, i0 = ti * TT;
  float dw[TT][TT];
#pragma unroll
  for (int p = 0; p < TT; ++p)
#pragma unroll
    for (int q = 0; q < TT; ++q) dw[p][q] = 0.f;

  float* r_w = r_all + warp * SW * LR;
  float* e_w = e_all + warp * SW * LE;
  // the usual case (a tree): one consumer row, resolved once instead of per element
  const float* grow = nullptr;
  if (a.gs.cons_ptr == nullptr) grow = a.gs.garena + (int64_t)f * a.gs.B * a.Ko;
  else if (a.gs.cons_ptr[f + 1] - a.gs.cons_ptr[f] == 1) grow = a.gs.garena + a.gs.B * a.gs.cons_rows[a.gs.cons_ptr[f]];
  const int64_t b_begin = (int64_t)blockIdx.x * a.chunk;
  const int64_t b_end = min(a.B, b_begin + a.chunk);
  for (int64_t t0 = b_begin; t0 < b_end; t0 += 8 * SW) {
    const int64_t b0 = t0 + warp * SW;
    // ---- phase A: e and r of this warp's samples
    static_assert(SW % 4 == 0, "sub-batches of 4 samples");
    if (a.concat) {
      // concatenating sums (rare, small): one sample at a time
#pragma unroll 1
      for (int s = 0; s < SW; ++s) {
        const int64_t b = b0 + s;
        float* es = e_w + s * LE;
        float* rs = r_w + s * LR;
        const bool valid = b < b_end;
        float m = 0.f;
        if (valid) m = load_u(a, rows, b, lane, es);
        __syncwarp();
        for (int k = lane; k < LE; k += 32) es[k] = (valid && k < a.Kred) ? expf(es[k] - m) : 0.f;
        for (int o = lane; o < LR; o += 32) {
          float r = 0.f;
          if (valid && o < a.Ko) {
            const float g = pull_grad(a.gs, f, b, a.Ko, o);
            r = (g == 0.f) ? 0.f : g * expf(m - a.y[((int64_t)f * a.B + b) * a.Ko + o]);
          }
          rs[o] = r;
        }
      }
    } else {
      // four samples at a time with all their loads (inputs, y, g) in flight together
#pragma unroll 1
      for (int s0 = 0; s0 < SW; s0 += 4) {
        float uu[4][NI], yv[4][NI], gv[4][NI];
        gather_u<4, NI>(a, rows, b0 + s0, lane, uu);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int64_t b = b0 + s0 + s;
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const int o = lane + 32 * j;
            const bool ok = b < b_end && o < a.Ko;
            yv[s][j] = ok ? __ldg(a.y + ((int64_t)f * a.B + b) * a.Ko + o) : 0.f;
            gv[s][j] = !ok ? 0.f : (grow ? __ldg(grow + b * a.Ko + o) : pull_grad(a.gs, f, b, a.Ko, o));
          }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const bool valid = b0 + s0 + s < b_end;
          float* es = e_w + (s0 + s) * LE;
          float* rs = r_w + (s0 + s) * LR;
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < NI; ++j) m = fmaxf(m, uu[s][j]);
          m = valid ? clamp_max(warp_max(m)) : 0.f;
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const int k = lane + 32 * j;
            if (k < LE) es[k] = (valid && k < a.Kred) ? expf(uu[s][j] - m) : 0.f;
            if (k < LR) rs[k] = (valid && k < a.Ko && gv[s][j] != 0.f) ? gv[s][j] * expf(m - yv[s][j]) : 0.f;
          }
        }
      }
    }
    __syncwarp();
    // ---- phase B: t[i] = sum_o r[o] W[o,i];  du[i] = e[i] t[i]
    {
      float acc[SW][NI];
#pragma unroll
      for (int s = 0; s < SW; ++s)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[s][j] = 0.f;
      for (int o = 0; o < LR; o += 4) {
        float w[4][NI];
#pragma unroll
        for (int oo = 0; oo < 4; ++oo)
#pragma unroll
          for (int j = 0; j < NI; ++j) w[oo][j] = Wn[(o + oo) * KS + lane + 32 * j];
#pragma unroll
        for (int s = 0; s < SW; ++s) {
          const float4 r4 = *reinterpret_cast<const float4*>(r_w + s * LR + o);
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            acc[s][j] = fmaf(r4.x, w[0][j], acc[s][j]);
            acc[s][j] = fmaf(r4.y, w[1][j], acc[s][j]);
            acc[s][j] = fmaf(r4.z, w[2][j], acc[s][j]);
            acc[s][j] = fmaf(r4.w, w[3][j], acc[s][j]);
          }
        }
      }
#pragma unroll
      for (int s = 0; s < SW; ++s) {
        const int64_t b = b0 + s;
        if (b < b_end) {
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const int i = lane + 32 * j;
            if (i < a.Kred) {
              const float du = e_w[s * LE + i] * acc[s][j];
              if (!a.concat) {
                a.gin[((int64_t)f * a.B + b) * a.Ki + i] = du;
              } else {
                const int h = i / a.Ki;
                a.gin[(((int64_t)f * a.H + h) * a.B + b) * a.Ki + (i - h * a.Ki)] = du;
              }
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- phase C: dW += r^T e over the 8*SW samples of the tile
    if (c_active) {
      const int n_s = (int)min64(8 * SW, b_end - t0);
      for (int s = 0; s < n_s; ++s) {
        float rv[TT], ev[TT];
        const float* rp = r_all + s * LR + o0;
        const float* ep = e_all + s * LE + i0;
        if constexpr (TT >= 4) {
#pragma unroll
          for (int p = 0; p < TT; p += 4) {
            const float4 r4 = *reinterpret_cast<const float4*>(rp + p);
            const float4 e4 = *reinterpret_cast<const float4*>(ep + p);
            rv[p] = r4.x; rv[p + 1] = r4.y; rv[p + 2] = r4.z; rv[p + 3] = r4.w;
            ev[p] = e4.x; ev[p + 1] = e4.y; ev[p + 2] = e4.z; ev[p + 3] = e4.w;
          }
        } else {
#pragma unroll
          for (int p = 0; p < TT; ++p) { rv[p] = rp[p]; ev[p] = ep[p]; }
        }
#pragma unroll
        for (int p = 0; p < TT; ++p)
#pragma unroll
          for (int q = 0; q < TT; ++q) dw[p][q] = fmaf(rv[p], ev[q], dw[p][q]);
      }
    }
    __syncthreads();
  }
  if (c_active) {
    float* out = a.dWp + ((int64_t)blockIdx.x * gridDim.y + f) * a.Ko * a.Kred;
#pragma unroll
    for (int p = 0; p < TT; ++p)
#pragma unroll
      for (int q = 0; q < TT; ++q)
        if (o0 + p < a.Ko && i0 + q < a.Kred) out[(int64_t)(o0 + p) * a.Kred + i0 + q] = dw[p][q];
  }
}

static void dense_bwd_config(int F, int Ko, int Kred, int64_t B, int& SW, int& splits, int64_t& chunk) {
  SW = max(Ko, Kred) <= 64 ? 16 : 8;
  const int tile = 8 * SW;
  const int64_t want = ceil_div(3 * kNumSMs, F);
  splits = (int)max64(1, min64(want, ceil_div(B, tile)));
  chunk = (int64_t)ceil_div(ceil_div(B, splits), tile) * tile;
  splits = ceil_div(B, chunk);
}

static bool dense_small_ok(int H, int Ko, int Kred) { return Kred <= 128 && Ko <= 128 && H <= kMaxH; }

template <int NI, int TT, int SW>
static int launch_dense_bwd_small(const DenseArgs& a, int F, int splits, Ctx& c) {
  constexpr int KS = 32 * NI + 1;
  constexpr int AL = TT > 4 ? TT : 4;
  const int LR = max(4, (a.Ko + AL - 1) / AL * AL);
  const int LE = (a.Kred + AL - 1) / AL * AL;
  const size_t smem = ((size_t)LR * KS + 8 * SW * (LR + LE)) * 4 + 16;
  static PerDeviceOnce attr;
  if (attr.first()) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_bwd_small<NI, TT, SW>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  dim3 grid(splits, F);
  dense_bwd_small<NI, TT, SW><<<grid, 256, smem, c.stream>>>(a);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// backward, any shape, part 1: one warp per (fold, sample) writes du and keeps the row shift m[f,b]
// for part 2.
__global__ void dense_bwd_generic(DenseArgs a, int e_stride, int r_stride, float* mrow) {
  extern __shared__ __align__(16) float smem[];
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* e = smem + warp * (e_stride + r_stride);
  float* r = e + e_stride;
  const float* Wf = a.W + (int64_t)f * a.Ko * a.Kred;
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < a.B; b += (int64_t)gridDim.x * nwarps) {
    for (int h = 0; h < a.H; ++h) {
      const float* xr = in_row(a, f, h) + b * a.Ki;
      for (int k = lane; k < a.Ki; k += 32) {
        const int kk = a.concat ? h * a.Ki + k : k;
        e[kk] = (a.concat || h == 0) ? xr[k] : e[kk] + xr[k];
      }
    }
    __syncwarp();
    float m = -INFINITY;
    for (int k = lane; k < a.Kred; k += 32) m = fmaxf(m, e[k]);
    m = clamp_max(warp_max(m));
    if (mrow && lane == 0) mrow[(int64_t)f * a.B + b] = m;
    for (int k = lane; k < a.Kred; k += 32) e[k] = expf(e[k] - m);
    for (int o = lane; o < a.Ko; o += 32) {
      const float g = pull_grad(a.gs, f, b, a.Ko, o);
      r[o] = (g == 0.f) ? 0.f : g * expf(m - a.y[((int64_t)f * a.B + b) * a.Ko + o]);
    }
    __syncwarp();
    for (int i = lane; i < a.Kred; i += 32) {
      float t = 0.f;
      for (int o = 0; o < a.Ko; ++o) t = fmaf(r[o], Wf[(int64_t)o * a.Kred + i], t);
      const float du = e[i] * t;
      if (!a.concat) {
        a.gin[((int64_t)f * a.B + b) * a.Ki + i] = du;
      } else {
        const int h = i / a.Ki;
        a.gin[(((int64_t)f * a.H + h) * a.B + b) * a.Ki + (i - h * a.Ki)] = du;
      }
    }
    __syncwarp();
  }
}

// part 2: dW[f,o,i] = sum_b r[b,o] e[b,i] with r = g exp(m - y), e = exp(u - m) rebuilt from the
// stored row shifts.  A CTA owns a 32 x 64 tile of one fold's dW and walks the batch in order, a
// thread owns 2 x 4 of its entries: no atomics, the sum over samples has one fixed order.
constexpr int kDwTo = 32, kDwTi = 64, kDwTb = 32;
__global__ void __launch_bounds__(256) dense_dw_generic(DenseArgs a, const float* __restrict__ mrow,
                                                        float* __restrict__ dW) {
  __shared__ float rs[kDwTb][kDwTo + 1];
  __shared__ __align__(16) float es[kDwTb][kDwTi + 4];
  const int f = blockIdx.z, o0 = blockIdx.y * kDwTo, i0 = blockIdx.x * kDwTi;
  const int tid = threadIdx.x;
  const int to = (tid >> 4) * 2, ti = (tid & 15) * 4;  // 16 x 16 threads -> 2 x 4 entries each
  float acc[2][4] = {};
  for (int64_t b0 = 0; b0 < a.B; b0 += kDwTb) {
    for (int q = tid; q < kDwTb * kDwTo; q += 256) {
      const int bb = q / kDwTo, o = o0 + q % kDwTo;
      const int64_t b = b0 + bb;
      float r = 0.f;
      if (b < a.B && o < a.Ko) {
        const float g = pull_grad(a.gs, f, b, a.Ko, o);
        if (g != 0.f) r = g * expf(mrow[(int64_t)f * a.B + b] - a.y[((int64_t)f * a.B + b) * a.Ko + o]);
      }
      rs[bb][q % kDwTo] = r;
    }
    for (int q = tid; q < kDwTb * kDwTi; q += 256) {
      const int bb = q / kDwTi, i = i0 + q % kDwTi;
      const int64_t b = b0 + bb;
      float e = 0.f;
      if (b < a.B && i < a.Kred) {
        float u;
        if (a.concat) {
          const int h = i / a.Ki;
          u = in_row(a, f, h)[b * a.Ki + (i - h * a.Ki)];
        } else {
          u = 0.f;
          for (int h = 0; h < a.H; ++h) u += in_row(a, f, h)[b * a.Ki + i];
        }
        e = expf(u - mrow[(int64_t)f * a.B + b]);
      }
      es[bb][q % kDwTi] = e;
    }
    __syncthreads();
#pragma unroll 8
    for (int bb = 0; bb < kDwTb; ++bb) {
      const float r0 = rs[bb][to], r1 = rs[bb][to + 1];
      const float4 ev = *reinterpret_cast<const float4*>(&es[bb][ti]);
      acc[0][0] = fmaf(r0, ev.x, acc[0][0]);
      acc[0][1] = fmaf(r0, ev.y, acc[0][1]);
      acc[0][2] = fmaf(r0, ev.z, acc[0][2]);
      acc[0][3] = fmaf(r0, ev.w, acc[0][3]);
      acc[1][0] = fmaf(r1, ev.x, acc[1][0]);
      acc[1][1] = fmaf(r1, ev.y, acc[1][1]);
      acc[1][2] = fmaf(r1, ev.z, acc[1][2]);
      acc[1][3] = fmaf(r1, ev.w, acc[1][3]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (o0 + to + p < a.Ko && i0 + ti + q < a.Kred)
        dW[((int64_t)f * a.Ko + o0 + to + p) * a.Kred + i0 + ti + q] = acc[p][q];
}

static size_t run_dense_bwd_ws_default(int F, int H, int Ko, int Kred, int64_t B);
static size_t run_dense_bwd_ws(int F, int H, int Ko, int Kred, int64_t B) {
  const size_t base = run_dense_bwd_ws_default(F, H, Ko, Kred, B);
  // experimental K = 128 tcgen05 backward (off by default): room for its row shifts and partials
  if ((tc_flags() & 512) && Ko == 128 && Kred == 128 && H <= 2) {
    const size_t need = dense128_tc_bwd_ws(F, B);
    return need > base ? need : base;
  }
  return base;
}
static size_t run_dense_bwd_ws_default(int F, int H, int Ko, int Kred, int64_t B) {
  if (Ko == 1 && Kred <= 32 * kKo1MaxPerLane) return (size_t)dense_ko1_blocks(F, B) * F * Kred * 4;
  const size_t tc = dense_tc_bwd_ws(F, H, Ko, Kred, B);
  if (Ko == 32 && Kred == 32 && H <= 2) return dense32_bwd_ws(F, B);
  if (!dense_small_ok(H, Ko, Kred)) return max64(tc, (size_t)F * B * 4);  // row shifts of the generic path
  int SW, splits;
  int64_t chunk;
  dense_bwd_config(F, Ko, Kred, B, SW, splits, chunk);
  const size_t simt = splits > 1 ? (size_t)splits * F * Ko * Kred * 4 : 0;
  return simt > tc ? simt : tc;
}

static int run_dense_bwd(DenseArgs a, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes) {
  {
    const int rc = dense_tc_bwd(a, F, dW, c, ws, ws_bytes);  // tensor-core path for the hot shape
    if (rc <= 0) return rc;
  }
  const size_t n = (size_t)F * a.Ko * a.Kred;
  if (dense32_ok(a)) return dense32_bwd(a, F, dW, c, ws, ws_bytes);
  if (dense128_tc_ok(a)) {  // Ki = Ko = 128 on tcgen05
    const int rc = dense128_tc_bwd(a, F, dW, c, ws, ws_bytes);
    if (rc <= 0) return rc;
  }
  if (dense_ko1_ok(a)) {
    const int blocks = dense_ko1_blocks(F, a.B);
    a.dWp = dW;
    if (dW && blocks > 1) {
      if (ws_bytes < blocks * n * 4) {
        set_error("dense_bwd: workspace too small (%zu < %zu)", ws_bytes, blocks * n * 4);
        return CKB_ERR_WORKSPACE;
      }
      a.dWp = (float*)ws;
    }
    dim3 grid(blocks, F);
    CKB_CUDA_CHECK(launch_pdl(dense_ko1_bwd_kernel, grid, dim3(256), 0, c.stream, a));
    c.launches++;
    if (dW && blocks > 1) return reduce_partials(a.dWp, dW, (int64_t)n, blocks, c);
    return CKB_OK;
  }
  if (dense_small_ok(a.H, a.Ko, a.Kred)) {
    int SW, splits;
    int64_t chunk;
    dense_bwd_config(F, a.Ko, a.Kred, a.B, SW, splits, chunk);
    a.chunk = chunk;
    a.dWp = dW;
    if (dW && splits > 1) {
      if (ws_bytes < splits * n * 4) {
        set_error("dense_bwd: workspace too small (%zu < %zu)", ws_bytes, splits * n * 4);
        return CKB_ERR_WORKSPACE;
      }
      a.dWp = (float*)ws;
    }
    const int kmax = max(a.Ko, a.Kred);
    int rc;
    if (kmax <= 32) rc = launch_dense_bwd_small<1, 2, 16>(a, F, splits, c);
    else if (kmax <= 64) rc = launch_dense_bwd_small<2, 4, 16>(a, F, splits, c);
    else rc = launch_dense_bwd_small<4, 8, 8>(a, F, splits, c);
    if (rc != CKB_OK) return rc;
    if (dW && splits > 1) return reduce_partials(a.dWp, dW, (int64_t)n, splits, c);
    return CKB_OK;
  }
  const int e_stride = (a.Kred + 3) & ~3, r_stride = (a.Ko + 3) & ~3;
  int nwarps = 8;
  while ((size_t)nwarps * (e_stride + r_stride) * 4 > 96 * 1024 && nwarps > 1) nwarps >>= 1;
  const size_t smem = (size_t)nwarps * (e_stride + r_stride) * 4;
  if (smem > 200 * 1024) {
    set_error("dense_bwd: reduction length %d too large", a.Kred);
    return CKB_ERR_UNSUPPORTED;
  }
  static PerDeviceOnce attr;
  if (attr.first()) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_bwd_generic,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  float* mrow = nullptr;
  if (dW || ws_bytes >= (size_t)F * a.B * 4) {
    if (ws_bytes < (size_t)F * a.B * 4) {
      set_error("dense_bwd: workspace too small (%zu < %zu)", ws_bytes, (size_t)F * a.B * 4);
      return CKB_ERR_WORKSPACE;
    }
    mrow = (float*)ws;
  }
  if (mrow != nullptr) {  // tiled du (see dense_fwd_tiled_generic)
    dim3 g1((int)min64(ceil_div(a.B, 8), 8 * kNumSMs), F);
    dense_rowmax_generic<<<g1, 256, 0, c.stream>>>(a, mrow);
    CKB_LAUNCH_CHECK();
    dim3 g2(ceil_div(a.B, kGtB), ceil_div(a.Kred, kGtN), F);
    dense_du_tiled_generic<<<g2, 256, 0, c.stream>>>(a, mrow);
    CKB_LAUNCH_CHECK();
    c.launches += 2;
  } else {
    dim3 grid((int)min64(ceil_div(a.B, nwarps), 8 * kNumSMs), F);
    dense_bwd_generic<<<grid, nwarps * 32, smem, c.stream>>>(a, e_stride, r_stride, mrow);
    CKB_LAUNCH_CHECK();
    c.launches++;
  }
  if (dW) {
    dim3 g2(ceil_div(a.Kred, kDwTi), ceil_div(a.Ko, kDwTo), F);
    dense_dw_generic<<<g2, 256, 0, c.stream>>>(a, mrow, dW);
    CKB_LAUNCH_CHECK();
    c.launches++;
  }
  return CKB_OK;
}

// ------------------------------------------------------------------------------------------
// step entry points
// ------------------------------------------------------------------------------------------
static DenseArgs make_args(const ckb_step_desc_t& d, Ctx& c) {
  DenseArgs a{};
  a.W = c.tensors[d.slot[0]];
  a.in_rows = d.in_rows;
  a.arena = c.arena;
  a.y = c.arena + c.B * d.out_off;
  a.B = c.B;
  a.H = d.arity;
  a.Ki = d.k_in;
  a.Ko = d.k_out;
  a.concat = (d.flags & CKB_DENSE_CONCAT) ? 1 : 0;
  a.Kred = a.concat ? d.arity * d.k_in : d.k_in;
  return a;
}

// CKB_STEP_TABLE_INPUT: the layer runs on the gathered block u (arity 1, rows f*B*Ki)
static DenseArgs table_input_args(const ckb_step_desc_t& d, Ctx& c) {
  DenseArgs a = make_args(d, c);
  a.in_rows = nullptr;
  a.arena = c.arena + c.B * d.aux_off;
  a.H = 1;
  return a;
}

int dense_fwd(const ckb_step_desc_t& d, Ctx& c) {
  if (d.flags & CKB_STEP_TABLE_INPUT) {
    if (d.flags & CKB_DENSE_CONCAT) {
      set_error("table-input step over concatenated inputs");
      return CKB_ERR_UNSUPPORTED;
    }
    if (int rc = table_pair_gather(d, c, c.arena + c.B * d.aux_off)) return rc;
    return run_dense_fwd(table_input_args(d, c), d.num_folds, c);
  }
  return run_dense_fwd(make_args(d, c), d.num_folds, c);
}

size_t dense_bwd_ws(const ckb_step_desc_t& d, int64_t B) {
  const int concat = (d.flags & CKB_DENSE_CONCAT) ? 1 : 0;
  return run_dense_bwd_ws(d.num_folds, d.arity, d.k_out, concat ? d.arity * d.k_in : d.k_in, B);
}

int dense_bwd(const ckb_step_desc_t& d, Ctx& c) {
  DenseArgs a = (d.flags & CKB_STEP_TABLE_INPUT) ? table_input_args(d, c) : make_args(d, c);
  a.gs = GradSrc{c.garena, d.cons_ptr, d.cons_rows, c.B};
  a.gin = c.garena + c.B * d.gin_off;
  a.max_cons = d.max_consumers;
  a.rows64 = (d.flags & CKB_STEP_ROWS64) ? 1 : 0;
  return run_dense_bwd(a, d.num_folds, c.grads[d.slot[0]], c, c.ws, c.ws_bytes);
}

// ------------------------------------------------------------------------------------------
// Tucker (layers/optimized.py:89-103), FP32 SIMT route for small K: the Kronecker product of the
// two shifted inputs is formed in scratch and fed to the dense block with Kred = Ki^2 (per-input
// shifts m1, m2 and a single shift over the K^2 sums agree: max_ij (x1_i + x2_j) = m1 + m2).
// ------------------------------------------------------------------------------------------
int kronecker_into(const ckb_step_desc_t& d, Ctx& c, float* dst);
int kronecker_bwd_from(const ckb_step_desc_t& d, Ctx& c, const float* gsrc);

bool tucker_root_ok(const ckb_step_desc_t& d);
size_t tucker_root_ws(const ckb_step_desc_t& d, int64_t B);
int tucker_root_fwd(const ckb_step_desc_t& d, Ctx& c);
int tucker_root_bwd(const ckb_step_desc_t& d, Ctx& c);
bool tucker_tc_ok(const ckb_step_desc_t& d);
size_t tucker_tc_ws(const ckb_step_desc_t& d, int64_t B);
int tucker_tc_fwd(const ckb_step_desc_t& d, Ctx& c);
int tucker_tc_bwd(const ckb_step_desc_t& d, Ctx& c);

size_t tucker_ws(const ckb_step_desc_t& d, int64_t B) {
  if (tucker_tc_ok(d)) return tucker_tc_ws(d, B) + 256;
  if (tucker_root_ok(d)) return tucker_root_ws(d, B) + 256;
  const size_t kron = (size_t)d.num_folds * B * d.k_in * d.k_in * 4;
  return 2 * kron + run_dense_bwd_ws(d.num_folds, 1, d.k_out, d.k_in * d.k_in, B) + 256;
}

static int tucker_check(const ckb_step_desc_t& d) {
  if (d.arity != 2) {
    set_error("tucker: arity %d has no kernel (2 only)", d.arity);
    return CKB_ERR_UNSUPPORTED;
  }
  return CKB_OK;
}

static DenseArgs tucker_args(const ckb_step_desc_t& d, Ctx& c, float* kron) {
  DenseArgs a{};
  a.W = c.tensors[d.slot[0]];
  a.in_rows = nullptr;
  a.arena = kron;
  a.y = c.arena + c.B * d.out_off;
  a.B = c.B;
  a.H = 1;
  a.Ki = d.k_in * d.k_in;
  a.Ko = d.k_out;
  a.concat = 0;
  a.Kred = a.Ki;
  return a;
}

int tucker_fwd(const ckb_step_desc_t& d, Ctx& c) {
  if (int rc = tucker_check(d)) return rc;
  if (tucker_tc_ok(d)) return tucker_tc_fwd(d, c);
  if (tucker_root_ok(d)) return tucker_root_fwd(d, c);
  const size_t kron_bytes = (size_t)d.num_folds * c.B * d.k_in * d.k_in * 4;
  if (c.ws_bytes < kron_bytes) {
    set_error("tucker_fwd: workspace too small");
    return CKB_ERR_WORKSPACE;
  }
  float* kron = (float*)c.ws;
  if (int rc = kronecker_into(d, c, kron)) return rc;
  // the dense block may use scratch of its own (row shifts of the tiled generic kernels): behind kron
  const size_t used = (kron_bytes + 255) & ~(size_t)255;
  Ctx c2 = c;
  c2.ws = c.ws_bytes > used ? c.ws + used : nullptr;
  c2.ws_bytes = c.ws_bytes > used ? c.ws_bytes - used : 0;
  const int rc = run_dense_fwd(tucker_args(d, c, kron), d.num_folds, c2);
  c.launches = c2.launches;
  return rc;
}

int tucker_bwd(const ckb_step_desc_t& d, Ctx& c) {
  if (int rc = tucker_check(d)) return rc;
  if (tucker_tc_ok(d)) return tucker_tc_bwd(d, c);
  if (tucker_root_ok(d)) return tucker_root_bwd(d, c);
  const size_t kron_bytes = ((size_t)d.num_folds * c.B * d.k_in * d.k_in * 4 + 255) & ~(size_t)255;
  if (c.ws_bytes < 2 * kron_bytes) {
    set_error("tucker_bwd: workspace too small");
    return CKB_ERR_WORKSPACE;
  }
  float* kron = (float*)c.ws;
  float* gkron = (float*)(c.ws + kron_bytes);
  if (int rc = kronecker_into(d, c, kron)) return rc;
  DenseArgs a = tucker_args(d, c, kron);
  a.gs = GradSrc{c.garena, d.cons_ptr, d.cons_rows, c.B};
  a.gin = gkron;
  a.max_cons = d.max_consumers;
  if (int rc = run_dense_bwd(a, d.num_folds, c.grads[d.slot[0]], c, c.ws + 2 * kron_bytes,
                             c.ws_bytes - 2 * kron_bytes))
    return rc;
  return kronecker_bwd_from(d, c, gkron);
}

// ------------------------------------------------------------------------------------------
// TABLE_DENSE: the dense block applied to the table rows instead of to every sample.
// ------------------------------------------------------------------------------------------
static DenseArgs table_dense_args(const ckb_step_desc_t& d, Ctx& c) {
  DenseArgs a{};
  a.W = c.tensors[d.slot[1]];
  a.in_rows = nullptr;
  a.arena = c.tensors[d.slot[0]];  // (F, V, Ki): "samples" are the V states
  a.y = c.tensors[d.slot[2]];      // (F, V, Ko)
  a.B = d.num_states;
  a.H = 1;
  a.Ki = d.k_in;
  a.Ko = d.k_out;
  a.concat = 0;
  a.Kred = d.k_in;
  return a;
}

static ckb_step_desc_t as_table(const ckb_step_desc_t& d) {
  ckb_step_desc_t t = d;
  t.kind = CKB_STEP_TABLE;
  t.slot[0] = d.slot[2];
  t.int_slot = -1;
  return t;
}

size_t table_dense_ws(const ckb_step_desc_t& d, int64_t B) {
  const size_t a = (table_bwd_ws(as_table(d), B) + 255) & ~(size_t)255;
  return a + run_dense_bwd_ws(d.num_folds, 1, d.k_out, d.k_in, d.num_states) + 256;
}

int table_dense_fwd(const ckb_step_desc_t& d, Ctx& c) {
  if (c.maskT != nullptr) {
    set_error("table_dense: integration masks need the unfused plan");
    return CKB_ERR_UNSUPPORTED;
  }
  if (int rc = run_dense_fwd(table_dense_args(d, c), d.num_folds, c)) return rc;
  if (d.flags & CKB_STEP_NO_GATHER) return CKB_OK;  // the consumer gathers T2 rows itself
  return table_fwd(as_table(d), c);
}

int table_dense_bwd(const ckb_step_desc_t& d, Ctx& c) {
  float* dT2 = c.grads[d.slot[2]];
  float* dT = c.grads[d.slot[0]];
  if (dT2 == nullptr || dT == nullptr) {
    set_error("table_dense_bwd: gradient buffers of T and T2 are required");
    return CKB_ERR_INVALID;
  }
  if (int rc = table_bwd(as_table(d), c)) return rc;  // dT2[f,v,:] = sum of g over samples in state v
  DenseArgs a = table_dense_args(d, c);
  a.gs = GradSrc{dT2, nullptr, nullptr, (int64_t)d.num_states};
  a.gin = dT;
  a.max_cons = 1;
  a.rows64 = 1;  // tables are dense (F, V, K) blocks
  const size_t off = (table_bwd_ws(as_table(d), c.B) + 255) & ~(size_t)255;
  return run_dense_bwd(a, d.num_folds, c.grads[d.slot[1]], c, c.ws + off,
                       c.ws_bytes > off ? c.ws_bytes - off : 0);
}

}  // namespace ckb
