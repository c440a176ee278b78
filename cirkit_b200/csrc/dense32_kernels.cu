// The fused sum-product block for Ki = Ko = 32 (config 2 of BASELINE.json: QuadTree K = 32), FP32.
// Reference: TorchCPTLayer / TorchSumLayer (arity 1) through LSESumSemiring.apply_reduce,
// cirkit/backend/torch/layers/optimized.py:171-178, inner.py:266-273, semiring.py:382-408.
// At K = 32 a row is exactly one warp-wide 128-byte access and the block is HBM-bound (5 flop/B),
// so the kernels are organised around the loads: a warp keeps its weight rows in registers, takes
// four samples at a time (all their loads in flight together), lane = unit, and broadcasts e / r
// through a few hundred bytes of shared memory.  No per-sample synchronisation chain as in the
// generic small-shape kernels (measured 0.64 ms -> see DESIGN.md for the F = 392 forward launch).
#include "dense.cuh"
#include "sm100.cuh"

namespace ckb {
namespace {

constexpr int K32 = 32;
constexpr int kWarps = 8;
constexpr int kS = 4;   // samples per warp iteration (backward)
constexpr int kSF = 4;  // forward (8 in flight measured slower: 80 registers, one CTA fewer per SM)

__device__ __forceinline__ float ldg_or(const float* p, bool ok, float other) { return ok ? __ldg(p) : other; }

__global__ void __launch_bounds__(kWarps * 32) dense32_fwd_kernel(DenseArgs a) {
  __shared__ __align__(16) float es[kWarps][kSF][K32];
  const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* x0 = in_row(a, f, 0);
  const float* x1 = a.H == 2 ? in_row(a, f, 1) : nullptr;
  float w[K32];  // W[f][o = lane][:]
  {
    const float4* wr = reinterpret_cast<const float4*>(a.W + ((int64_t)f * K32 + lane) * K32);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 v = __ldg(wr + c);
      w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
    }
  }
  float* yf = a.y + (int64_t)f * a.B * K32;
  for (int64_t b0 = ((int64_t)blockIdx.x * kWarps + warp) * kSF; b0 < a.B; b0 += (int64_t)gridDim.x * kWarps * kSF) {
    float u[kSF], m[kSF];
#pragma unroll
    for (int s = 0; s < kSF; ++s) {
      const bool ok = b0 + s < a.B;
      u[s] = ldg_or(x0 + (b0 + s) * K32 + lane, ok, 0.f);
      if (x1) u[s] += ldg_or(x1 + (b0 + s) * K32 + lane, ok, 0.f);
    }
#pragma unroll
    for (int s = 0; s < kSF; ++s) {
      m[s] = clamp_max(warp_max(u[s]));
      es[warp][s][lane] = sm100::fast_exp(u[s] - m[s]);
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kSF; ++s) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 e4 = *reinterpret_cast<const float4*>(&es[warp][s][4 * c]);
        acc = fmaf(e4.x, w[4 * c], acc);
        acc = fmaf(e4.y, w[4 * c + 1], acc);
        acc = fmaf(e4.z, w[4 * c + 2], acc);
        acc = fmaf(e4.w, w[4 * c + 3], acc);
      }
      if (b0 + s < a.B) yf[(b0 + s) * K32 + lane] = sm100::fast_log(acc) + m[s];
    }
    __syncwarp();
  }
}

// backward: r = g * exp(m - y); du[i] = e[i] * sum_o r[o] W[o,i]; dW[o,i] += r[o] e[i]
__global__ void __launch_bounds__(kWarps * 32, 2) dense32_bwd_kernel(DenseArgs a) {
  __shared__ __align__(16) float es[kWarps][kS][K32];
  __shared__ __align__(16) float rs[kWarps][kS][K32];
  __shared__ float red[kWarps][K32][K32 + 1];
  const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* x0 = in_row(a, f, 0);
  const float* x1 = a.H == 2 ? in_row(a, f, 1) : nullptr;
  const float* Wf = a.W + (int64_t)f * K32 * K32;
  float wt[K32];  // W[f][:, i = lane]  (column of W: the row of W^T this lane contracts with r)
  float dw[K32];  // dW[f][o = lane][:] partial of this warp
#pragma unroll
  for (int o = 0; o < K32; ++o) {
    wt[o] = __ldg(Wf + o * K32 + lane);
    dw[o] = 0.f;
  }
  const float* yf = a.y + (int64_t)f * a.B * K32;
  float* gin = a.gin + (int64_t)f * a.B * K32;
  const bool want_dw = a.dWp != nullptr;
  // the usual case (a tree): one consumer row, resolved once instead of per sample
  const float* grow = nullptr;
  if (a.gs.cons_ptr == nullptr) {
    grow = a.gs.garena + (int64_t)f * a.gs.B * K32;
  } else if (a.gs.cons_ptr[f + 1] - a.gs.cons_ptr[f] == 1) {
    grow = a.gs.garena + a.gs.B * a.gs.cons_rows[a.gs.cons_ptr[f]];
  }
  for (int64_t b0 = ((int64_t)blockIdx.x * kWarps + warp) * kS; b0 < a.B; b0 += (int64_t)gridDim.x * kWarps * kS) {
    float u[kS], yv[kS], g[kS];
#pragma unroll
    for (int s = 0; s < kS; ++s) {
      const bool ok = b0 + s < a.B;
      u[s] = ldg_or(x0 + (b0 + s) * K32 + lane, ok, 0.f);
      if (x1) u[s] += ldg_or(x1 + (b0 + s) * K32 + lane, ok, 0.f);
      yv[s] = ldg_or(yf + (b0 + s) * K32 + lane, ok, 0.f);
      g[s] = !ok ? 0.f : (grow ? __ldg(grow + (b0 + s) * K32 + lane) : pull_grad(a.gs, f, b0 + s, K32, lane));
    }
    float e[kS];
#pragma unroll
    for (int s = 0; s < kS; ++s) {
      const float m = clamp_max(warp_max(u[s]));
      e[s] = sm100::fast_exp(u[s] - m);
      const float r = (g[s] == 0.f) ? 0.f : g[s] * sm100::fast_exp_finite(fminf(fmaxf(m - yv[s], -104.f), 88.f));
      es[warp][s][lane] = e[s];
      rs[warp][s][lane] = r;
      g[s] = r;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kS; ++s) {
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 r4 = *reinterpret_cast<const float4*>(&rs[warp][s][4 * c]);
        t = fmaf(r4.x, wt[4 * c], t);
        t = fmaf(r4.y, wt[4 * c + 1], t);
        t = fmaf(r4.z, wt[4 * c + 2], t);
        t = fmaf(r4.w, wt[4 * c + 3], t);
      }
      if (b0 + s < a.B) gin[(b0 + s) * K32 + lane] = e[s] * t;
      if (want_dw) {
        const float r = g[s];  // r[o = lane]
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 e4 = *reinterpret_cast<const float4*>(&es[warp][s][4 * c]);
          dw[4 * c] = fmaf(r, e4.x, dw[4 * c]);
          dw[4 * c + 1] = fmaf(r, e4.y, dw[4 * c + 1]);
          dw[4 * c + 2] = fmaf(r, e4.z, dw[4 * c + 2]);
          dw[4 * c + 3] = fmaf(r, e4.w, dw[4 * c + 3]);
        }
      }
    }
    __syncwarp();
  }
  if (!want_dw) return;
  // deterministic combination of the 8 warps' partials, then one slab per CTA
#pragma unroll
  for (int i = 0; i < K32; ++i) red[warp][lane][i] = dw[i];
  __syncthreads();
  float* out = a.dWp + ((int64_t)blockIdx.x * gridDim.y + f) * K32 * K32;
  for (int idx = threadIdx.x; idx < K32 * K32; idx += kWarps * 32) {
    const int o = idx >> 5, i = idx & 31;
    float sacc = 0.f;
#pragma unroll
    for (int wv = 0; wv < kWarps; ++wv) sacc += red[wv][o][i];
    out[idx] = sacc;
  }
}

int dense32_splits(int F, int64_t B) {
  return (int)max64(1, min64(ceil_div(B, kWarps * kSF), ceil_div(8 * kNumSMs, F)));
}

}  // namespace

bool dense32_ok(const DenseArgs& a) {
  return a.Ki == K32 && a.Ko == K32 && a.Kred == K32 && !a.concat && a.H >= 1 && a.H <= 2;
}

size_t dense32_bwd_ws(int F, int64_t B) {
  const int splits = dense32_splits(F, B);
  return splits > 1 ? (size_t)splits * F * K32 * K32 * 4 : 0;
}

int dense32_fwd(const DenseArgs& a, int F, Ctx& c) {
  dim3 grid(dense32_splits(F, a.B), F);
  dense32_fwd_kernel<<<grid, kWarps * 32, 0, c.stream>>>(a);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

int dense32_bwd(const DenseArgs& a_in, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes) {
  DenseArgs a = a_in;
  const int splits = dense32_splits(F, a.B);
  const size_t n = (size_t)F * K32 * K32;
  a.dWp = dW;
  if (dW && splits > 1) {
    if (ws_bytes < splits * n * 4) {
      set_error("dense32_bwd: workspace too small (%zu < %zu)", ws_bytes, splits * n * 4);
      return CKB_ERR_WORKSPACE;
    }
    a.dWp = (float*)ws;
  }
  dim3 grid(splits, F);
  dense32_bwd_kernel<<<grid, kWarps * 32, 0, c.stream>>>(a);
  CKB_LAUNCH_CHECK();
  c.launches++;
  if (dW && splits > 1) return reduce_partials(a.dWp, dW, (int64_t)n, splits, c);
  return CKB_OK;
}

}  // namespace ckb
