// Product layers (Hadamard, Kronecker) and the mixing sum layer.
#include "common.cuh"

namespace ckb {

// ------------------------------------------------------------------------------------------
// Hadamard, layers/inner.py:126-127: a product is a sum in log space.
// ------------------------------------------------------------------------------------------
__global__ void hadamard_fwd_kernel(const float* __restrict__ arena, const int64_t* __restrict__ in_rows,
                                    float* __restrict__ y, int64_t B, int H, int K) {
  const int f = blockIdx.y;
  const int64_t total = B * K;
  float* yf = y + (int64_t)f * total;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int h = 0; h < H; ++h) s += arena[B * in_rows[f * H + h] + idx];
    yf[idx] = s;
  }
}

int hadamard_fwd(const ckb_step_desc_t& d, Ctx& c) {
  const int bx = (int)min64(ceil_div(c.B * d.k_out, 256), 4 * kNumSMs);
  dim3 grid(max(bx, 1), d.num_folds);
  hadamard_fwd_kernel<<<grid, 256, 0, c.stream>>>(c.arena, d.in_rows, c.arena + c.B * d.out_off,
                                                  c.B, d.arity, d.k_out);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// every input of a fold receives the fold's output gradient (gin_h == 1)
__global__ void hadamard_bwd_kernel(GradSrc gs, float* __restrict__ gin, int64_t B, int K) {
  const int f = blockIdx.y;
  const int64_t total = B * K;
  float* gf = gin + (int64_t)f * total;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / K;
    gf[idx] = pull_grad(gs, f, b, K, (int)(idx - b * K));
  }
}

int hadamard_bwd(const ckb_step_desc_t& d, Ctx& c) {
  GradSrc gs{c.garena, d.cons_ptr, d.cons_rows, c.B};
  const int bx = (int)min64(ceil_div(c.B * d.k_out, 256), 4 * kNumSMs);
  dim3 grid(max(bx, 1), d.num_folds);
  hadamard_bwd_kernel<<<grid, 256, 0, c.stream>>>(gs, c.garena + c.B * d.gin_off, c.B, d.k_out);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// ------------------------------------------------------------------------------------------
// Kronecker (arity 2), layers/inner.py:178-187: y[(i,j)] = x0[i] + x1[j], i major.
// ------------------------------------------------------------------------------------------
__global__ void kronecker_fwd_kernel(const float* __restrict__ arena, const int64_t* __restrict__ in_rows,
                                     float* __restrict__ y, int64_t B, int K) {
  const int f = blockIdx.y;
  const int KK = K * K;
  const float* x0 = arena + B * in_rows[f * 2 + 0];
  const float* x1 = arena + B * in_rows[f * 2 + 1];
  const int64_t total = B * KK;
  float* yf = y + (int64_t)f * total;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / KK;
    const int ij = (int)(idx - b * KK);
    const int i = ij / K, j = ij - i * K;
    yf[idx] = x0[b * K + i] + x1[b * K + j];
  }
}

int kronecker_into(const ckb_step_desc_t& d, Ctx& c, float* dst) {
  if (d.arity != 2) {
    set_error("kronecker: arity %d has no kernel (2 only)", d.arity);
    return CKB_ERR_UNSUPPORTED;
  }
  const int bx = (int)min64(ceil_div(c.B * d.k_in * d.k_in, 256), 8 * kNumSMs);
  dim3 grid(max(bx, 1), d.num_folds);
  kronecker_fwd_kernel<<<grid, 256, 0, c.stream>>>(c.arena, d.in_rows, dst, c.B, d.k_in);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

int kronecker_fwd(const ckb_step_desc_t& d, Ctx& c) {
  return kronecker_into(d, c, c.arena + c.B * d.out_off);
}

// one warp per (fold, sample): row sums -> d/dx0, column sums -> d/dx1
__global__ void kronecker_bwd_kernel(GradSrc gs, float* __restrict__ gin, int64_t B, int K) {
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int KK = K * K;
  float* g0 = gin + ((int64_t)f * 2 + 0) * B * K;
  float* g1 = gin + ((int64_t)f * 2 + 1) * B * K;
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < B; b += (int64_t)gridDim.x * nwarps) {
    for (int j0 = 0; j0 < K; j0 += 32) {
      const int j = j0 + lane;
      float col = 0.f;
      for (int i = 0; i < K; ++i) {
        const float g = (j < K) ? pull_grad(gs, f, b, KK, i * K + j) : 0.f;
        col += g;
        const float row = warp_sum(g);
        if (lane == 0) {
          if (j0 == 0) g0[b * K + i] = row;
          else g0[b * K + i] += row;
        }
      }
      if (j < K) g1[b * K + j] = col;
    }
  }
}

int kronecker_bwd_from(const ckb_step_desc_t& d, Ctx& c, const float* gsrc) {
  GradSrc gs{gsrc, nullptr, nullptr, c.B};
  dim3 grid((int)min64(ceil_div(c.B, 8), 4 * kNumSMs), d.num_folds);
  kronecker_bwd_kernel<<<grid, 256, 0, c.stream>>>(gs, c.garena + c.B * d.gin_off, c.B, d.k_in);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

int kronecker_bwd(const ckb_step_desc_t& d, Ctx& c) {
  GradSrc gs{c.garena, d.cons_ptr, d.cons_rows, c.B};
  dim3 grid((int)min64(ceil_div(c.B, 8), 4 * kNumSMs), d.num_folds);
  kronecker_bwd_kernel<<<grid, 256, 0, c.stream>>>(gs, c.garena + c.B * d.gin_off, c.B, d.k_in);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// ------------------------------------------------------------------------------------------
// Mixing: a sum layer of arity H whose (F,K,H*K) weight is the block-diagonal expansion of
// (F,K,H) mixing weights (parameters/nodes.py:857-862).  The reference materialises that dense
// weight every step and runs the dense LSE-einsum; the result only depends on the H*K
// non-zeros:  y[o] = log sum_h w[o,h] exp(x_h[o] - m) + m,  m = max over ALL H*K inputs
// (one shift per row, as `apply_reduce` takes the max over the flattened axis).
// One warp per (fold, sample).
// ------------------------------------------------------------------------------------------
__global__ void mixing_fwd_kernel(const float* __restrict__ arena, const int64_t* __restrict__ in_rows,
                                  const float* __restrict__ w, float* __restrict__ y, int64_t B,
                                  int H, int K) {
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* wf = w + (int64_t)f * K * H;
  for (int64_t b = (int64_t)blockIdx.x * nwarps + warp; b < B; b += (int64_t)gridDim.x * nwarps) {
    float m = -INFINITY;
    for (int h = 0; h < H; ++h) {
      const float* xr = arena + B * in_rows[f * H + h] + b * K;
      for (int k = lane; k < K; k += 32) m = fmaxf(m, xr[k]);
    }
    m = clamp_max(warp_max(m));
    for (int k = lane; k < K; k += 32) {
      float s = 0.f;
      for (int h = 0; h < H; ++h)
        s += wf[k * H + h] * expf(arena[B * in_rows[f * H + h] + b * K + k] - m);
      y[((int64_t)f * B + b) * K + k] = logf(s) + m;
    }
  }
}

int mixing_fwd(const ckb_step_desc_t& d, Ctx& c) {
  dim3 grid((int)min64(ceil_div(c.B, 8), 4 * kNumSMs), d.num_folds);
  mixing_fwd_kernel<<<grid, 256, 0, c.stream>>>(c.arena, d.in_rows, c.tensors[d.slot[0]],
                                                c.arena + c.B * d.out_off, c.B, d.arity, d.k_out);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// r[o] = g[o] / S[o] with S[o] = exp(y[o] - m);  dx_h[o] = r[o] w[o,h] e_h[o];
// dw[o,h] = sum_b r[o] e_h[o].  Deterministic: a lane owns the units k = lane, lane + 32, ... of
// its warp's private accumulator block (no two threads ever add to the same word), the warps'
// blocks are summed in warp order at the end, and the CTAs' slabs by reduce_partials.
__global__ void mixing_bwd_kernel(const float* __restrict__ arena, const int64_t* __restrict__ in_rows,
                                  const float* __restrict__ w, const float* __restrict__ y, GradSrc gs,
                                  float* __restrict__ gin, float* __restrict__ dw_out, int64_t B, int H,
                                  int K, int64_t chunk) {
  extern __shared__ float dw_s[];  // [nwarps][K][H]
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* wf = w + (int64_t)f * K * H;
  float* mine = dw_s + (size_t)warp * K * H;
  for (int i = threadIdx.x; i < nwarps * K * H; i += blockDim.x) dw_s[i] = 0.f;
  __syncthreads();
  const int64_t b_begin = (int64_t)blockIdx.x * chunk, b_end = min(B, b_begin + chunk);
  for (int64_t b = b_begin + warp; b < b_end; b += nwarps) {
    float m = -INFINITY;
    for (int h = 0; h < H; ++h) {
      const float* xr = arena + B * in_rows[f * H + h] + b * K;
      for (int k = lane; k < K; k += 32) m = fmaxf(m, xr[k]);
    }
    m = clamp_max(warp_max(m));
    for (int k = lane; k < K; k += 32) {
      const float g = pull_grad(gs, f, b, K, k);
      const float r = (g == 0.f) ? 0.f : g * expf(m - y[((int64_t)f * B + b) * K + k]);
      for (int h = 0; h < H; ++h) {
        const float e = expf(arena[B * in_rows[f * H + h] + b * K + k] - m);
        gin[(((int64_t)f * H + h) * B + b) * K + k] = r * wf[k * H + h] * e;
        if (dw_out) mine[k * H + h] += r * e;
      }
    }
  }
  __syncthreads();
  if (dw_out) {
    float* o = dw_out + ((int64_t)blockIdx.x * gridDim.y + f) * K * H;
    for (int i = threadIdx.x; i < K * H; i += blockDim.x) {
      float acc = 0.f;
      for (int q = 0; q < nwarps; ++q) acc += dw_s[(size_t)q * K * H + i];
      o[i] = acc;
    }
  }
}

static void mixing_bwd_config(const ckb_step_desc_t& d, int64_t B, int& splits, int64_t& chunk) {
  const int64_t want = ceil_div(4 * kNumSMs, d.num_folds);
  splits = (int)max64(1, min64(want, ceil_div(B, 64)));
  chunk = ceil_div(B, splits);
  splits = ceil_div(B, chunk);
}

size_t mixing_bwd_ws(const ckb_step_desc_t& d, int64_t B) {
  int splits;
  int64_t chunk;
  mixing_bwd_config(d, B, splits, chunk);
  return splits > 1 ? (size_t)splits * d.num_folds * d.k_out * d.arity * 4 : 0;
}

int mixing_bwd(const ckb_step_desc_t& d, Ctx& c) {
  int splits;
  int64_t chunk;
  mixing_bwd_config(d, c.B, splits, chunk);
  float* dw = c.grads[d.slot[0]];
  const size_t n = (size_t)d.num_folds * d.k_out * d.arity;
  float* out = dw;
  if (dw && splits > 1) {
    if (c.ws_bytes < splits * n * 4) {
      set_error("mixing_bwd: workspace too small");
      return CKB_ERR_WORKSPACE;
    }
    out = (float*)c.ws;
  }
  int nwarps = 8;  // one private (K, H) accumulator block per warp
  while (nwarps > 1 && (size_t)nwarps * d.k_out * d.arity * 4 > 48 * 1024) nwarps >>= 1;
  const size_t smem = (size_t)nwarps * d.k_out * d.arity * 4;
  if (smem > 48 * 1024) {
    set_error("mixing_bwd: K*H = %d*%d exceeds the shared accumulator", d.k_out, d.arity);
    return CKB_ERR_UNSUPPORTED;
  }
  GradSrc gs{c.garena, d.cons_ptr, d.cons_rows, c.B};
  dim3 grid(splits, d.num_folds);
  mixing_bwd_kernel<<<grid, 32 * nwarps, smem, c.stream>>>(c.arena, d.in_rows, c.tensors[d.slot[0]],
                                                   c.arena + c.B * d.out_off, gs,
                                                   c.garena + c.B * d.gin_off, out, c.B, d.arity,
                                                   d.k_out, chunk);
  CKB_LAUNCH_CHECK();
  c.launches++;
  if (dw && splits > 1) return reduce_partials(out, dw, (int64_t)n, splits, c);
  return CKB_OK;
}

}  // namespace ckb
