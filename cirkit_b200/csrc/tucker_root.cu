// Tucker layer with a single output unit (the root of a Tucker circuit: Ko = 1, Ki <= 64), FP32.
// Reference: TorchTuckerLayer.forward, cirkit/backend/torch/layers/optimized.py:89-103.
//   y[b] = log( sum_ij W[i,j] e1[b,i] e2[b,j] ) + m1 + m2
// One warp per sample computes t_i = sum_j W[i,j] e2[j] and u_j = sum_i W[i,j] e1[i] from the
// weight slice in shared memory (no (B, Ki^2) Kronecker scratch as in the generic route), the
// batch reduction of dW runs in registers per CTA and is combined by reduce_partials.
#include "dense.cuh"

namespace ckb {
namespace {

constexpr int kRootMaxK = 64;
constexpr int kRootWarps = 8;

struct RootSample {
  float e1[2], e2[2];  // units lane, lane + 32
  float ms;
};

__device__ __forceinline__ RootSample root_load(const DenseArgs& a, int f, int64_t b, int lane, int K) {
  const float* x1 = in_row(a, f, 0) + b * K;
  const float* x2 = in_row(a, f, 1) + b * K;
  float v1[2], v2[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int k = lane + 32 * t;
    v1[t] = k < K ? x1[k] : -INFINITY;
    v2[t] = k < K ? x2[k] : -INFINITY;
  }
  const float m1 = clamp_max(warp_max(fmaxf(v1[0], v1[1])));
  const float m2 = clamp_max(warp_max(fmaxf(v2[0], v2[1])));
  RootSample s;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    s.e1[t] = expf(v1[t] - m1);
    s.e2[t] = expf(v2[t] - m2);
  }
  s.ms = fmaxf(m1 + m2, -FLT_MAX);
  return s;
}

// W (K x K) -> shared memory with rows padded to K + 1 floats
__device__ __forceinline__ void root_stage_w(const float* Wf, float* w, int K) {
  for (int idx = threadIdx.x; idx < K * K; idx += blockDim.x) w[(idx / K) * (K + 1) + idx % K] = Wf[idx];
}

__global__ void __launch_bounds__(kRootWarps * 32) tucker_root_fwd_kernel(DenseArgs a, int K) {
  extern __shared__ float sm[];
  float* w = sm;                               // [K][K+1]
  float* ev = sm + K * (K + 1);                // [warps][2][64]
  const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  root_stage_w(a.W + (int64_t)f * K * K, w, K);
  __syncthreads();
  float* e1s = ev + warp * 128;
  float* e2s = e1s + 64;
  for (int64_t b = (int64_t)blockIdx.x * kRootWarps + warp; b < a.B; b += (int64_t)gridDim.x * kRootWarps) {
    const RootSample s = root_load(a, f, b, lane, K);
    __syncwarp();
    e2s[lane] = s.e2[0];
    e2s[lane + 32] = s.e2[1];
    __syncwarp();
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = lane + 32 * t;
      if (i < K) {
        float ti = 0.f;
        for (int j = 0; j < K; ++j) ti = fmaf(w[i * (K + 1) + j], e2s[j], ti);
        acc = fmaf(s.e1[t], ti, acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) a.y[(int64_t)f * a.B + b] = logf(acc) + s.ms;
  }
}

__global__ void __launch_bounds__(kRootWarps * 32)
tucker_root_bwd_kernel(DenseArgs a, int K, float* dWp) {
  extern __shared__ float sm[];
  float* w = sm;                    // [K][K+1]
  float* ev = sm + K * (K + 1);     // [warps][2][64]: r*e1 and e2 of the warp's sample
  const int f = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  root_stage_w(a.W + (int64_t)f * K * K, w, K);
  // this thread's 16 entries of dW: row di, columns dj0 .. dj0+15 (K = 64: 256 threads x 16)
  const int per = (K * K + 255) / 256;
  float dw[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) dw[n] = 0.f;
  __syncthreads();
  float* a1s = ev + warp * 128;
  float* e2s = a1s + 64;
  float* g1 = a.gin + ((int64_t)f * 2 + 0) * a.B * K;
  float* g2 = a.gin + ((int64_t)f * 2 + 1) * a.B * K;
  const int64_t n_iter = (a.B + (int64_t)gridDim.x * kRootWarps - 1) / ((int64_t)gridDim.x * kRootWarps);
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t b = (it * gridDim.x + blockIdx.x) * kRootWarps + warp;
    const bool valid = b < a.B;
    float r = 0.f;
    RootSample s{};
    if (valid) {
      s = root_load(a, f, b, lane, K);
      const float g = pull_grad(a.gs, f, b, 1, 0);
      r = (g == 0.f) ? 0.f : g * expf(fminf(s.ms - a.y[(int64_t)f * a.B + b], 88.f));
    }
    __syncthreads();  // the previous iteration's dW pass has read the staging rows
    a1s[lane] = s.e1[0];
    a1s[lane + 32] = s.e1[1];
    e2s[lane] = s.e2[0];
    e2s[lane + 32] = s.e2[1];
    __syncwarp();
    if (valid) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int k = lane + 32 * t;
        if (k < K) {
          float ti = 0.f, uj = 0.f;
          for (int j = 0; j < K; ++j) {
            ti = fmaf(w[k * (K + 1) + j], e2s[j], ti);   // sum_j W[k,j] e2[j]
            uj = fmaf(w[j * (K + 1) + k], a1s[j], uj);   // sum_i W[i,k] e1[i]
          }
          g1[b * K + k] = s.e1[t] * r * ti;
          g2[b * K + k] = s.e2[t] * r * uj;
        }
      }
    }
    __syncwarp();
    a1s[lane] = r * s.e1[0];
    a1s[lane + 32] = r * s.e1[1];
    __syncthreads();
    if (dWp != nullptr) {
      for (int wv = 0; wv < kRootWarps; ++wv) {
        const float* pa = ev + wv * 128;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          const int idx = tid * per + n;
          if (n < per && idx < K * K) dw[n] = fmaf(pa[idx / K], pa[64 + idx % K], dw[n]);
        }
      }
    }
  }
  if (dWp != nullptr) {
    float* o = dWp + ((int64_t)blockIdx.x * gridDim.y + f) * K * K;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int idx = tid * per + n;
      if (n < per && idx < K * K) o[idx] = dw[n];
    }
  }
}

int root_blocks(int F, int64_t B) {
  return (int)max64(1, min64(ceil_div(B, kRootWarps), ceil_div(2 * kNumSMs, F)));
}
size_t root_smem(int K) { return (size_t)(K * (K + 1) + kRootWarps * 128) * 4; }

}  // namespace

bool tucker_root_ok(const ckb_step_desc_t& d) {
  return d.arity == 2 && d.k_out == 1 && d.k_in <= kRootMaxK && d.k_in * d.k_in <= 256 * 16;
}

size_t tucker_root_ws(const ckb_step_desc_t& d, int64_t B) {
  return (size_t)root_blocks(d.num_folds, B) * d.num_folds * d.k_in * d.k_in * 4;
}

static DenseArgs root_args(const ckb_step_desc_t& d, Ctx& c) {
  DenseArgs a{};
  a.W = c.tensors[d.slot[0]];
  a.in_rows = d.in_rows;
  a.arena = c.arena;
  a.y = c.arena + c.B * d.out_off;
  a.B = c.B;
  a.H = 2;
  a.Ki = d.k_in;
  a.Ko = 1;
  a.Kred = d.k_in * d.k_in;
  return a;
}

int tucker_root_fwd(const ckb_step_desc_t& d, Ctx& c) {
  const DenseArgs a = root_args(d, c);
  dim3 grid(root_blocks(d.num_folds, c.B), d.num_folds);
  tucker_root_fwd_kernel<<<grid, kRootWarps * 32, root_smem(d.k_in), c.stream>>>(a, d.k_in);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

int tucker_root_bwd(const ckb_step_desc_t& d, Ctx& c) {
  DenseArgs a = root_args(d, c);
  a.gs = GradSrc{c.garena, d.cons_ptr, d.cons_rows, c.B};
  a.gin = c.garena + c.B * d.gin_off;
  float* dW = c.grads[d.slot[0]];
  const int blocks = root_blocks(d.num_folds, c.B);
  const size_t n = (size_t)d.num_folds * d.k_in * d.k_in;
  float* dWp = dW;
  if (dW && blocks > 1) {
    if (c.ws_bytes < blocks * n * 4) {
      set_error("tucker_root_bwd: workspace too small (%zu < %zu)", c.ws_bytes, blocks * n * 4);
      return CKB_ERR_WORKSPACE;
    }
    dWp = (float*)c.ws;
  }
  dim3 grid(blocks, d.num_folds);
  tucker_root_bwd_kernel<<<grid, kRootWarps * 32, root_smem(d.k_in), c.stream>>>(a, d.k_in, dWp);
  CKB_LAUNCH_CHECK();
  c.launches++;
  if (dW && blocks > 1) return reduce_partials(dWp, dW, (int64_t)n, blocks, c);
  return CKB_OK;
}

}  // namespace ckb
