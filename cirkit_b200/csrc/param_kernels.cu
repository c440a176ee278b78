// Parameter re-parameterisation ops (cirkit/backend/torch/parameters/nodes.py) and their backward.
// The reference re-evaluates every layer's parameter graph on each forward
// (parameters/parameter.py:180-188): softmax over all sum weights, softmax + log over all
// Categorical tables.  These kernels do the same work in one pass per tensor and emit the layout
// the layer kernels read ((F,V,K) log-tables).
#include "common.cuh"

namespace ckb {

// ---- softmax over the last axis, one warp per row ------------------------------------------
__global__ void softmax_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                   int64_t rows, int cols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < rows; r += (int64_t)gridDim.x * nwarps) {
    const float* s = src + r * cols;
    float m = -INFINITY;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, s[c]);
    m = warp_max(m);
    float z = 0.f;
    for (int c = lane; c < cols; c += 32) z += expf(s[c] - m);
    z = warp_sum(z);
    const float inv = 1.f / z;
    for (int c = lane; c < cols; c += 32) dst[r * cols + c] = expf(s[c] - m) * inv;
  }
}

// d theta = W * (dW - sum_j W_j dW_j)
__global__ void softmax_bwd_kernel(const float* __restrict__ W, const float* __restrict__ dW,
                                   float* __restrict__ dsrc, int64_t rows, int cols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < rows; r += (int64_t)gridDim.x * nwarps) {
    const float* w = W + r * cols;
    const float* g = dW + r * cols;
    float dot = 0.f;
    for (int c = lane; c < cols; c += 32) dot = fmaf(w[c], g[c], dot);
    dot = warp_sum(dot);
    for (int c = lane; c < cols; c += 32) dsrc[r * cols + c] = w[c] * (g[c] - dot);
  }
}

// ---- (F,K,V) -> (F,V,K) table builders; one CTA per (fold, unit tile of 32) -----------------
// MODE 0: log_softmax over V, MODE 1: log, MODE 2: copy
template <int MODE>
__global__ void table_fwd_t_kernel(const float* __restrict__ src, float* __restrict__ dst, int K, int V) {
  __shared__ float tile[32][33];
  __shared__ float lse[32];
  const int f = blockIdx.y, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
  const float* s = src + (int64_t)f * K * V;
  float* d = dst + (int64_t)f * V * K;
  if (MODE == 0) {
    // warp ty handles rows k0 + ty, +8, ...: logsumexp over V
    for (int kk = ty; kk < 32; kk += 8) {
      const int k = k0 + kk;
      float m = -INFINITY, z = 0.f;
      if (k < K) {
        for (int v = tx; v < V; v += 32) m = fmaxf(m, s[(int64_t)k * V + v]);
        m = warp_max(m);
        for (int v = tx; v < V; v += 32) z += expf(s[(int64_t)k * V + v] - m);
        z = warp_sum(z);
      }
      if (tx == 0) lse[kk] = (k < K) ? m + logf(z) : 0.f;
    }
    __syncthreads();
  }
  for (int v0 = 0; v0 < V; v0 += 32) {
    for (int kk = ty; kk < 32; kk += 8) {
      const int k = k0 + kk, v = v0 + tx;
      float val = 0.f;
      if (k < K && v < V) {
        val = s[(int64_t)k * V + v];
        if (MODE == 0) val -= lse[kk];
        if (MODE == 1) val = logf(val);
      }
      tile[kk][tx] = val;
    }
    __syncthreads();
    for (int vv = ty; vv < 32; vv += 8) {
      const int v = v0 + vv, k = k0 + tx;
      if (v < V && k < K) d[(int64_t)v * K + k] = tile[tx][vv];
    }
    __syncthreads();
  }
}

// backward: dsrc[f,k,v] from dT[f,v,k]
//  MODE 0: dT[v,k] - exp(T[v,k]) * sum_v dT[v,k]     (T = log softmax)
//  MODE 1: dT[v,k] / src[k,v]
//  MODE 2: dT[v,k]
template <int MODE>
__global__ void table_bwd_t_kernel(const float* __restrict__ src, const float* __restrict__ T,
                                   const float* __restrict__ dT, float* __restrict__ dsrc, int K, int V) {
  __shared__ float tile[32][33];
  __shared__ float tile2[32][33];
  __shared__ float colsum[32];
  const int f = blockIdx.y, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float* t = T + (int64_t)f * V * K;
  const float* g = dT + (int64_t)f * V * K;
  float* d = dsrc + (int64_t)f * K * V;
  if (MODE == 0) {
    // sum over v of dT[v, k0+tx]; slices over ty
    float a = 0.f;
    const int k = k0 + tx;
    if (k < K)
      for (int v = ty; v < V; v += 8) a += g[(int64_t)v * K + k];
    tile[ty][tx] = a;
    __syncthreads();
    if (ty == 0) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += tile[j][tx];
      colsum[tx] = s;
    }
    __syncthreads();
  }
  for (int v0 = 0; v0 < V; v0 += 32) {
    for (int vv = ty; vv < 32; vv += 8) {
      const int v = v0 + vv, k = k0 + tx;
      float gv = 0.f, tv = 0.f;
      if (v < V && k < K) {
        gv = g[(int64_t)v * K + k];
        if (MODE == 0) tv = t[(int64_t)v * K + k];
      }
      tile[vv][tx] = gv;
      if (MODE == 0) tile2[vv][tx] = tv;
    }
    __syncthreads();
    for (int kk = ty; kk < 32; kk += 8) {
      const int k = k0 + kk, v = v0 + tx;
      if (k < K && v < V) {
        float val = tile[tx][kk];
        if (MODE == 0) val -= expf(tile2[tx][kk]) * colsum[kk];
        if (MODE == 1) val /= src[(int64_t)f * K * V + (int64_t)k * V + v];
        d[(int64_t)k * V + v] = val;
      }
    }
    __syncthreads();
  }
}

// ---- the same two ops with the whole (32 units x V states) tile of a fold in shared memory ----
// One CTA per (fold, 32-unit tile): every element is read from global memory ONCE, with all of a
// thread's loads in flight together, and both directions of the transpose go through one padded
// tile (row stride V + 1: row-wise and column-wise accesses are bank-conflict free).  The versions
// above walk the tile in 32 x 32 pieces with two barriers each and re-read the source for the
// log-sum-exp: 45 / 60 us for the 51 MB north-star table against 16 / 24 us of HBM time.
template <int MODE>
__global__ void __launch_bounds__(256) table_fwd_t_smem_kernel(const float* __restrict__ src,
                                                               float* __restrict__ dst, int K, int V) {
  extern __shared__ float tile[];  // [32][V + 1]
  __shared__ float lse[32];
  pdl_wait();
  pdl_launch_dependents();
  const int f = blockIdx.y, k0 = blockIdx.x * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = min(32, K - k0), ld = V + 1;
  const float* s = src + ((int64_t)f * K + k0) * V;  // nk contiguous rows of V floats
  float* d = dst + (int64_t)f * V * K + k0;
  const int64_t n = (int64_t)nk * V;
  if ((V & 3) == 0) {
    for (int64_t i = 4 * (int64_t)tid; i < n; i += 1024) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(s + i));
      const int k = (int)(i / V), v = (int)(i - (int64_t)k * V);
      float* t = tile + k * ld + v;
      t[0] = x.x; t[1] = x.y; t[2] = x.z; t[3] = x.w;
    }
  } else {
    for (int64_t i = tid; i < n; i += 256) tile[(i / V) * ld + (i % V)] = __ldg(s + i);
  }
  __syncthreads();
  if (MODE == 0) {
    for (int k = warp; k < nk; k += 8) {
      float m = -INFINITY, z = 0.f;
      for (int v = lane; v < V; v += 32) m = fmaxf(m, tile[k * ld + v]);
      m = warp_max(m);
      for (int v = lane; v < V; v += 32) z += expf(tile[k * ld + v] - m);
      z = warp_sum(z);
      if (lane == 0) lse[k] = m + logf(z);
    }
    __syncthreads();
  }
  if (lane < nk) {
    const float shift = MODE == 0 ? lse[lane] : 0.f;
    for (int v = warp; v < V; v += 8) {
      float val = tile[lane * ld + v];
      if (MODE == 0) val -= shift;
      if (MODE == 1) val = logf(val);
      d[(int64_t)v * K + lane] = val;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) table_bwd_t_smem_kernel(const float* __restrict__ src,
                                                               const float* __restrict__ T,
                                                               const float* __restrict__ dT,
                                                               float* __restrict__ dsrc, int K, int V) {
  extern __shared__ float tile[];  // g [32][V + 1], then (MODE 0) exp(T) [32][V + 1]
  __shared__ float colsum[32];
  pdl_wait();
  pdl_launch_dependents();
  const int f = blockIdx.y, k0 = blockIdx.x * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = min(32, K - k0), ld = V + 1;
  float* tg = tile;
  float* te = tile + 32 * ld;
  const float* g = dT + (int64_t)f * V * K + k0;
  const float* t = T + (int64_t)f * V * K + k0;
  if (lane < nk) {
    for (int v = warp; v < V; v += 8) {
      tg[lane * ld + v] = __ldg(g + (int64_t)v * K + lane);
      if (MODE == 0) te[lane * ld + v] = __ldg(t + (int64_t)v * K + lane);
    }
  }
  __syncthreads();
  if (MODE == 0) {
    for (int k = warp; k < nk; k += 8) {  // fixed order: lanes stride the states, then the shuffle tree
      float a = 0.f;
      for (int v = lane; v < V; v += 32) a += tg[k * ld + v];
      a = warp_sum(a);
      if (lane == 0) colsum[k] = a;
    }
    __syncthreads();
  }
  float* d = dsrc + ((int64_t)f * K + k0) * V;
  const float* s = src + ((int64_t)f * K + k0) * V;
  for (int k = warp; k < nk; k += 8) {
    const float cs = MODE == 0 ? colsum[k] : 0.f;
    for (int v = lane; v < V; v += 32) {
      float val = tg[k * ld + v];
      if (MODE == 0) val -= expf(te[k * ld + v]) * cs;
      if (MODE == 1) val /= __ldg(s + (int64_t)k * V + v);
      d[(int64_t)k * V + v] = val;
    }
  }
}

// launchers: shared-memory tiles when they fit (V up to ~850 states), else the tiled versions
template <int MODE>
static int launch_table_fwd_t(const float* src, float* dst, int64_t F, int K, int V, cudaStream_t st) {
  const size_t smem = (size_t)32 * (V + 1) * 4;
  dim3 grid(ceil_div(K, 32), (unsigned)F);
  if (smem <= 64 * 1024) {
    static PerDeviceOnce attr;
    if (attr.first())
      CKB_CUDA_CHECK(cudaFuncSetAttribute(table_fwd_t_smem_kernel<MODE>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CKB_CUDA_CHECK(launch_pdl(table_fwd_t_smem_kernel<MODE>, grid, dim3(256), smem, st, src, dst, K, V));
  } else {
    table_fwd_t_kernel<MODE><<<grid, dim3(32, 8), 0, st>>>(src, dst, K, V);
  }
  return CKB_OK;
}
template <int MODE>
static int launch_table_bwd_t(const float* src, const float* T, const float* dT, float* dsrc, int64_t F,
                              int K, int V, cudaStream_t st) {
  const size_t smem = (size_t)(MODE == 0 ? 2 : 1) * 32 * (V + 1) * 4;
  dim3 grid(ceil_div(K, 32), (unsigned)F);
  if (smem <= 72 * 1024) {
    static PerDeviceOnce attr;
    if (attr.first())
      CKB_CUDA_CHECK(cudaFuncSetAttribute(table_bwd_t_smem_kernel<MODE>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    CKB_CUDA_CHECK(launch_pdl(table_bwd_t_smem_kernel<MODE>, grid, dim3(256), smem, st, src, T, dT, dsrc, K, V));
  } else {
    table_bwd_t_kernel<MODE><<<grid, dim3(32, 8), 0, st>>>(src, T, dT, dsrc, K, V);
  }
  return CKB_OK;
}

// ---- elementwise ------------------------------------------------------------------------------
__global__ void scaled_sigmoid_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                          int64_t n, float a, float b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (1.f / (1.f + expf(-src[i]))) * (b - a) + a;
}
__global__ void scaled_sigmoid_bwd_kernel(const float* __restrict__ src, const float* __restrict__ g,
                                          float* __restrict__ dsrc, int64_t n, float a, float b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float s = 1.f / (1.f + expf(-src[i]));
    dsrc[i] = g[i] * (b - a) * s * (1.f - s);
  }
}
__global__ void log_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = logf(src[i]);
}
__global__ void log_bwd_kernel(const float* __restrict__ src, const float* __restrict__ g,
                               float* __restrict__ dsrc, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dsrc[i] = g[i] / src[i];
}
__global__ void lse_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows,
                                int cols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < rows; r += (int64_t)gridDim.x * nwarps) {
    const float* s = src + r * cols;
    float m = -INFINITY, z = 0.f;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, s[c]);
    m = warp_max(m);
    for (int c = lane; c < cols; c += 32) z += expf(s[c] - m);
    z = warp_sum(z);
    if (lane == 0) dst[r] = m + logf(z);
  }
}
// d/dsrc of logsumexp: softmax(src) * g, ADDED to dsrc (the same logits also feed the table op,
// whose backward has already written dsrc when this one runs)
__global__ void lse_rows_bwd_kernel(const float* __restrict__ src, const float* __restrict__ lse,
                                    const float* __restrict__ g, float* __restrict__ dsrc,
                                    int64_t rows, int cols) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    dsrc[i] += g[r] * expf(src[i] - lse[r]);
  }
}

// ---- all softmax ops of a plan in one launch ------------------------------------------------
// Every sum layer re-normalises its weights each step (TorchSoftmaxParameter, nodes.py:764-772);
// a circuit has one such tensor per folded layer, mostly tiny, so they are batched: the op table
// travels in the kernel parameters and a warp finds its (op, row) by a linear scan.
constexpr int kMultiOps = 24;
struct MultiSoftmax {
  int n;
  int all64;  // every op of the batch has 64 columns and 16-byte aligned tensors
  int cols[kMultiOps];
  int64_t row_end[kMultiOps];  // cumulative row counts
  const float* a[kMultiOps];   // fwd: src      bwd: W
  const float* b[kMultiOps];   // fwd: unused   bwd: dW
  float* out[kMultiOps];       // fwd: dst      bwd: dsrc
};

// A warp works on kMsRows rows at once and keeps each row in registers (up to 4 elements per lane,
// rows of at most 128 columns -- every sum layer of width <= 128): one global read per element,
// the loads of all rows in flight together.  Longer rows take the three-pass loop.
constexpr int kMsRows = 4;
template <bool BWD>
__global__ void multi_softmax_kernel(const __grid_constant__ MultiSoftmax m) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t total = m.row_end[m.n - 1];
  if (m.all64) {
    // every row has 64 columns (K = 64 circuits): a half-warp per row, one 16-byte load per lane,
    // 8 rows per warp in flight, 4-step shuffle reductions inside the half-warps
    const int half = lane >> 4, l16 = lane & 15;
    const int64_t stride8 = (int64_t)gridDim.x * nwarps * 8;
    for (int64_t g0 = ((int64_t)blockIdx.x * nwarps + warp) * 8; g0 < total; g0 += stride8) {
      float4 x[4], y[4];
      float4* dst[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t gr = g0 + 2 * q + half;
        dst[q] = nullptr;
        x[q] = BWD ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        y[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < total) {
          int op = 0;
          while (gr >= m.row_end[op]) ++op;
          const int64_t at = (gr - (op ? m.row_end[op - 1] : 0)) * 64 + 4 * l16;
          x[q] = __ldg(reinterpret_cast<const float4*>(m.a[op] + at));
          if (BWD) y[q] = __ldg(reinterpret_cast<const float4*>(m.b[op] + at));
          dst[q] = reinterpret_cast<float4*>(m.out[op] + at);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 r;
        if (!BWD) {
          float mx = fmaxf(fmaxf(x[q].x, x[q].y), fmaxf(x[q].z, x[q].w));
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          r = make_float4(expf(x[q].x - mx), expf(x[q].y - mx), expf(x[q].z - mx), expf(x[q].w - mx));
          float z = (r.x + r.y) + (r.z + r.w);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
          const float inv = 1.f / z;
          r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
        } else {
          float dot = fmaf(x[q].x, y[q].x, fmaf(x[q].y, y[q].y, fmaf(x[q].z, y[q].z, x[q].w * y[q].w)));
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
          r = make_float4(x[q].x * (y[q].x - dot), x[q].y * (y[q].y - dot), x[q].z * (y[q].z - dot),
                          x[q].w * (y[q].w - dot));
        }
        if (dst[q] != nullptr) *dst[q] = r;
      }
    }
    return;
  }
  const int64_t stride = (int64_t)gridDim.x * nwarps * kMsRows;
  for (int64_t g0 = ((int64_t)blockIdx.x * nwarps + warp) * kMsRows; g0 < total; g0 += stride) {
    const float* a[kMsRows];
    const float* g[kMsRows];
    float* out[kMsRows];
    int cols[kMsRows];
    bool small = true;
#pragma unroll
    for (int r = 0; r < kMsRows; ++r) {
      const int64_t gr = g0 + r;
      cols[r] = 0;
      a[r] = g[r] = nullptr;
      out[r] = nullptr;
      if (gr < total) {
        int op = 0;
        while (gr >= m.row_end[op]) ++op;
        const int64_t row = gr - (op ? m.row_end[op - 1] : 0);
        cols[r] = m.cols[op];
        a[r] = m.a[op] + row * cols[r];
        out[r] = m.out[op] + row * cols[r];
        if (BWD) g[r] = m.b[op] + row * cols[r];
        small &= cols[r] <= 128;
      }
    }
    if (small) {
      float x[kMsRows][4], y[kMsRows][4];
#pragma unroll
      for (int r = 0; r < kMsRows; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = lane + 32 * j;
          const bool ok = c < cols[r];
          x[r][j] = ok ? __ldg(a[r] + c) : (BWD ? 0.f : -INFINITY);
          if (BWD) y[r][j] = ok ? __ldg(g[r] + c) : 0.f;
        }
#pragma unroll
      for (int r = 0; r < kMsRows; ++r) {
        if (cols[r] == 0) continue;  // warp-uniform
        if (!BWD) {
          const float mx = warp_max(fmaxf(fmaxf(x[r][0], x[r][1]), fmaxf(x[r][2], x[r][3])));
          float e[4], z = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            e[j] = expf(x[r][j] - mx);  // exp(-inf) = 0 for the padding
            z += e[j];
          }
          const float inv = 1.f / warp_sum(z);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (lane + 32 * j < cols[r]) out[r][lane + 32 * j] = e[j] * inv;
        } else {
          float dot = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) dot = fmaf(x[r][j], y[r][j], dot);
          dot = warp_sum(dot);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (lane + 32 * j < cols[r]) out[r][lane + 32 * j] = x[r][j] * (y[r][j] - dot);
        }
      }
      continue;
    }
    for (int r = 0; r < kMsRows; ++r) {
      if (cols[r] == 0) continue;
      const int n = cols[r];
      if (!BWD) {
        float mx = -INFINITY;
        for (int c = lane; c < n; c += 32) mx = fmaxf(mx, a[r][c]);
        mx = warp_max(mx);
        float z = 0.f;
        for (int c = lane; c < n; c += 32) z += expf(a[r][c] - mx);
        z = warp_sum(z);
        const float inv = 1.f / z;
        for (int c = lane; c < n; c += 32) out[r][c] = expf(a[r][c] - mx) * inv;
      } else {
        float dot = 0.f;
        for (int c = lane; c < n; c += 32) dot = fmaf(a[r][c], g[r][c], dot);
        dot = warp_sum(dot);
        for (int c = lane; c < n; c += 32) out[r][c] = a[r][c] * (g[r][c] - dot);
      }
    }
  }
}

// Long rows (the (Ko, Ki^2) weights of Tucker layers: 4096 columns): one block per row, the row
// stays in registers as 16-byte vectors, so every element is read once and written once.
template <bool BWD, int NV>
__global__ void __launch_bounds__(256) softmax_wide_kernel(const float* __restrict__ a,
                                                          const float* __restrict__ g,
                                                          float* __restrict__ out, int cols) {
  __shared__ float red[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t base = (int64_t)blockIdx.x * cols;
  float4 v[NV], w[NV];
#pragma unroll
  for (int n = 0; n < NV; ++n) {
    const int c = (n * 256 + tid) * 4;
    v[n] = *reinterpret_cast<const float4*>(a + base + c);
    if (BWD) w[n] = *reinterpret_cast<const float4*>(g + base + c);
  }
  auto block_reduce = [&](float x, bool is_max) {
    x = is_max ? warp_max(x) : warp_sum(x);
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
  };
  if (!BWD) {
    float mx = -INFINITY;
#pragma unroll
    for (int n = 0; n < NV; ++n) mx = fmaxf(mx, fmaxf(fmaxf(v[n].x, v[n].y), fmaxf(v[n].z, v[n].w)));
    mx = block_reduce(mx, true);
    float z = 0.f;
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      v[n].x = expf(v[n].x - mx); v[n].y = expf(v[n].y - mx);
      v[n].z = expf(v[n].z - mx); v[n].w = expf(v[n].w - mx);
      z += (v[n].x + v[n].y) + (v[n].z + v[n].w);
    }
    const float inv = 1.f / block_reduce(z, false);
#pragma unroll
    for (int n = 0; n < NV; ++n)
      *reinterpret_cast<float4*>(out + base + (n * 256 + tid) * 4) =
          make_float4(v[n].x * inv, v[n].y * inv, v[n].z * inv, v[n].w * inv);
  } else {
    float dot = 0.f;
#pragma unroll
    for (int n = 0; n < NV; ++n)
      dot += fmaf(v[n].x, w[n].x, v[n].y * w[n].y) + fmaf(v[n].z, w[n].z, v[n].w * w[n].w);
    dot = block_reduce(dot, false);
#pragma unroll
    for (int n = 0; n < NV; ++n)
      *reinterpret_cast<float4*>(out + base + (n * 256 + tid) * 4) =
          make_float4(v[n].x * (w[n].x - dot), v[n].y * (w[n].y - dot), v[n].z * (w[n].z - dot),
                      v[n].w * (w[n].w - dot));
  }
}

template <bool BWD>
static bool softmax_wide(const float* a, const float* g, float* out, int64_t rows, int cols, Ctx& c) {
  if (rows <= 0 || rows > 0x7fffffff) return false;
  switch (cols) {
    case 1024: softmax_wide_kernel<BWD, 1><<<(unsigned)rows, 256, 0, c.stream>>>(a, g, out, cols); break;
    case 2048: softmax_wide_kernel<BWD, 2><<<(unsigned)rows, 256, 0, c.stream>>>(a, g, out, cols); break;
    case 4096: softmax_wide_kernel<BWD, 4><<<(unsigned)rows, 256, 0, c.stream>>>(a, g, out, cols); break;
    default: return false;
  }
  c.launches++;
  return true;
}

// Runs every CKB_POP_SOFTMAX op of `ops` (forward, or backward when `bwd`); returns how many
// launches it made through c.launches.
int multi_softmax(const ckb_param_op_t* ops, int n_ops, bool bwd, Ctx& c) {
  MultiSoftmax m;
  m.n = 0;
  int64_t rows = 0;
  auto flush = [&]() -> int {
    if (m.n == 0) return CKB_OK;
    m.all64 = 1;
    for (int i = 0; i < m.n; ++i)
      if (m.cols[i] != 64 || ((uintptr_t)m.a[i] & 15) || ((uintptr_t)m.out[i] & 15) ||
          (bwd && ((uintptr_t)m.b[i] & 15)))
        m.all64 = 0;
    const int blocks = (int)max64(1, min64(ceil_div(rows, 8 * kMsRows), 8 * kNumSMs));
    if (bwd) CKB_CUDA_CHECK(launch_pdl(multi_softmax_kernel<true>, dim3(blocks), dim3(256), 0, c.stream, m));
    else CKB_CUDA_CHECK(launch_pdl(multi_softmax_kernel<false>, dim3(blocks), dim3(256), 0, c.stream, m));
    c.launches++;
    m.n = 0;
    rows = 0;
    return CKB_OK;
  };
  for (int i = 0; i < n_ops; ++i) {
    const ckb_param_op_t& op = ops[i];
    if (op.kind != CKB_POP_SOFTMAX) continue;
    if (bwd) {
      if (c.grads[op.src] == nullptr) continue;
      if (c.grads[op.dst] == nullptr) {
        set_error("softmax op: gradient of slot %d requested but slot %d has none", op.src, op.dst);
        return CKB_ERR_INVALID;
      }
      if (softmax_wide<true>(c.tensors[op.dst], c.grads[op.dst], c.grads[op.src], op.rows, op.cols, c)) {
        CKB_LAUNCH_CHECK();
        continue;
      }
      m.a[m.n] = c.tensors[op.dst];
      m.b[m.n] = c.grads[op.dst];
      m.out[m.n] = c.grads[op.src];
    } else {
      if (softmax_wide<false>(c.tensors[op.src], nullptr, c.tensors[op.dst], op.rows, op.cols, c)) {
        CKB_LAUNCH_CHECK();
        continue;
      }
      m.a[m.n] = c.tensors[op.src];
      m.b[m.n] = nullptr;
      m.out[m.n] = c.tensors[op.dst];
    }
    m.cols[m.n] = op.cols;
    rows += op.rows;
    m.row_end[m.n] = rows;
    if (++m.n == kMultiOps)
      if (int rc = flush()) return rc;
  }
  return flush();
}

static int grid1d(int64_t n, int per_block) {
  return (int)max64(1, min64(ceil_div(n, per_block), 8 * kNumSMs));
}

int param_op_fwd(const ckb_param_op_t& op, Ctx& c) {
  const float* src = c.tensors[op.src];
  float* dst = c.tensors[op.dst];
  const int64_t n = op.rows * op.cols * (op.aux > 0 ? op.aux : 1);
  switch (op.kind) {
    case CKB_POP_SOFTMAX:
      softmax_fwd_kernel<<<grid1d(op.rows, 8), 256, 0, c.stream>>>(src, dst, op.rows, op.cols);
      break;
    case CKB_POP_LOG_SOFTMAX_T:
      if (int rc = launch_table_fwd_t<0>(src, dst, op.rows, op.aux, op.cols, c.stream)) return rc;
      break;
    case CKB_POP_LOG_T:
      if (int rc = launch_table_fwd_t<1>(src, dst, op.rows, op.aux, op.cols, c.stream)) return rc;
      break;
    case CKB_POP_COPY_T:
      if (int rc = launch_table_fwd_t<2>(src, dst, op.rows, op.aux, op.cols, c.stream)) return rc;
      break;
    case CKB_POP_SCALED_SIGMOID:
      scaled_sigmoid_fwd_kernel<<<grid1d(n, 256), 256, 0, c.stream>>>(src, dst, n, op.a, op.b);
      break;
    case CKB_POP_LOG:
      log_fwd_kernel<<<grid1d(n, 256), 256, 0, c.stream>>>(src, dst, n);
      break;
    case CKB_POP_LSE_ROWS:
      lse_rows_kernel<<<grid1d(op.rows, 8), 256, 0, c.stream>>>(src, dst, op.rows, op.cols);
      break;
    default:
      set_error("unknown parameter op %d", op.kind);
      return CKB_ERR_INVALID;
  }
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

int param_op_bwd(const ckb_param_op_t& op, Ctx& c) {
  float* dsrc = c.grads[op.src];
  const float* g = c.grads[op.dst];
  if (dsrc == nullptr) return CKB_OK;
  if (op.kind == CKB_POP_LSE_ROWS && g == nullptr) return CKB_OK;  // value used, gradient not wanted
  if (g == nullptr) {
    set_error("parameter op %d: gradient of slot %d requested but slot %d has none", op.kind,
              op.src, op.dst);
    return CKB_ERR_INVALID;
  }
  const float* src = c.tensors[op.src];
  const float* dst = c.tensors[op.dst];
  const int64_t n = op.rows * op.cols * (op.aux > 0 ? op.aux : 1);
  switch (op.kind) {
    case CKB_POP_SOFTMAX:
      softmax_bwd_kernel<<<grid1d(op.rows, 8), 256, 0, c.stream>>>(dst, g, dsrc, op.rows, op.cols);
      break;
    case CKB_POP_LOG_SOFTMAX_T:
      if (int rc = launch_table_bwd_t<0>(src, dst, g, dsrc, op.rows, op.aux, op.cols, c.stream)) return rc;
      break;
    case CKB_POP_LOG_T:
      if (int rc = launch_table_bwd_t<1>(src, dst, g, dsrc, op.rows, op.aux, op.cols, c.stream)) return rc;
      break;
    case CKB_POP_COPY_T:
      if (int rc = launch_table_bwd_t<2>(src, dst, g, dsrc, op.rows, op.aux, op.cols, c.stream)) return rc;
      break;
    case CKB_POP_SCALED_SIGMOID:
      scaled_sigmoid_bwd_kernel<<<grid1d(n, 256), 256, 0, c.stream>>>(src, g, dsrc, n, op.a, op.b);
      break;
    case CKB_POP_LOG:
      log_bwd_kernel<<<grid1d(n, 256), 256, 0, c.stream>>>(src, g, dsrc, n);
      break;
    case CKB_POP_LSE_ROWS:
      lse_rows_bwd_kernel<<<grid1d(n, 256), 256, 0, c.stream>>>(src, dst, g, dsrc, op.rows, op.cols);
      break;
    default:
      set_error("unknown parameter op %d", op.kind);
      return CKB_ERR_INVALID;
  }
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

}  // namespace ckb
