// 'complex-lse-sum' semiring in the plan executor (BASELINE.json configs[4]: squared circuits).
//
// Reference semantics (cirkit/backend/torch):
//   * activations are complex logarithms; here they are interleaved (re, im) float pairs in the
//     same (fold, batch, unit) arena blocks as the real path, so a block of step s starts at float
//     offset 2 * B * out_off[s] (offsets stay in units per sample);
//   * ComplexLSESumSemiring.apply_reduce, semiring.py:440-476: m = clamp(max_i Re u_i),
//     e_i = exp(u_i - m), S_o = sum_i W[o,i] e_i, y_o = csafelog(S_o) + m;
//   * csafelog, utils.py:32-50: log z forward, nan_to_num(g / conj(z)) backward;
//   * TorchEmbeddingLayer.forward layers/input.py:258-266: y = csafelog(W[f,:,x]); a Categorical
//     layer under this semiring is its real log-probability cast to complex (semiring.py:511-514);
//   * TorchConjugateParameter nodes.py:742-746 (CKB_POP_CONJ);
//   * gradients follow PyTorch's convention for complex tensors (stored gradient = dL/d conj z):
//     through a holomorphic f, g_in = g_out conj(f'(z)); through z = w e, g_w = g_z conj(e).
// Kernels are FP32 SIMT; weight gradients are accumulated in a fixed order (thread = weight
// entries, loop over the CTA's samples; per-CTA slabs + reduce_partials): no atomics.
#include "common.cuh"

namespace ckb {
namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
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
// nan_to_num(g / conj(z)), the quotient as c10::complex computes it
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

struct CGradSrc {
  const float2* garena;
  const int32_t* cons_ptr;
  const int64_t* cons_rows;
  int64_t B;
};
__device__ __forceinline__ float2 cpull(const CGradSrc& g, int f, int64_t b, int K, int k) {
  float2 acc = make_float2(0.f, 0.f);
  const int c1 = g.cons_ptr[f + 1];
  for (int c = g.cons_ptr[f]; c < c1; ++c) {
    const float2 v = g.garena[g.B * g.cons_rows[c] + b * K + k];
    acc.x += v.x;
    acc.y += v.y;
  }
  return acc;
}

// ---------------------------------------------------------------------------------- tables
// MODE 0: Embedding, W complex (F, K, V): y = csafelog(W[f, k, x]).
// MODE 1: real log-table T (F, V, K) produced by the real parameter ops: y = (T[f, x, k], 0).
template <int MODE>
__global__ void ctable_fwd_kernel(const float* __restrict__ T, const int32_t* __restrict__ scope_var,
                                  const void* __restrict__ xT, int x_is_float, float2* __restrict__ y,
                                  int64_t B, int K, int V) {
  const int f = blockIdx.y;
  const int var = scope_var[f];
  const int64_t total = B * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / K;
    const int k = (int)(idx - b * K);
    int v = read_state(xT, x_is_float, (int64_t)var * B + b);
    v = min(max(v, 0), V - 1);
    float2 out;
    if (MODE == 0)
      out = clog(reinterpret_cast<const float2*>(T)[((int64_t)f * K + k) * V + v]);
    else
      out = make_float2(T[((int64_t)f * V + v) * K + k], 0.f);
    y[(int64_t)f * total + idx] = out;
  }
}

// One CTA per (fold, batch chunk), one thread per unit k: the thread walks the chunk's samples in
// order and adds to ITS row of a shared-memory gradient table (no conflicts, fixed order).
template <int MODE>
__global__ void ctable_bwd_kernel(CGradSrc gs, const float* __restrict__ T,
                                  const int32_t* __restrict__ scope_var, const void* __restrict__ xT,
                                  int x_is_float, float* __restrict__ out, int64_t B, int K, int V,
                                  int64_t chunk) {
  extern __shared__ float smem_t[];  // MODE 0: [K][V] complex, MODE 1: [V][K] real
  const int f = blockIdx.y;
  const int var = scope_var[f];
  const int64_t b0 = (int64_t)blockIdx.x * chunk, b1 = min64(B, b0 + chunk);
  const int n = (MODE == 0 ? 2 : 1) * K * V;
  for (int i = threadIdx.x; i < n; i += blockDim.x) smem_t[i] = 0.f;
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    for (int64_t b = b0; b < b1; ++b) {
      int v = read_state(xT, x_is_float, (int64_t)var * B + b);
      v = min(max(v, 0), V - 1);
      const float2 g = cpull(gs, f, b, K, k);
      if (MODE == 0) {
        const float2 w = reinterpret_cast<const float2*>(T)[((int64_t)f * K + k) * V + v];
        const float2 d = safe_div_conj(g, w);
        smem_t[2 * (k * V + v)] += d.x;
        smem_t[2 * (k * V + v) + 1] += d.y;
      } else {
        smem_t[v * K + k] += g.x;  // real table under a complex cast: the real part of the gradient
      }
    }
  }
  __syncthreads();
  float* o = out + ((int64_t)blockIdx.x * gridDim.y + f) * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = smem_t[i];
}

// ---------------------------------------------------------------------------------- sum layers
// One CTA per (fold, batch chunk), 8 warps; the fold's weights sit in shared memory.  Reduction
// index: the H inputs are summed (Hadamard-fused, CPT / arity-1 sums), or -- TUCKER -- the
// reduction runs over the pairs (i, j) of two inputs, e = e1_i e2_j, one shift per input.
constexpr int kCWarps = 8;

struct CDenseArgs {
  const float2* W;          // (F, Ko, Kred)
  const int64_t* in_rows;   // (F*H) per-sample unit offsets
  const float2* arena;
  float2* y;                // (F, B, Ko)
  int64_t B, chunk;
  int H, Ki, Ko, Kred, tucker;
  int Kq;  // TENSORDOT (layers/optimized.py:205-300): every sample row holds Kq interleaved vectors --
           // input element i of vector q sits at i*Kq + q, output o at q*Ko + o; 1 for the sum layers
  CGradSrc gs;
  float2* gin;              // (F, gin_h, B, Ki)
  float* dWp;               // [chunks][F][Ko][Kred] complex, or nullptr
};

// e[0..Kred) of one sample into shared memory; returns the shift (sum of the per-input shifts)
__device__ __forceinline__ float cload_e(const CDenseArgs& a, int f, int64_t s, int lane, float2* e,
                                         float2* e12) {
  const int64_t b = s / a.Kq;
  const int q = (int)(s - b * a.Kq);
  if (!a.tucker) {
    float m = -INFINITY;
    for (int i = lane; i < a.Ki; i += 32) {
      float2 u = make_float2(0.f, 0.f);
      for (int h = 0; h < a.H; ++h) {
        const float2 v = a.arena[a.B * a.in_rows[f * a.H + h] + b * a.Ki * a.Kq + i * a.Kq + q];
        u.x += v.x;
        u.y += v.y;
      }
      e[i] = u;
      m = fmaxf(m, u.x);
    }
    m = clamp_max(warp_max(m));
    __syncwarp();
    for (int i = lane; i < a.Ki; i += 32) e[i] = cexp_shifted(e[i], m);
    __syncwarp();
    return m;
  }
  float ms[2];
  for (int h = 0; h < 2; ++h) {
    const float2* x = a.arena + a.B * a.in_rows[f * 2 + h] + b * a.Ki;
    float m = -INFINITY;
    for (int i = lane; i < a.Ki; i += 32) m = fmaxf(m, x[i].x);
    m = clamp_max(warp_max(m));
    for (int i = lane; i < a.Ki; i += 32) e12[h * a.Ki + i] = cexp_shifted(x[i], m);
    ms[h] = m;
  }
  __syncwarp();
  for (int ij = lane; ij < a.Kred; ij += 32) e[ij] = cmul(e12[ij / a.Ki], e12[a.Ki + ij % a.Ki]);
  __syncwarp();
  return ms[0] + ms[1];
}

__global__ void __launch_bounds__(kCWarps * 32) cdense_fwd_kernel(CDenseArgs a) {
  extern __shared__ float2 smem_d[];
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* Ws = smem_d;                                   // [Ko][Kred + 1]
  float2* e = Ws + a.Ko * (a.Kred + 1) + warp * (a.Kred + 2 * a.Ki);
  float2* e12 = e + a.Kred;
  const float2* Wf = a.W + (int64_t)f * a.Ko * a.Kred;
  for (int i = threadIdx.x; i < a.Ko * a.Kred; i += blockDim.x) Ws[(i / a.Kred) * (a.Kred + 1) + i % a.Kred] = Wf[i];
  __syncthreads();
  const int KP = a.Kred + 1;  // padded row: lanes (= rows o) hit different banks
  const int64_t b0 = (int64_t)blockIdx.x * a.chunk, b1 = min64(a.B * a.Kq, b0 + a.chunk);
  for (int64_t b = b0 + warp; b < b1; b += kCWarps) {
    const float m = cload_e(a, f, b, lane, e, e12);
    for (int o = lane; o < a.Ko; o += 32) {
      const float2* wr = Ws + o * KP;
      float2 s = make_float2(0.f, 0.f);
      for (int i = 0; i < a.Kred; ++i) {
        const float2 p = cmul(wr[i], e[i]);
        s.x += p.x;
        s.y += p.y;
      }
      float2 l = clog(s);
      l.x += m;
      a.y[(int64_t)f * a.B * a.Ko * a.Kq + b * a.Ko + o] = l;  // row b = (sample, q): q*Ko + o
    }
    __syncwarp();
  }
}

// NE: weight-gradient entries per thread (Ko * Kred <= 256 * NE)
template <int NE>
__global__ void __launch_bounds__(kCWarps * 32) cdense_bwd_kernel(CDenseArgs a) {
  extern __shared__ float2 smem_d[];
  const int f = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nW = a.Ko * a.Kred;
  const int KP = a.Kred + 1;
  float2* Ws = smem_d;                                  // [Ko][Kred + 1]
  float2* es = Ws + a.Ko * KP;                          // [warps][Kred]
  float2* rs = es + kCWarps * a.Kred;                   // [warps][Ko]
  float2* e12 = rs + kCWarps * a.Ko + warp * 2 * a.Ki;  // [warps][2 Ki]  (tucker)
  float2* ge = e12 + (kCWarps - warp) * 2 * a.Ki + warp * a.Kred;  // [warps][Kred] (tucker)
  float2* e = es + warp * a.Kred;
  float2* r = rs + warp * a.Ko;
  const float2* Wf = a.W + (int64_t)f * nW;
  for (int i = threadIdx.x; i < nW; i += blockDim.x) Ws[(i / a.Kred) * KP + i % a.Kred] = Wf[i];
  float2 acc[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j) acc[j] = make_float2(0.f, 0.f);
  __syncthreads();
  const int64_t b0 = (int64_t)blockIdx.x * a.chunk, b1 = min64(a.B * a.Kq, b0 + a.chunk);
  for (int64_t bb = b0; bb < b1; bb += kCWarps) {
    const int64_t b = bb + warp;  // row (sample, q)
    const bool live = b < b1;
    const int64_t smp = b / a.Kq;
    const int q = (int)(b - smp * a.Kq);
    if (live) {
      const float m = cload_e(a, f, b, lane, e, e12);
      for (int o = lane; o < a.Ko; o += 32) {
        // S_o = exp(y_o - m): the sum the forward took the logarithm of
        const float2 s = cexp_shifted(a.y[(int64_t)f * a.B * a.Ko * a.Kq + b * a.Ko + o], m);
        r[o] = safe_div_conj(cpull(a.gs, f, smp, a.Ko * a.Kq, q * a.Ko + o), s);
      }
      __syncwarp();
      if (!a.tucker) {
        for (int i = lane; i < a.Ki; i += 32) {
          float2 g = make_float2(0.f, 0.f);
          for (int o = 0; o < a.Ko; ++o) {
            const float2 p = cmul_conj(r[o], Ws[o * KP + i]);
            g.x += p.x;
            g.y += p.y;
          }
          // every input of the fold receives the same gradient (u is their sum)
          a.gin[((int64_t)f * a.B + smp) * a.Ki * a.Kq + i * a.Kq + q] = cmul_conj(g, e[i]);
        }
      } else {
        for (int ij = lane; ij < a.Kred; ij += 32) {
          float2 g = make_float2(0.f, 0.f);
          for (int o = 0; o < a.Ko; ++o) {
            const float2 p = cmul_conj(r[o], Ws[o * KP + ij]);
            g.x += p.x;
            g.y += p.y;
          }
          ge[ij] = g;  // d/d(e1_i e2_j)
        }
        __syncwarp();
        for (int i = lane; i < a.Ki; i += 32) {
          float2 g1 = make_float2(0.f, 0.f), g2 = make_float2(0.f, 0.f);
          for (int j = 0; j < a.Ki; ++j) {
            const float2 p = cmul_conj(ge[i * a.Ki + j], e12[a.Ki + j]);  // d/de1_i
            g1.x += p.x;
            g1.y += p.y;
            const float2 q = cmul_conj(ge[j * a.Ki + i], e12[j]);         // d/de2_i
            g2.x += q.x;
            g2.y += q.y;
          }
          a.gin[(((int64_t)f * 2 + 0) * a.B + b) * a.Ki + i] = cmul_conj(g1, e12[i]);
          a.gin[(((int64_t)f * 2 + 1) * a.B + b) * a.Ki + i] = cmul_conj(g2, e12[a.Ki + i]);
        }
      }
    } else {  // no sample for this warp in the last group: its slots must not contribute
      for (int o = lane; o < a.Ko; o += 32) r[o] = make_float2(0.f, 0.f);
      for (int i = lane; i < a.Kred; i += 32) e[i] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    if (a.dWp) {
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        const int idx = threadIdx.x + j * (kCWarps * 32);
        if (idx < nW) {
          const int o = idx / a.Kred, i = idx - o * a.Kred;
#pragma unroll
          for (int w = 0; w < kCWarps; ++w) {  // fixed order over the group's samples
            const float2 p = cmul_conj(rs[w * a.Ko + o], es[w * a.Kred + i]);
            acc[j].x += p.x;
            acc[j].y += p.y;
          }
        }
      }
    }
    __syncthreads();
  }
  if (a.dWp) {
    float2* out = reinterpret_cast<float2*>(a.dWp) + ((int64_t)blockIdx.x * gridDim.y + f) * nW;
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      const int idx = threadIdx.x + j * (kCWarps * 32);
      if (idx < nW) out[idx] = acc[j];
    }
  }
}

// ---------------------------------------------------------------------------------- products
__global__ void chadamard_fwd_kernel(const float2* __restrict__ arena, const int64_t* __restrict__ in_rows,
                                     float2* __restrict__ y, int64_t B, int H, int K) {
  const int f = blockIdx.y;
  const int64_t total = B * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    float2 s = make_float2(0.f, 0.f);
    for (int h = 0; h < H; ++h) {
      const float2 v = arena[B * in_rows[f * H + h] + idx];
      s.x += v.x;
      s.y += v.y;
    }
    y[(int64_t)f * total + idx] = s;
  }
}
__global__ void chadamard_bwd_kernel(CGradSrc gs, float2* __restrict__ gin, int64_t B, int K) {
  const int f = blockIdx.y;
  const int64_t total = B * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / K;
    gin[(int64_t)f * total + idx] = cpull(gs, f, b, K, (int)(idx - b * K));
  }
}

// ---------------------------------------------------------------------------------- constants
// value (F, K) complex, already in log space (the runtime applies csafelog on the host for
// linear-space constants: a batch-free (F, K) tensor)
__global__ void cconstant_fwd_kernel(const float2* __restrict__ value, float2* __restrict__ y,
                                     int64_t B, int K) {
  const int f = blockIdx.y;
  const int64_t total = B * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    y[(int64_t)f * total + idx] = value[(int64_t)f * K + (int)(idx % K)];
}
__global__ void cconstant_bwd_kernel(CGradSrc gs, float2* __restrict__ dvalue, int64_t B, int K) {
  const int f = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float2 a = make_float2(0.f, 0.f);
    for (int64_t b = 0; b < B; ++b) {
      const float2 g = cpull(gs, f, b, K, k);
      a.x += g.x;
      a.y += g.y;
    }
    dvalue[(int64_t)f * K + k] = a;
  }
}

__global__ void conj_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = make_float2(src[i].x, -src[i].y);
}

size_t cdense_smem(int Ki, int Ko, int Kred, bool bwd) {
  size_t n = (size_t)Ko * (Kred + 1);
  if (!bwd) n += (size_t)kCWarps * (Kred + 2 * Ki);
  else n += (size_t)kCWarps * (Kred + Ko + 2 * Ki + Kred);
  return n * sizeof(float2);
}

int grid_x(int64_t n, int per_block, int cap) { return (int)max64(1, min64(ceil_div(n, per_block), cap)); }

}  // namespace

static void cdense_config(int F, int64_t B, int& chunks, int64_t& chunk) {
  chunks = (int)max64(1, min64(ceil_div(B, 4 * kCWarps), ceil_div(2 * kNumSMs, F)));
  chunk = (int64_t)ceil_div(ceil_div(B, chunks), kCWarps) * kCWarps;
  chunks = ceil_div(B, chunk);
}
static void ctable_config(int F, int64_t B, int& chunks, int64_t& chunk) {
  chunks = (int)max64(1, min64(ceil_div(B, 64), ceil_div(2 * kNumSMs, F)));
  chunk = ceil_div(B, chunks);
  chunks = ceil_div(B, chunk);
}

// TENSORDOT steps: k_in = Kj * Kq and k_out = Kq * Kk; num_states carries Kq
static int tdot_kq(const ckb_step_desc_t& d) { return d.kind == CKB_STEP_TENSORDOT ? d.num_states : 1; }

static CDenseArgs cdense_args(const ckb_step_desc_t& d, Ctx& c) {
  CDenseArgs a{};
  a.W = reinterpret_cast<const float2*>(c.tensors[d.slot[0]]);
  a.in_rows = d.in_rows;
  a.arena = reinterpret_cast<const float2*>(c.arena);
  a.y = reinterpret_cast<float2*>(c.arena) + c.B * d.out_off;
  a.B = c.B;
  a.H = d.arity;
  a.Kq = tdot_kq(d);
  a.Ki = d.k_in / a.Kq;
  a.Ko = d.k_out / a.Kq;
  a.tucker = d.kind == CKB_STEP_TUCKER;
  a.Kred = a.tucker ? a.Ki * a.Ki : a.Ki;
  return a;
}

static int cdense_check(const ckb_step_desc_t& d) {
  if (d.flags & CKB_DENSE_CONCAT) {
    set_error("complex semiring: sum layers over concatenated inputs have no kernel");
    return CKB_ERR_UNSUPPORTED;
  }
  const int kq = tdot_kq(d);
  if (kq <= 0 || d.k_in % kq || d.k_out % kq) {
    set_error("complex tensordot: %d / %d units are not multiples of Kq = %d", d.k_in, d.k_out, kq);
    return CKB_ERR_INVALID;
  }
  const int Ki = d.k_in / kq, Ko = d.k_out / kq;
  const int Kred = d.kind == CKB_STEP_TUCKER ? Ki * Ki : Ki;
  if ((int64_t)Ko * Kred > 256 * 16 || cdense_smem(Ki, Ko, Kred, true) > 200 * 1024) {
    set_error("complex semiring: %d x %d weights exceed the kernels' shared-memory tiles", Ko, Kred);
    return CKB_ERR_UNSUPPORTED;
  }
  if (d.kind == CKB_STEP_TUCKER && d.arity != 2) {
    set_error("complex semiring: tucker arity %d", d.arity);
    return CKB_ERR_UNSUPPORTED;
  }
  return CKB_OK;
}

size_t complex_step_ws(const ckb_step_desc_t& d, int64_t B) {
  int chunks;
  int64_t chunk;
  switch (d.kind) {
    case CKB_STEP_TABLE:
      ctable_config(d.num_folds, B, chunks, chunk);
      return (size_t)chunks * d.num_folds * d.k_out * d.num_states * 8;
    case CKB_STEP_DENSE:
    case CKB_STEP_TENSORDOT:
    case CKB_STEP_TUCKER: {
      const int kq = tdot_kq(d) > 0 ? tdot_kq(d) : 1;
      cdense_config(d.num_folds, B * kq, chunks, chunk);
      const int Ki = d.k_in / kq, Ko = d.k_out / kq;
      const int Kred = d.kind == CKB_STEP_TUCKER ? Ki * Ki : Ki;
      return (size_t)chunks * d.num_folds * Ko * Kred * 8;
    }
    default: return 0;
  }
}

int complex_step_fwd(const ckb_step_desc_t& d, Ctx& c) {
  float2* arena = reinterpret_cast<float2*>(c.arena);
  float2* y = arena + c.B * d.out_off;
  const int F = d.num_folds, K = d.k_out;
  switch (d.kind) {
    case CKB_STEP_TABLE: {
      dim3 grid(grid_x(c.B * K, 256, 4 * kNumSMs), F);
      if (d.flags & CKB_STEP_REAL_TABLE)
        ctable_fwd_kernel<1><<<grid, 256, 0, c.stream>>>(c.tensors[d.slot[0]], d.scope_var, c.xT,
                                                        c.x_is_float, y, c.B, K, d.num_states);
      else
        ctable_fwd_kernel<0><<<grid, 256, 0, c.stream>>>(c.tensors[d.slot[0]], d.scope_var, c.xT,
                                                        c.x_is_float, y, c.B, K, d.num_states);
      break;
    }
    case CKB_STEP_CONSTANT: {
      dim3 grid(grid_x(c.B * K, 256, 4 * kNumSMs), F);
      cconstant_fwd_kernel<<<grid, 256, 0, c.stream>>>(
          reinterpret_cast<const float2*>(c.tensors[d.slot[0]]), y, c.B, K);
      break;
    }
    case CKB_STEP_HADAMARD: {
      dim3 grid(grid_x(c.B * K, 256, 4 * kNumSMs), F);
      chadamard_fwd_kernel<<<grid, 256, 0, c.stream>>>(arena, d.in_rows, y, c.B, d.arity, K);
      break;
    }
    case CKB_STEP_DENSE:
    case CKB_STEP_TENSORDOT:
    case CKB_STEP_TUCKER: {
      if (int rc = cdense_check(d)) return rc;
      CDenseArgs a = cdense_args(d, c);
      int chunks;
      cdense_config(F, c.B * a.Kq, chunks, a.chunk);
      const size_t smem = cdense_smem(a.Ki, a.Ko, a.Kred, false);
      static PerDeviceOnce attr;
      if (attr.first())
        CKB_CUDA_CHECK(cudaFuncSetAttribute(cdense_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            200 * 1024));
      cdense_fwd_kernel<<<dim3(chunks, F), kCWarps * 32, smem, c.stream>>>(a);
      break;
    }
    default:
      set_error("complex semiring: step kind %d has no kernel", d.kind);
      return CKB_ERR_UNSUPPORTED;
  }
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

template <int NE>
static int launch_cdense_bwd(const CDenseArgs& a, int chunks, int F, size_t smem, Ctx& c) {
  static PerDeviceOnce attr;
  if (attr.first())
    CKB_CUDA_CHECK(cudaFuncSetAttribute(cdense_bwd_kernel<NE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        200 * 1024));
  cdense_bwd_kernel<NE><<<dim3(chunks, F), kCWarps * 32, smem, c.stream>>>(a);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

int complex_step_bwd(const ckb_step_desc_t& d, Ctx& c) {
  const int F = d.num_folds, K = d.k_out;
  CGradSrc gs{reinterpret_cast<const float2*>(c.garena), d.cons_ptr, d.cons_rows, c.B};
  float2* gin = d.gin_off >= 0 ? reinterpret_cast<float2*>(c.garena) + c.B * d.gin_off : nullptr;
  switch (d.kind) {
    case CKB_STEP_TABLE: {
      float* dT = c.grads[d.slot[0]];
      if (!dT) return CKB_OK;
      const bool real = (d.flags & CKB_STEP_REAL_TABLE) != 0;
      int chunks;
      int64_t chunk;
      ctable_config(F, c.B, chunks, chunk);
      const size_t n = (size_t)(real ? 1 : 2) * F * K * d.num_states;
      const size_t smem = (size_t)(real ? 1 : 2) * K * d.num_states * 4;
      if (smem > 200 * 1024) {
        set_error("complex semiring: a %d x %d table does not fit shared memory", K, d.num_states);
        return CKB_ERR_UNSUPPORTED;
      }
      float* out = dT;
      if (chunks > 1) {
        if (c.ws_bytes < (size_t)chunks * n * 4) {
          set_error("complex table backward: workspace too small");
          return CKB_ERR_WORKSPACE;
        }
        out = (float*)c.ws;
      }
      static PerDeviceOnce attr;
      if (attr.first()) {
        CKB_CUDA_CHECK(cudaFuncSetAttribute(ctable_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CKB_CUDA_CHECK(cudaFuncSetAttribute(ctable_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      }
      const int threads = min(256, (K + 31) / 32 * 32);
      if (real)
        ctable_bwd_kernel<1><<<dim3(chunks, F), threads, smem, c.stream>>>(
            gs, c.tensors[d.slot[0]], d.scope_var, c.xT, c.x_is_float, out, c.B, K, d.num_states, chunk);
      else
        ctable_bwd_kernel<0><<<dim3(chunks, F), threads, smem, c.stream>>>(
            gs, c.tensors[d.slot[0]], d.scope_var, c.xT, c.x_is_float, out, c.B, K, d.num_states, chunk);
      CKB_LAUNCH_CHECK();
      c.launches++;
      if (chunks > 1) return reduce_partials(out, dT, (int64_t)n, chunks, c);
      return CKB_OK;
    }
    case CKB_STEP_CONSTANT: {
      float* dv = c.grads[d.slot[0]];
      if (!dv) return CKB_OK;
      cconstant_bwd_kernel<<<F, 128, 0, c.stream>>>(gs, reinterpret_cast<float2*>(dv), c.B, K);
      CKB_LAUNCH_CHECK();
      c.launches++;
      return CKB_OK;
    }
    case CKB_STEP_HADAMARD: {
      dim3 grid(grid_x(c.B * K, 256, 4 * kNumSMs), F);
      chadamard_bwd_kernel<<<grid, 256, 0, c.stream>>>(gs, gin, c.B, K);
      CKB_LAUNCH_CHECK();
      c.launches++;
      return CKB_OK;
    }
    case CKB_STEP_DENSE:
    case CKB_STEP_TENSORDOT:
    case CKB_STEP_TUCKER: {
      if (int rc = cdense_check(d)) return rc;
      CDenseArgs a = cdense_args(d, c);
      a.gs = gs;
      a.gin = gin;
      int chunks;
      cdense_config(F, c.B * a.Kq, chunks, a.chunk);
      float* dW = c.grads[d.slot[0]];
      const size_t n = (size_t)2 * F * a.Ko * a.Kred;
      a.dWp = dW;
      if (dW && chunks > 1) {
        if (c.ws_bytes < (size_t)chunks * n * 4) {
          set_error("complex dense backward: workspace too small");
          return CKB_ERR_WORKSPACE;
        }
        a.dWp = (float*)c.ws;
      }
      const size_t smem = cdense_smem(a.Ki, a.Ko, a.Kred, true);
      const int ne = ceil_div((int64_t)a.Ko * a.Kred, kCWarps * 32);
      int rc;
      if (ne <= 1) rc = launch_cdense_bwd<1>(a, chunks, F, smem, c);
      else if (ne <= 4) rc = launch_cdense_bwd<4>(a, chunks, F, smem, c);
      else rc = launch_cdense_bwd<16>(a, chunks, F, smem, c);
      if (rc) return rc;
      if (dW && chunks > 1) return reduce_partials(a.dWp, dW, (int64_t)n, chunks, c);
      return CKB_OK;
    }
    default:
      set_error("complex semiring: step kind %d has no kernel", d.kind);
      return CKB_ERR_UNSUPPORTED;
  }
}

int complex_conj(const float* src, float* dst, int64_t n, Ctx& c) {
  conj_kernel<<<grid_x(n, 256, 8 * kNumSMs), 256, 0, c.stream>>>(reinterpret_cast<const float2*>(src),
                                                                reinterpret_cast<float2*>(dst), n);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

}  // namespace ckb
