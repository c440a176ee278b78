// The fused sum-product block for Ki = Ko = 128 (BASELINE.json configs[3]) on tcgen05, TS form:
// the generated operand (e = exp(u - m) forward, r / r^T backward) lives in TMEM, the other operand
// in shared memory.  Forward first; the two backward kernels follow further down.
//
// Written blind at the end of round 1, validated on the B200 in round 2
// (tests/test_gpu_zzz_dense128.py: forward and gradients against the SIMT route and the fp64 oracle)
// and dispatched by default since (bit 9 of CKB_OPT_TC_FAST_MATH, on in the default value).
//
//   y[b,o] = log( sum_i W[o,i] exp(u[b,i] - m[b]) ) + m[b],  u = x_0 (+ x_1),  m = max_i u
//   (TorchCPTLayer.forward layers/optimized.py:171-178 / TorchSumLayer.forward
//   layers/inner.py:266-273 through LSESumSemiring.apply_reduce semiring.py:382-408)
//
// CTA = (fold, a run of 128-sample tiles), 16 warps.  Thread = (sample row q*32 + lane, column
// group cg of 32 units): it loads its 32 pre-activations, the four column groups of a row combine
// their maxima through shared memory, e is split into (hi, lo) and written to TMEM with
// tcgen05.st, warp 0 issues 16 k-steps of  e_hi x [W_hi | W_lo] (N = 256: main | correction)  and
// e_lo x W_hi (N = 128: correction), and every thread reads its 32 outputs back for the log
// epilogue.  TMEM: e_hi [0,128) e_lo [128,256) main [256,384) correction [384,512) -- all 512
// columns, one CTA per SM.  Phases of a tile are serial (no room to double-buffer); the
// feasibility arithmetic is in DESIGN.md section 8, item 4.
#include "dense.cuh"
#include "sm100.cuh"
#include "tc_util.cuh"

namespace ckb {
using namespace sm100;

namespace {

constexpr int TM = 128;   // samples per tile (UMMA M)
constexpr int K128 = 128; // Ki = Ko
constexpr int kThreads128 = 512;
constexpr uint32_t kWBlk = 256 * 128;  // bytes of one k-block: [hi o 0..127 | lo o 0..127] x 32 fp32

struct __align__(1024) Fwd128Smem {
  float w[4][256 * 32];  // [k-block][hi rows | lo rows][32], 128-byte swizzle        128 KB
  float part[4][TM];     // per column group row maxima
  uint64_t a_full, d_full;
  uint32_t tmem_base;
};

template <bool FAST>
__global__ void __launch_bounds__(kThreads128, 1)
dense128_tc_fwd_kernel(DenseArgs a, int tiles_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  Fwd128Smem& s = *reinterpret_cast<Fwd128Smem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)((a.B + TM - 1) / TM);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(n_tiles_total, t_begin + tiles_per_cta) - t_begin;
  if (n_tiles <= 0) return;

  const float* row0 = in_row(a, f, 0);
  const float* row1 = a.H == 2 ? in_row(a, f, 1) : nullptr;

  if (tid == 0) {
    mbar_init(&s.a_full, kThreads128 / 32);
    mbar_init(&s.d_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);

  // ---- the fold's weights -> (hi | lo) swizzled K-major image; loads first, stores after
  {
    const float* Wf = a.W + (int64_t)f * K128 * K128;
    const uint32_t wbase = smem_u32(s.w);
    constexpr int PER = K128 * K128 / 4 / kThreads128;  // 8 float4 per thread
    float4 v[PER];
#pragma unroll
    for (int n = 0; n < PER; ++n) v[n] = __ldg(reinterpret_cast<const float4*>(Wf) + tid + n * kThreads128);
#pragma unroll
    for (int n = 0; n < PER; ++n) {
      const int p = tid + n * kThreads128;
      const int o = p >> 5, c4 = p & 31;  // row o, float4 number c4 of its 32
      const uint32_t kb = c4 >> 3, chunk = c4 & 7;
      float4 hi, lo;
      split4(v[n], hi, lo);
      const uint32_t off = kb * kWBlk + (uint32_t)o * 128u + (((chunk ^ (uint32_t)o) & 7u) << 4);
      sts128(wbase + off, hi);
      sts128(wbase + off + 128u * 128u, lo);  // rows 128 + o: same swizzle phase
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  const int q = warp & 3, cg = warp >> 2;
  const int row = q * 32 + lane;
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  constexpr uint32_t kColLo = 128, kColMain = 256, kColCorr = 384;

  for (int it = 0; it < n_tiles; ++it) {
    const int64_t b = (int64_t)(t_begin + it) * TM + row;
    const bool valid = b < a.B;
    const int64_t off = (valid ? b : 0) * K128 + cg * 32;
    // ---- this thread's 32 pre-activations
    float u[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 v = ldg_stream(row0 + off + 4 * c);
      if (row1) {
        const float4 w = ldg_stream(row1 + off + 4 * c);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      u[4 * c] = v.x; u[4 * c + 1] = v.y; u[4 * c + 2] = v.z; u[4 * c + 3] = v.w;
    }
    float m = u[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) m = fmaxf(m, u[j]);
    s.part[cg][row] = m;
    __syncthreads();  // also: every thread is past the previous tile's epilogue
    m = clamp_max(fmaxf(fmaxf(s.part[0][row], s.part[1][row]), fmaxf(s.part[2][row], s.part[3][row])));
    // ---- e = exp(u - m) -> (hi, lo) -> TMEM columns cg*32 .. cg*32+31 of lane `row`
    // (the previous tile's MMAs have completed: this thread waited for d_full in its epilogue)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e = valid ? exp_nonpos<FAST>(u[16 * h + j] - m) : 0.f;
        split_tf32(e, hi[j], lo[j]);
      }
      tmem_st16(lane_base + cg * 32 + 16 * h, hi);
      tmem_st16(lane_base + kColLo + cg * 32 + 16 * h, lo);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.a_full);
    // ---- MMA issue: whole warp converged, instructions on the elected lane
    if (warp == 0) {
      mbar_wait(&s.a_full, it & 1);
      tc_fence_after_sync();
      constexpr uint32_t idesc_n256 = make_idesc_tf32(TM, 2 * K128, 0, 0);
      constexpr uint32_t idesc_n128 = make_idesc_tf32(TM, K128, 0, 0);
      const uint64_t d_w = make_desc(smem_u32(s.w), 16, 1024);
#pragma unroll
      for (int ks = 0; ks < K128 / 8; ++ks) {
        const uint64_t b_w = desc_at(d_w, (ks >> 2) * kWBlk + (ks & 3) * 32);
        mma_tf32_ts_warp(tmem_base + kColMain, tmem_base + ks * 8, b_w, idesc_n256, ks ? 1u : 0u);
        mma_tf32_ts_warp(tmem_base + kColCorr, tmem_base + kColLo + ks * 8, b_w, idesc_n128, 1u);
      }
      mma_commit_warp(&s.d_full);
      __syncwarp();
    }
    // ---- epilogue: y = log(main + correction) + m for this thread's 32 outputs
    mbar_wait(&s.d_full, it & 1);
    tc_fence_after_sync();
    float* yo = a.y + ((int64_t)f * a.B + (valid ? b : 0)) * K128 + cg * 32;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[16], w[16];
      tmem_ld16(lane_base + kColMain + cg * 32 + 16 * h, v);
      tmem_ld16(lane_base + kColCorr + cg * 32 + 16 * h, w);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 o;
          o.x = log_<FAST>(v[j] + w[j]) + m;
          o.y = log_<FAST>(v[j + 1] + w[j + 1]) + m;
          o.z = log_<FAST>(v[j + 2] + w[j + 2]) + m;
          o.w = log_<FAST>(v[j + 3] + w[j + 3]) + m;
          *reinterpret_cast<float4*>(yo + 16 * h + j) = o;
        }
      }
    }
    tc_fence_before_sync();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}


// ==========================================================================================
// Backward.  r = g exp(m - y), e = exp(u - m):
//   part 1  du[b,i] = e[b,i] * sum_o r[b,o] W[o,i]      -- the forward's skeleton with A = r, the
//           TRANSPOSED weight image and a multiplying epilogue; also stores the row shifts m[f,b];
//   part 2  dW[o,i] = sum_b r[b,o] e[b,i]               -- contraction over the samples: a thread
//           owns (unit, 32 samples), so r^T goes to TMEM (lane = o) and e^T to shared memory
//           (row = i, K = samples) with plain stores, no transposes; the accumulator stays in
//           TMEM over all tiles of the CTA and batch splits are summed by reduce_partials.
// Two kernels because each needs all 512 TMEM columns (operand 256 + accumulators 256).
// ==========================================================================================
struct Cons128 {
  const float* g0;   // first consumer row block (or null: no consumer)
  int cons0, n_cons;
};
__device__ __forceinline__ Cons128 consumers_of(const DenseArgs& a, int f) {
  Cons128 c{nullptr, 0, 1};
  if (a.gs.cons_ptr == nullptr) {
    c.g0 = a.gs.garena + (int64_t)f * a.gs.B * K128;
  } else {
    c.cons0 = a.gs.cons_ptr[f];
    c.n_cons = a.gs.cons_ptr[f + 1] - c.cons0;
    if (c.n_cons > 0) c.g0 = a.gs.garena + a.gs.B * a.gs.cons_rows[c.cons0];
  }
  return c;
}

template <bool FAST>
__global__ void __launch_bounds__(kThreads128, 1)
dense128_tc_bwd_du_kernel(DenseArgs a, int tiles_per_cta, float* __restrict__ m_out) {
  extern __shared__ uint8_t smem_raw[];
  Fwd128Smem& s = *reinterpret_cast<Fwd128Smem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)((a.B + TM - 1) / TM);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(n_tiles_total, t_begin + tiles_per_cta) - t_begin;
  if (n_tiles <= 0) return;

  const float* row0 = in_row(a, f, 0);
  const float* row1 = a.H == 2 ? in_row(a, f, 1) : nullptr;
  const float* yrow = a.y + (int64_t)f * a.B * K128;
  const Cons128 cons = consumers_of(a, f);

  if (tid == 0) {
    mbar_init(&s.a_full, kThreads128 / 32);
    mbar_init(&s.d_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);

  // ---- W^T image: rows i (N), K = o.  Element (i, o) = W[o][i].  The lanes of a warp take 32
  // consecutive o's (the K axis, contiguous inside an image row), so the scalar stores of one
  // instruction fill one 128-byte row segment: no bank conflicts.
  {
    const float* Wf = a.W + (int64_t)f * K128 * K128;
    const uint32_t wbase = smem_u32(s.w);
    constexpr int PER = K128 * K128 / 4 / kThreads128;
    float4 v[PER];
#pragma unroll
    for (int n = 0; n < PER; ++n) {
      const int p = tid + n * kThreads128;
      v[n] = __ldg(reinterpret_cast<const float4*>(Wf + (p & 127) * K128 + (p >> 7) * 4));
    }
#pragma unroll
    for (int n = 0; n < PER; ++n) {
      const int p = tid + n * kThreads128;
      const uint32_t o = p & 127, i0 = (uint32_t)(p >> 7) * 4;
      float4 hi, lo;
      split4(v[n], hi, lo);
      const float h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
      const uint32_t kb = o >> 5, chunk = (o & 31) >> 2, within = (o & 3) * 4;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t i = i0 + t;
        const uint32_t off = kb * kWBlk + i * 128u + (((chunk ^ i) & 7u) << 4) + within;
        sts32(wbase + off, h[t]);
        sts32(wbase + off + 128u * 128u, l[t]);
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  const int q = warp & 3, cg = warp >> 2;
  const int row = q * 32 + lane;
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  constexpr uint32_t kColLo = 128, kColMain = 256, kColCorr = 384;

  for (int it = 0; it < n_tiles; ++it) {
    const int64_t b = (int64_t)(t_begin + it) * TM + row;
    const bool valid = b < a.B;
    const int64_t off = (valid ? b : 0) * K128 + cg * 32;
    float u[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 v = ldg_stream(row0 + off + 4 * c);
      if (row1) {
        const float4 w = ldg_stream(row1 + off + 4 * c);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      u[4 * c] = v.x; u[4 * c + 1] = v.y; u[4 * c + 2] = v.z; u[4 * c + 3] = v.w;
    }
    float m = u[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) m = fmaxf(m, u[j]);
    s.part[cg][row] = m;
    __syncthreads();
    m = clamp_max(fmaxf(fmaxf(s.part[0][row], s.part[1][row]), fmaxf(s.part[2][row], s.part[3][row])));
    if (cg == 0 && valid && m_out) m_out[(int64_t)f * a.B + b] = m;
    // ---- r = g exp(m - y) for this thread's 32 outputs -> (hi, lo) -> TMEM
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float hi[16], lo[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = 16 * h + 4 * c;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && cons.g0) {
          g = ldg_stream(cons.g0 + off + col);
          for (int k = 1; k < cons.n_cons; ++k) {
            const float4 z = ldg_stream(a.gs.garena + a.gs.B * a.gs.cons_rows[cons.cons0 + k] + off + col);
            g.x += z.x; g.y += z.y; g.z += z.z; g.w += z.w;
          }
        }
        const float4 yv = ldg_stream(yrow + off + col);
        const float r0 = g.x == 0.f ? 0.f : g.x * exp_capped<FAST>(m - yv.x);
        const float r1 = g.y == 0.f ? 0.f : g.y * exp_capped<FAST>(m - yv.y);
        const float r2 = g.z == 0.f ? 0.f : g.z * exp_capped<FAST>(m - yv.z);
        const float r3 = g.w == 0.f ? 0.f : g.w * exp_capped<FAST>(m - yv.w);
        split_tf32(r0, hi[4 * c], lo[4 * c]);
        split_tf32(r1, hi[4 * c + 1], lo[4 * c + 1]);
        split_tf32(r2, hi[4 * c + 2], lo[4 * c + 2]);
        split_tf32(r3, hi[4 * c + 3], lo[4 * c + 3]);
      }
      tmem_st16(lane_base + cg * 32 + 16 * h, hi);
      tmem_st16(lane_base + kColLo + cg * 32 + 16 * h, lo);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.a_full);
    if (warp == 0) {
      mbar_wait(&s.a_full, it & 1);
      tc_fence_after_sync();
      constexpr uint32_t idesc_n256 = make_idesc_tf32(TM, 2 * K128, 0, 0);
      constexpr uint32_t idesc_n128 = make_idesc_tf32(TM, K128, 0, 0);
      const uint64_t d_w = make_desc(smem_u32(s.w), 16, 1024);
#pragma unroll
      for (int ks = 0; ks < K128 / 8; ++ks) {  // 8 o's per step
        const uint64_t b_w = desc_at(d_w, (ks >> 2) * kWBlk + (ks & 3) * 32);
        mma_tf32_ts_warp(tmem_base + kColMain, tmem_base + ks * 8, b_w, idesc_n256, ks ? 1u : 0u);
        mma_tf32_ts_warp(tmem_base + kColCorr, tmem_base + kColLo + ks * 8, b_w, idesc_n128, 1u);
      }
      mma_commit_warp(&s.d_full);
      __syncwarp();
    }
    // ---- du[b,i] = e[b,i] * T[b,i] for this thread's 32 inputs
    mbar_wait(&s.d_full, it & 1);
    tc_fence_after_sync();
    float* du = a.gin + ((int64_t)f * a.B + (valid ? b : 0)) * K128 + cg * 32;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[16], w[16];
      tmem_ld16(lane_base + kColMain + cg * 32 + 16 * h, v);
      tmem_ld16(lane_base + kColCorr + cg * 32 + 16 * h, w);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 o;
          o.x = exp_nonpos<FAST>(u[16 * h + j] - m) * (v[j] + w[j]);
          o.y = exp_nonpos<FAST>(u[16 * h + j + 1] - m) * (v[j + 1] + w[j + 1]);
          o.z = exp_nonpos<FAST>(u[16 * h + j + 2] - m) * (v[j + 2] + w[j + 2]);
          o.w = exp_nonpos<FAST>(u[16 * h + j + 3] - m) * (v[j + 3] + w[j + 3]);
          *reinterpret_cast<float4*>(du + 16 * h + j) = o;
        }
      }
    }
    tc_fence_before_sync();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

struct __align__(1024) Dw128Smem {
  float eT[4][256 * 32];  // [sample block of 32][hi i 0..127 | lo i 0..127][32 samples]   128 KB
  uint64_t a_full, mma_done;
  uint32_t tmem_base;
};

template <bool FAST>
__global__ void __launch_bounds__(kThreads128, 1)
dense128_tc_bwd_dw_kernel(DenseArgs a, int tiles_per_cta, const float* __restrict__ m_in) {
  extern __shared__ uint8_t smem_raw[];
  Dw128Smem& s = *reinterpret_cast<Dw128Smem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)((a.B + TM - 1) / TM);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(n_tiles_total, t_begin + tiles_per_cta) - t_begin;
  if (n_tiles <= 0) return;

  const float* row0 = in_row(a, f, 0);
  const float* row1 = a.H == 2 ? in_row(a, f, 1) : nullptr;
  const float* yrow = a.y + (int64_t)f * a.B * K128;
  const float* mrow = m_in + (int64_t)f * a.B;
  const Cons128 cons = consumers_of(a, f);

  if (tid == 0) {
    mbar_init(&s.a_full, kThreads128 / 32);
    mbar_init(&s.mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  const int q = warp & 3, bg = warp >> 2;   // TMEM lane quadrant; block of 32 samples in the tile
  const int unit = q * 32 + lane;           // o for r^T (TMEM lane), i for e^T (shared-memory row)
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t eT = smem_u32(s.eT);
  constexpr uint32_t kColLo = 128, kColMain = 256, kColCorr = 384;

  for (int it = 0; it < n_tiles; ++it) {
    const int64_t bb = (int64_t)(t_begin + it) * TM + bg * 32;  // first sample of this thread's block
    // the previous tile's MMAs read r^T (TMEM) and e^T (shared memory): wait before overwriting
    if (it > 0) {
      mbar_wait(&s.mma_done, (it - 1) & 1);
      tc_fence_after_sync();
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float rhi[16], rlo[16], ehi[16], elo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int64_t b = bb + 16 * h + j;
        float r = 0.f, e = 0.f;
        if (b < a.B) {
          const int64_t at = b * K128 + unit;
          float u = __ldg(row0 + at);
          if (row1) u += __ldg(row1 + at);
          const float m = __ldg(mrow + b);
          e = exp_nonpos<FAST>(u - m);
          float g = 0.f;
          if (cons.g0) {
            g = __ldg(cons.g0 + at);
            for (int k = 1; k < cons.n_cons; ++k)
              g += __ldg(a.gs.garena + a.gs.B * a.gs.cons_rows[cons.cons0 + k] + at);
          }
          r = g == 0.f ? 0.f : g * exp_capped<FAST>(m - __ldg(yrow + at));
        }
        split_tf32(r, rhi[j], rlo[j]);
        split_tf32(e, ehi[j], elo[j]);
      }
      // r^T: TMEM lane `unit`, columns = the tile's samples bg*32 + 16h ..
      tmem_st16(lane_base + bg * 32 + 16 * h, rhi);
      tmem_st16(lane_base + kColLo + bg * 32 + 16 * h, rlo);
      // e^T: k-block bg, row `unit` (hi) / 128 + unit (lo), 16-byte chunks 4h .. 4h+3
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t chunk = 4 * h + c;
        const uint32_t off = (uint32_t)bg * kWBlk + (uint32_t)unit * 128u + (((chunk ^ (uint32_t)unit) & 7u) << 4);
        sts128(eT + off, make_float4(ehi[4 * c], ehi[4 * c + 1], ehi[4 * c + 2], ehi[4 * c + 3]));
        sts128(eT + off + 128u * 128u, make_float4(elo[4 * c], elo[4 * c + 1], elo[4 * c + 2], elo[4 * c + 3]));
      }
    }
    tmem_st_wait();
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.a_full);
    if (warp == 0) {
      mbar_wait(&s.a_full, it & 1);
      tc_fence_after_sync();
      constexpr uint32_t idesc_n256 = make_idesc_tf32(TM, 2 * K128, 0, 0);
      constexpr uint32_t idesc_n128 = make_idesc_tf32(TM, K128, 0, 0);
      const uint64_t d_e = make_desc(eT, 16, 1024);
#pragma unroll
      for (int ks = 0; ks < TM / 8; ++ks) {  // 8 samples per step
        const uint64_t b_e = desc_at(d_e, (ks >> 2) * kWBlk + (ks & 3) * 32);
        mma_tf32_ts_warp(tmem_base + kColMain, tmem_base + ks * 8, b_e, idesc_n256, (it || ks) ? 1u : 0u);
        mma_tf32_ts_warp(tmem_base + kColCorr, tmem_base + kColLo + ks * 8, b_e, idesc_n128, 1u);
      }
      mma_commit_warp(&s.mma_done);
      __syncwarp();
    }
  }
  // ---- dW[o][i] = main + correction: thread = (o = unit, 32 columns i of group bg)
  mbar_wait(&s.mma_done, (n_tiles - 1) & 1);
  tc_fence_after_sync();
  float* out = a.dWp + (((int64_t)blockIdx.x * gridDim.y + f) * K128 + unit) * K128 + bg * 32;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[16], w[16];
    tmem_ld16(lane_base + kColMain + bg * 32 + 16 * h, v);
    tmem_ld16(lane_base + kColCorr + bg * 32 + 16 * h, w);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      *reinterpret_cast<float4*>(out + 16 * h + j) =
          make_float4(v[j] + w[j], v[j + 1] + w[j + 1], v[j + 2] + w[j + 2], v[j + 3] + w[j + 3]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool dense128_tc_ok(const DenseArgs& a) {
  return (tc_flags() & 512) && !tc_disabled() && a.Ki == K128 && a.Ko == K128 && a.Kred == K128 &&
         !a.concat && a.H >= 1 && a.H <= 2;
}

int dense128_tc_fwd(const DenseArgs& a, int F, Ctx& c) {
  const size_t smem = sizeof(Fwd128Smem) + 1024;
  static PerDeviceOnce attr;
  if (attr.first()) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense128_tc_fwd_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense128_tc_fwd_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int n_tiles = ceil_div(a.B, TM);
  // one CTA per SM (all of TMEM); two waves of CTAs over the launch when the batch allows
  int splits = (int)max64(1, min64(n_tiles, ceil_div(2 * kNumSMs, F)));
  const int tiles_per_cta = ceil_div(n_tiles, splits);
  splits = ceil_div(n_tiles, tiles_per_cta);
  dim3 grid(splits, F);
  if ((tc_flags() & 3) == 3)
    dense128_tc_fwd_kernel<true><<<grid, kThreads128, smem, c.stream>>>(a, tiles_per_cta);
  else
    dense128_tc_fwd_kernel<false><<<grid, kThreads128, smem, c.stream>>>(a, tiles_per_cta);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// Workspace of the backward: the row shifts part 1 hands to part 2, then the dW partials.
static void dense128_bwd_config(int F, int64_t B, int& splits, int& tiles_per_cta) {
  const int n_tiles = ceil_div(B, TM);
  splits = (int)max64(1, min64(n_tiles, ceil_div(2 * kNumSMs, F)));
  tiles_per_cta = ceil_div(n_tiles, splits);
  splits = ceil_div(n_tiles, tiles_per_cta);
}

size_t dense128_tc_bwd_ws(int F, int64_t B) {
  int splits, tpc;
  dense128_bwd_config(F, B, splits, tpc);
  const size_t m_bytes = ((size_t)F * B * 4 + 255) & ~(size_t)255;
  return m_bytes + (size_t)splits * F * K128 * K128 * 4;
}

// Returns CKB_OK when it ran, 1 when the caller should use the SIMT kernels (workspace sized
// before the option was switched on), negative on errors.
int dense128_tc_bwd(const DenseArgs& a_in, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes) {
  if (ws_bytes < dense128_tc_bwd_ws(F, a_in.B)) return 1;
  DenseArgs a = a_in;
  int splits, tpc;
  dense128_bwd_config(F, a.B, splits, tpc);
  const size_t m_bytes = ((size_t)F * a.B * 4 + 255) & ~(size_t)255;
  float* m_buf = (float*)ws;
  const size_t n = (size_t)F * K128 * K128;
  a.dWp = splits > 1 ? (float*)(ws + m_bytes) : dW;
  const size_t smem_du = sizeof(Fwd128Smem) + 1024, smem_dw = sizeof(Dw128Smem) + 1024;
  static PerDeviceOnce attr;
  if (attr.first()) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense128_tc_bwd_du_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_du));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense128_tc_bwd_du_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_du));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense128_tc_bwd_dw_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dw));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense128_tc_bwd_dw_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dw));
  }
  dim3 grid(splits, F);
  const bool fast = (tc_flags() & 3) == 3;
  if (fast) dense128_tc_bwd_du_kernel<true><<<grid, kThreads128, smem_du, c.stream>>>(a, tpc, m_buf);
  else dense128_tc_bwd_du_kernel<false><<<grid, kThreads128, smem_du, c.stream>>>(a, tpc, m_buf);
  CKB_LAUNCH_CHECK();
  c.launches++;
  if (dW == nullptr) return CKB_OK;
  if (fast) dense128_tc_bwd_dw_kernel<true><<<grid, kThreads128, smem_dw, c.stream>>>(a, tpc, m_buf);
  else dense128_tc_bwd_dw_kernel<false><<<grid, kThreads128, smem_dw, c.stream>>>(a, tpc, m_buf);
  CKB_LAUNCH_CHECK();
  c.launches++;
  if (splits > 1) return reduce_partials(a.dWp, dW, (int64_t)n, splits, c);
  return CKB_OK;
}

}  // namespace ckb
