// Tensor-core (tcgen05 / TMEM) kernels of the Tucker layer for Ki = Ko = 64, sm_100a only.
// Reference: TorchTuckerLayer.forward, cirkit/backend/torch/layers/optimized.py:89-103 (einsum
// spec :62-66) through LSESumSemiring.apply_reduce, semiring.py:382-408.
//
//   forward   y[b,o] = log( sum_ij W[o,ij] e1[b,i] e2[b,j] ) + m1[b] + m2[b],  e = exp(x - m)
//   backward  r = g / S;  T[b,ij] = sum_o r[b,o] W[o,ij];
//             d/dx1[b,i] = e1[b,i] sum_j T[b,ij] e2[b,j];  d/dx2[b,j] = e2[b,j] sum_i T[b,ij] e1[b,i]
//             d/dW[o,ij] = sum_b r[b,o] e1[b,i] e2[b,j]
//
// The reference materialises nothing smaller than the einsum torch picks; its unfused route
// (Kronecker layer + sum layer) writes the (F, B, Ki^2) product to memory.  Here the Kronecker
// operand e1 (x) e2 is formed in registers and written straight into swizzled shared-memory MMA
// tiles, so only x1, x2, y, g and W touch HBM.  All products are 3xTF32 (hi*hi + hi*lo + lo*hi,
// see sm100.cuh), operand tiles are 128-byte-swizzled K-major.
//
// The tensor core truncates when it folds a product group into the fp32 accumulator (~0.6 ulp low
// per accumulating instruction, measured).  Over the 512 k-steps of a Ki^2 = 4096 reduction that
// would bias log S by ~1e-5 per layer, all with the same sign, so the forward accumulates in TMEM
// only over chunks of 8 k-steps and folds the chunks into fp32 registers with round-to-nearest
// adds (the same 8-step depth as the Ki = 64 sum-product block of dense_tc.cu).
#include <type_traits>

#include "dense.cuh"
#include "sm100.cuh"
#include "tc_util.cuh"

namespace ckb {
using namespace sm100;

namespace {

constexpr int TM = 128;        // UMMA M
constexpr int KK = 64;         // Ki = Ko
constexpr int KRED = KK * KK;  // 4096
constexpr int ROWS = 256;      // samples per CTA (two M tiles share every weight tile)
constexpr uint32_t kTile = TM * 128;  // bytes of a [128 rows][32 fp32] swizzled tile

__device__ __forceinline__ float4 ldg_nc(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float max4(const float4& v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }

// ==========================================================================================
// Forward.  CTA = (fold, 256 samples).  Warps 0-7: producers (one thread per sample row: the
// Kronecker operand tile and a share of the weight tile per k-block of 32 reduction indices),
// warps 8-15: chunk accumulation (TMEM -> registers) and the log epilogue.  Lane 0 of producer
// warp 0 issues the MMAs of a k-block right after its own share of the tile (a 17th warp would
// cap the kernel at 96 registers per thread: 5 warps on one scheduler's register file).
// ==========================================================================================
constexpr int kProdWarps = 8, kEpiWarps = 8;
constexpr int kFwdThreads = (kProdWarps + kEpiWarps) * 32;  // 512
constexpr int kChunkKb = 2;                          // k-blocks (of 4 k-steps) per TMEM chunk
constexpr int kNumKb = KRED / 32;                    // 128

struct __align__(1024) TkFwdSmem {
  float a_hi[2][2][TM * 32];  // [stage][M tile][row][32] swizzled     64 KB
  float a_lo[2][2][TM * 32];  //                                        64 KB
  float w[2][128 * 32];       // [stage][hi o 0..63 | lo o 0..63][32]   32 KB
  float msum[ROWS];
  uint64_t full[2], empty[2], tfull[2], tempty[2];
  uint32_t tmem_base;
};

template <bool FAST>
__global__ void __launch_bounds__(kFwdThreads, 1) tucker_tc_fwd_kernel(DenseArgs a) {
  extern __shared__ uint8_t smem_raw[];
  TkFwdSmem& s = *reinterpret_cast<TkFwdSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int64_t b0 = (int64_t)blockIdx.x * ROWS;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.full[i], kProdWarps);
      mbar_init(&s.empty[i], 1);
      mbar_init(&s.tfull[i], 1);
      mbar_init(&s.tempty[i], kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  if (warp < kProdWarps) {
    // ================= producers =================
    constexpr uint32_t idesc_n128 = make_idesc_tf32(TM, 2 * KK, 0, 0);
    constexpr uint32_t idesc_n64 = make_idesc_tf32(TM, KK, 0, 0);
    const uint64_t d_ahi = make_desc(smem_u32(s.a_hi), 16, 1024);
    const uint64_t d_alo = make_desc(smem_u32(s.a_lo), 16, 1024);
    const uint64_t d_w = make_desc(smem_u32(s.w), 16, 1024);
    const int p = tid;  // row inside the CTA
    const int64_t b = b0 + p;
    const bool valid = b < a.B;
    const float* x1 = in_row(a, f, 0) + (valid ? b : 0) * KK;
    const float* x2 = in_row(a, f, 1) + (valid ? b : 0) * KK;
    float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 16; ++c) m1 = fmaxf(m1, max4(ldg_nc(x1 + 4 * c)));
    m1 = clamp_max(m1);
    float e2[KK];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 v = ldg_nc(x2 + 4 * c);
      e2[4 * c] = v.x; e2[4 * c + 1] = v.y; e2[4 * c + 2] = v.z; e2[4 * c + 3] = v.w;
      m2 = fmaxf(m2, max4(v));
    }
    m2 = clamp_max(m2);
#pragma unroll
    for (int j = 0; j < KK; ++j) e2[j] = valid ? exp_nonpos<FAST>(e2[j] - m2) : 0.f;
    s.msum[p] = fmaxf(m1 + m2, -FLT_MAX);

    // this thread's two 16-byte pieces of every weight k-block: rows o_a / o_a + 32, chunk wc
    const int o_a = p >> 3, wc = p & 7;
    const float* w0 = a.W + ((int64_t)f * KK + o_a) * KRED + wc * 4;
    const float* w1 = w0 + (int64_t)32 * KRED;
    const uint32_t w_off0 = (uint32_t)o_a * 128u + ((((uint32_t)wc ^ (uint32_t)o_a) & 7u) << 4);
    const uint32_t w_off1 = w_off0 + 32u * 128u;  // (o_a + 32) & 7 == o_a & 7
    const uint32_t r = (uint32_t)p & 127u, t = (uint32_t)p >> 7;
    const uint32_t a_row = t * kTile + r * 128u;
    const uint32_t ahi = smem_u32(s.a_hi), alo = smem_u32(s.a_lo), wsm = smem_u32(s.w);

    float4 wn0 = ldg_stream(w0), wn1 = ldg_stream(w1);
    float x1n = __ldg(x1);
    auto body = [&](const int kb, auto JH, const float a1) {
      constexpr int jh = decltype(JH)::value;
      const float4 wc0 = wn0, wc1 = wn1;
      if (kb + 1 < kNumKb) {
        wn0 = ldg_stream(w0 + (kb + 1) * 32);
        wn1 = ldg_stream(w1 + (kb + 1) * 32);
      }
      const uint32_t stage = (uint32_t)kb & 1u;
      mbar_wait(&s.empty[stage], ((kb >> 1) & 1) ^ 1);
      const uint32_t abase = stage * (2 * kTile) + a_row;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 pr, hi, lo;
        pr.x = a1 * e2[jh * 32 + 4 * c];
        pr.y = a1 * e2[jh * 32 + 4 * c + 1];
        pr.z = a1 * e2[jh * 32 + 4 * c + 2];
        pr.w = a1 * e2[jh * 32 + 4 * c + 3];
        split4(pr, hi, lo);
        const uint32_t off = abase + ((((uint32_t)c ^ r) & 7u) << 4);
        sts128(ahi + off, hi);
        sts128(alo + off, lo);
      }
      {
        float4 hi, lo;
        const uint32_t wb = wsm + stage * kTile;
        split4(wc0, hi, lo);
        sts128(wb + w_off0, hi);
        sts128(wb + w_off0 + 64 * 128, lo);
        split4(wc1, hi, lo);
        sts128(wb + w_off1, hi);
        sts128(wb + w_off1 + 64 * 128, lo);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.full[stage]);
      if (warp == 0) {
        if (lane == 0) {
          // ---- MMA issue for this k-block
          const int chunk = kb / kChunkKb, buf = chunk & 1;
          const bool first = (kb % kChunkKb) == 0;
          if (first) mbar_wait(&s.tempty[buf], ((chunk >> 1) & 1) ^ 1);
          mbar_wait(&s.full[stage], (kb >> 1) & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) {
            const uint32_t d = tmem_base + tt * 256 + buf * 128;
            const uint32_t aoff = stage * (2 * kTile) + tt * kTile;
            //   e_hi x [W_hi | W_lo]  (N = 128): main | correction;  e_lo x W_hi (N = 64): correction
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_tf32(d, desc_at(d_ahi, aoff + ks * 32), desc_at(d_w, stage * kTile + ks * 32),
                       idesc_n128, (first && ks == 0) ? 0u : 1u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_tf32(d + KK, desc_at(d_alo, aoff + ks * 32), desc_at(d_w, stage * kTile + ks * 32),
                       idesc_n64, 1u);
          }
          mma_commit(&s.empty[stage]);
          if ((kb % kChunkKb) == kChunkKb - 1) mma_commit(&s.tfull[buf]);
        }
        __syncwarp();
      }
    };
    for (int i = 0; i < KK; ++i) {
      const float a1 = exp_nonpos<FAST>(x1n - m1);
      if (i + 1 < KK) x1n = __ldg(x1 + i + 1);
      body(2 * i, std::integral_constant<int, 0>{}, a1);
      body(2 * i + 1, std::integral_constant<int, 1>{}, a1);
    }
  } else {
    // ================= chunk accumulation + epilogue =================
    const int q = warp & 3, t = (warp - kProdWarps) >> 2;
    const int row = t * TM + q * 32 + lane;
    float acc[KK];
#pragma unroll
    for (int j = 0; j < KK; ++j) acc[j] = 0.f;
    constexpr int kChunks = kNumKb / kChunkKb;
    for (int c = 0; c < kChunks; ++c) {
      const int buf = c & 1;
      mbar_wait_relaxed(&s.tfull[buf], (c >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + t * 256 + buf * 128;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float v[16], w[16];
        tmem_ld16(taddr + cc * 16, v);
        tmem_ld16(taddr + KK + cc * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[cc * 16 + j] += v[j] + w[j];
      }
      tc_fence_before_sync();
      mbar_arrive(&s.tempty[buf]);
    }
    const int64_t b = b0 + row;
    if (b < a.B) {
      const float ms = s.msum[row];
      float* yo = a.y + ((int64_t)f * a.B + b) * KK;
#pragma unroll
      for (int j = 0; j < KK; j += 4) {
        float4 o;
        o.x = log_<FAST>(acc[j]) + ms;
        o.y = log_<FAST>(acc[j + 1]) + ms;
        o.z = log_<FAST>(acc[j + 2]) + ms;
        o.w = log_<FAST>(acc[j + 3]) + ms;
        *reinterpret_cast<float4*>(yo + j) = o;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================================
// Backward, part 1 (d/dx1, d/dx2 and the operands of part 2).  CTA = (fold, 256 samples).
// 16 worker warps; thread = (sample row, column half).  Prologue: row maxes, e2, r = g/S as
// (hi, lo) A tiles [sample][o]; the transposed, split r and the raw e1, e2 go to a scratch block
// per 32 samples, already in the swizzled image part 2 multiplies from.  Main loop over i:
// the 64x64 slice W[:, i, :] is staged transposed ([j][o], so that o is the K axis), GEMM
// T_i[b, j] = sum_o r[b,o] W[o,i,j] (K = 64: 8 k-steps, no long accumulation), and the workers
// fold T_i into d/dx1[b,i] (dot with e2) and d/dx2[b,:] (axpy with e1[b,i]).
// ==========================================================================================
constexpr int kDxWorkers = 16;
constexpr int kDxThreads = kDxWorkers * 32;  // 512; lane 0 of worker 0 issues the MMAs

// scratch block of 32 samples (floats): r stacked [hi o 0..63 | lo o 0..63][32 b] | e1 [i][32 b] |
// e2 [j][32 b]; every [unit][32] row is 128 bytes with its 16-byte chunks xor-swizzled by unit & 7
constexpr int kBlkFloats = 128 * 32 + 64 * 32 + 64 * 32;  // 8192 floats = 32 KB
constexpr int kBlkE1 = 128 * 32, kBlkE2 = 128 * 32 + 64 * 32;
__device__ __forceinline__ int blk_off(int unit, int bcol) {
  return unit * 32 + ((((bcol >> 2) ^ unit) & 7) << 2) + (bcol & 3);
}

struct __align__(1024) TkDxSmem {
  float r_hi[2][2][TM * 32];  // [M tile][o half][row][32]                        64 KB
  float r_lo[2][2][TM * 32];  //                                                   64 KB
  float w[2][2][128 * 32];    // [stage][o half][hi j 0..63 | lo j 0..63][32 o]    64 KB
  float part[2][ROWS];
  uint64_t wfull[2], wempty[2], tfull[2], tempty[2];
  uint32_t tmem_base;
};

template <bool FAST>
__global__ void __launch_bounds__(kDxThreads, 1)
tucker_tc_bwd_dx_kernel(DenseArgs a, float* scratch, int nblk) {
  extern __shared__ uint8_t smem_raw[];
  TkDxSmem& s = *reinterpret_cast<TkDxSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int64_t b0 = (int64_t)blockIdx.x * ROWS;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.wfull[i], kDxWorkers);
      mbar_init(&s.wempty[i], 1);
      mbar_init(&s.tfull[i], 1);
      mbar_init(&s.tempty[i], kDxWorkers * 32);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  {
    const int q = warp & 3, t = (warp >> 2) & 1, ch = warp >> 3;
    const int p = t * TM + q * 32 + lane;  // row inside the CTA
    const int64_t b = b0 + p;
    const bool valid = b < a.B;
    const int64_t bsafe = valid ? b : 0;
    const float* x1 = in_row(a, f, 0) + bsafe * KK;
    const float* x2 = in_row(a, f, 1) + bsafe * KK;
    float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      m1 = fmaxf(m1, max4(ldg_nc(x1 + 4 * c)));
      m2 = fmaxf(m2, max4(ldg_nc(x2 + 4 * c)));
    }
    m1 = clamp_max(m1);
    m2 = clamp_max(m2);
    const float ms = fmaxf(m1 + m2, -FLT_MAX);
    float* blk = scratch ? scratch + ((int64_t)f * nblk + (b0 + p) / 32) * kBlkFloats : nullptr;

    // e2 (this thread's 32 columns)
    float e2[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 v = ldg_nc(x2 + 32 * ch + 4 * c);
      e2[4 * c] = valid ? exp_nonpos<FAST>(v.x - m2) : 0.f;
      e2[4 * c + 1] = valid ? exp_nonpos<FAST>(v.y - m2) : 0.f;
      e2[4 * c + 2] = valid ? exp_nonpos<FAST>(v.z - m2) : 0.f;
      e2[4 * c + 3] = valid ? exp_nonpos<FAST>(v.w - m2) : 0.f;
    }
    if (blk) {
#pragma unroll
      for (int c = 0; c < 32; ++c) blk[kBlkE2 + blk_off(32 * ch + c, lane)] = e2[c];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = ldg_nc(x1 + 32 * ch + 4 * c);
        const int u = 32 * ch + 4 * c;
        blk[kBlkE1 + blk_off(u, lane)] = valid ? exp_nonpos<FAST>(v.x - m1) : 0.f;
        blk[kBlkE1 + blk_off(u + 1, lane)] = valid ? exp_nonpos<FAST>(v.y - m1) : 0.f;
        blk[kBlkE1 + blk_off(u + 2, lane)] = valid ? exp_nonpos<FAST>(v.z - m1) : 0.f;
        blk[kBlkE1 + blk_off(u + 3, lane)] = valid ? exp_nonpos<FAST>(v.w - m1) : 0.f;
      }
    }
    // r = g * exp(m1 + m2 - y) for this thread's 32 outputs
    {
      const float* yrow = a.y + ((int64_t)f * a.B + bsafe) * KK + 32 * ch;
      int cbeg = 0, cend = 0;
      if (valid) {
        cbeg = a.gs.cons_ptr[f];
        cend = a.gs.cons_ptr[f + 1];
      }
      const uint32_t rhi = smem_u32(s.r_hi) + (uint32_t)(t * 2 + ch) * kTile;
      const uint32_t rlo = smem_u32(s.r_lo) + (uint32_t)(t * 2 + ch) * kTile;
      const uint32_t rr = (uint32_t)(q * 32 + lane);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = cbeg; k < cend; ++k) {
          const float4 z = *reinterpret_cast<const float4*>(a.gs.garena + a.gs.B * a.gs.cons_rows[k] +
                                                            b * KK + 32 * ch + 4 * c);
          g.x += z.x; g.y += z.y; g.z += z.z; g.w += z.w;
        }
        const float4 yv = ldg_nc(yrow + 4 * c);
        float4 rv, hi, lo;
        rv.x = g.x * exp_capped<FAST>(ms - yv.x);
        rv.y = g.y * exp_capped<FAST>(ms - yv.y);
        rv.z = g.z * exp_capped<FAST>(ms - yv.z);
        rv.w = g.w * exp_capped<FAST>(ms - yv.w);
        split4(rv, hi, lo);
        const uint32_t off = rr * 128u + ((((uint32_t)c ^ rr) & 7u) << 4);
        sts128(rhi + off, hi);
        sts128(rlo + off, lo);
        if (blk) {
          const int o = 32 * ch + 4 * c;
          blk[blk_off(o, lane)] = hi.x;
          blk[blk_off(o + 1, lane)] = hi.y;
          blk[blk_off(o + 2, lane)] = hi.z;
          blk[blk_off(o + 3, lane)] = hi.w;
          blk[blk_off(64 + o, lane)] = lo.x;
          blk[blk_off(64 + o + 1, lane)] = lo.y;
          blk[blk_off(64 + o + 2, lane)] = lo.z;
          blk[blk_off(64 + o + 3, lane)] = lo.w;
        }
      }
    }

    // weight staging: unit u = 2*warp + n: rows o0..o0+3 of W[:, i, j0..j0+31], transposed on the fly
    const float* wsrc[2];
    uint32_t wdst[2];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const int u = 2 * warp + n;
      const int o0 = 4 * (u & 15), j0 = 32 * (u >> 4);
      wsrc[n] = a.W + ((int64_t)f * KK + o0 + (lane & 3)) * KRED + j0 + 4 * (lane >> 2);
      const uint32_t j = (uint32_t)(j0 + lane);
      wdst[n] = (uint32_t)(o0 >> 5) * kTile + j * 128u + (((((uint32_t)o0 & 31u) >> 2) ^ j) & 7u) * 16u;
    }
    const uint32_t wsm = smem_u32(s.w);
    float4 wn[2];
    wn[0] = ldg_stream(wsrc[0]);
    wn[1] = ldg_stream(wsrc[1]);
    auto stage_w = [&](int i) {  // writes W[:, i, :] (held in wn) and fetches slice i + 1
      const uint32_t st = (uint32_t)i & 1u;
      mbar_wait(&s.wempty[st], ((i >> 1) & 1) ^ 1);
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        float4 hi, lo;
        split4(transpose4(wn[n], lane & 3), hi, lo);
        const uint32_t dst = wsm + st * (2 * kTile) + wdst[n];
        sts128(dst, hi);
        sts128(dst + 64 * 128, lo);
      }
      if (i + 1 < KK) {
        wn[0] = ldg_stream(wsrc[0] + (i + 1) * KK);
        wn[1] = ldg_stream(wsrc[1] + (i + 1) * KK);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.wfull[st]);
    };
    constexpr uint32_t idesc_n128 = make_idesc_tf32(TM, 2 * KK, 0, 0);
    constexpr uint32_t idesc_n64 = make_idesc_tf32(TM, KK, 0, 0);
    const uint64_t d_rhi = make_desc(smem_u32(s.r_hi), 16, 1024);
    const uint64_t d_rlo = make_desc(smem_u32(s.r_lo), 16, 1024);
    const uint64_t d_w = make_desc(smem_u32(s.w), 16, 1024);
    auto issue_mma = [&](int i) {  // worker 0: T_i = r W[:, i, :] for both M tiles
      if (warp == 0) {
        if (lane == 0) {
          const uint32_t st = i & 1, buf = i & 1;
          mbar_wait(&s.tempty[buf], ((i >> 1) & 1) ^ 1);
          mbar_wait(&s.wfull[st], (i >> 1) & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) {
            const uint32_t d = tmem_base + tt * 256 + buf * 128;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_tf32(d, desc_at(d_rhi, (tt * 2 + (ks >> 2)) * kTile + (ks & 3) * 32),
                       desc_at(d_w, (st * 2 + (ks >> 2)) * kTile + (ks & 3) * 32), idesc_n128,
                       ks ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_tf32(d + KK, desc_at(d_rlo, (tt * 2 + (ks >> 2)) * kTile + (ks & 3) * 32),
                       desc_at(d_w, (st * 2 + (ks >> 2)) * kTile + (ks & 3) * 32), idesc_n64, 1u);
          }
          mma_commit(&s.wempty[st]);
          mma_commit(&s.tfull[buf]);
        }
        __syncwarp();
      }
    };
    stage_w(0);
    issue_mma(0);

    float acc2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc2[j] = 0.f;
    float* gin1 = a.gin + (((int64_t)f * 2 + 0) * a.B + bsafe) * KK;
    float x1n = __ldg(x1);
    for (int i = 0; i < KK; ++i) {
      if (i + 1 < KK) {
        stage_w(i + 1);
        issue_mma(i + 1);
      }
      const float e1i = valid ? exp_nonpos<FAST>(x1n - m1) : 0.f;
      if (i + 1 < KK) x1n = __ldg(x1 + i + 1);
      const int buf = i & 1;
      mbar_wait(&s.tfull[buf], (i >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + t * 256 + buf * 128 + ch * 32;
      float dot = 0.f;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float v[8], w[8];
        tmem_ld8(taddr + h * 8, v);
        tmem_ld8(taddr + KK + h * 8, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float T = v[j] + w[j];
          dot = fmaf(T, e2[h * 8 + j], dot);
          acc2[h * 8 + j] = fmaf(T, e1i, acc2[h * 8 + j]);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&s.tempty[buf]);
      if (ch == 1) s.part[buf][p] = dot;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 7)) : "memory");
      if (ch == 0 && valid) gin1[i] = e1i * (dot + s.part[buf][p]);
    }
    if (valid) {
      float* gin2 = a.gin + (((int64_t)f * 2 + 1) * a.B + b) * KK + 32 * ch;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(gin2 + 4 * c) =
            make_float4(e2[4 * c] * acc2[4 * c], e2[4 * c + 1] * acc2[4 * c + 1],
                        e2[4 * c + 2] * acc2[4 * c + 2], e2[4 * c + 3] * acc2[4 * c + 3]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================================
// Backward, part 2 (d/dW).  CTA = (fold, 256 consecutive reduction indices n = (i, j): four i's),
// loops over ALL samples in blocks of 32 (the K axis of this GEMM):
//   dW[o, n] = sum_b r[b,o] P[b,n],  P[b,(i,j)] = e1[b,i] e2[b,j]
// A = stacked [r_hi^T ; r_lo^T] (128 x 32 per block, bulk-copied as written by part 1),
// B = P_hi^T / P_lo^T (256 x 32), formed by 256 producer threads (one per n) from the raw e1 / e2
// rows of the block.  Two M128 x N256 accumulators (A x P_hi, A x P_lo) fill the 512 TMEM columns;
// dW = D0[o] + D0[64 + o] + D1[o]  (the r_lo x P_lo quadrant is dropped).
// ==========================================================================================
constexpr int kDwProdWarps = 8;
constexpr int kDwTmaWarp = 8, kDwMmaWarp = 9;
constexpr int kDwThreads = 10 * 32;
constexpr int NCH = 256;  // reduction indices per CTA

struct __align__(1024) TkDwSmem {
  float rstack[2][128 * 32];  // 32 KB
  float p_hi[2][NCH * 32];    // 64 KB
  float p_lo[2][NCH * 32];    // 64 KB
  float e2raw[2][64 * 32];    // 16 KB
  float e1raw[2][4 * 32];     //  1 KB
  uint64_t raw_full[2], p_full[2], empty[2], done;
  uint32_t tmem_base;
};
static_assert(NCH * (KK + 1) * 4 <= 2 * sizeof(float) * 2 * NCH * 32, "exchange buffer fits p_hi + p_lo");

__global__ void __launch_bounds__(kDwThreads, 1)
tucker_tc_bwd_dw_kernel(const float* __restrict__ scratch, int nblk_alloc, int nblk, float* dW) {
  extern __shared__ uint8_t smem_raw[];
  TkDwSmem& s = *reinterpret_cast<TkDwSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int i0 = 4 * blockIdx.x, n0 = NCH * blockIdx.x;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.raw_full[i], 1);
      mbar_init(&s.p_full[i], kDwProdWarps);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.done, 1);
    fence_barrier_init();
  }
  if (warp == kDwMmaWarp) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;
  const float* fblk = scratch + (int64_t)f * nblk_alloc * kBlkFloats;

  if (warp == kDwTmaWarp) {
    if (lane == 0) {
      for (int kb = 0; kb < nblk; ++kb) {
        const int st = kb & 1;
        mbar_wait(&s.empty[st], ((kb >> 1) & 1) ^ 1);
        const float* blk = fblk + (int64_t)kb * kBlkFloats;
        mbar_arrive_expect_tx(&s.raw_full[st], 128 * 128 + 64 * 128 + 4 * 128);
        bulk_g2s(s.rstack[st], blk, 128 * 128, &s.raw_full[st]);
        bulk_g2s(s.e2raw[st], blk + kBlkE2, 64 * 128, &s.raw_full[st]);
        bulk_g2s(s.e1raw[st], blk + kBlkE1 + i0 * 32, 4 * 128, &s.raw_full[st]);
      }
    }
  } else if (warp == kDwMmaWarp) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, NCH, 0, 0);
      const uint64_t d_r = make_desc(smem_u32(s.rstack), 16, 1024);
      const uint64_t d_ph = make_desc(smem_u32(s.p_hi), 16, 1024);
      const uint64_t d_pl = make_desc(smem_u32(s.p_lo), 16, 1024);
      for (int kb = 0; kb < nblk; ++kb) {
        const uint32_t st = kb & 1;
        mbar_wait(&s.raw_full[st], (kb >> 1) & 1);
        mbar_wait(&s.p_full[st], (kb >> 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t da = desc_at(d_r, st * (128 * 128) + ks * 32);
          mma_tf32(tmem_base, da, desc_at(d_ph, st * (NCH * 128) + ks * 32), idesc, (kb | ks) ? 1u : 0u);
          mma_tf32(tmem_base + NCH, da, desc_at(d_pl, st * (NCH * 128) + ks * 32), idesc,
                   (kb | ks) ? 1u : 0u);
        }
        mma_commit(&s.empty[st]);
      }
      mma_commit(&s.done);
    }
  } else {
    // ================= producers: P^T tiles =================
    const int n = tid;                 // row of the P^T tile: (i0 + il, j)
    const int il = n >> 6, j = n & 63;
    const uint32_t e1row = (uint32_t)il * 128u, e1key = (uint32_t)(i0 + il) & 7u;
    const uint32_t e2row = (uint32_t)j * 128u, e2key = (uint32_t)j & 7u;
    const uint32_t prow = (uint32_t)n * 128u, pkey = (uint32_t)n & 7u;
    const uint32_t e1s = smem_u32(s.e1raw), e2s = smem_u32(s.e2raw);
    const uint32_t phs = smem_u32(s.p_hi), pls = smem_u32(s.p_lo);
    for (int kb = 0; kb < nblk; ++kb) {
      const uint32_t st = kb & 1;
      mbar_wait(&s.raw_full[st], (kb >> 1) & 1);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 u = lds128(e1s + st * (4 * 128) + e1row + ((((uint32_t)c ^ e1key) & 7u) << 4));
        const float4 v = lds128(e2s + st * (64 * 128) + e2row + ((((uint32_t)c ^ e2key) & 7u) << 4));
        float4 pr, hi, lo;
        pr.x = u.x * v.x; pr.y = u.y * v.y; pr.z = u.z * v.z; pr.w = u.w * v.w;
        split4(pr, hi, lo);
        const uint32_t off = st * (NCH * 128) + prow + ((((uint32_t)c ^ pkey) & 7u) << 4);
        sts128(phs + off, hi);
        sts128(pls + off, lo);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.p_full[st]);
    }
    // ================= epilogue =================
    mbar_wait_relaxed(&s.done, 0);
    tc_fence_after_sync();
    const int q = warp & 3, half = warp >> 2;  // TMEM lane quadrant; 128-column half
    const uint32_t xch = smem_u32(s.p_hi);     // [256 columns][64 + 1] exchange buffer over p_hi | p_lo (dead)
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + half * 128;
    if (q >= 2) {
      const int o = (q - 2) * 32 + lane;
#pragma unroll 2
      for (int cc = 0; cc < 8; ++cc) {
        float v[16];
        tmem_ld16(taddr + cc * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj)
          sts32(xch + (uint32_t)((half * 128 + cc * 16 + jj) * (KK + 1) + o) * 4u, v[jj]);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kDwProdWarps * 32) : "memory");
    if (q < 2) {
      const int o = q * 32 + lane;
      float* out = dW + ((int64_t)f * KK + o) * KRED + n0 + half * 128;
#pragma unroll 2
      for (int cc = 0; cc < 8; ++cc) {
        float v[16], w[16];
        tmem_ld16(taddr + cc * 16, v);
        tmem_ld16(taddr + NCH + cc * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj)
          v[jj] += w[jj] + lds32(xch + (uint32_t)((half * 128 + cc * 16 + jj) * (KK + 1) + o) * 4u);
#pragma unroll
        for (int jj = 0; jj < 16; jj += 4)
          *reinterpret_cast<float4*>(out + cc * 16 + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kDwMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  CKB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return CKB_OK;
}

}  // namespace

bool tucker_tc_ok(const ckb_step_desc_t& d) {
  return !tc_disabled() && d.arity == 2 && d.k_in == KK && d.k_out == KK;
}

size_t tucker_tc_ws(const ckb_step_desc_t& d, int64_t B) {
  const int64_t nblk = (B + ROWS - 1) / ROWS * (ROWS / 32);
  return (size_t)d.num_folds * nblk * kBlkFloats * 4;
}

static DenseArgs tucker_tc_args(const ckb_step_desc_t& d, Ctx& c) {
  DenseArgs a{};
  a.W = c.tensors[d.slot[0]];
  a.in_rows = d.in_rows;
  a.arena = c.arena;
  a.y = c.arena + c.B * d.out_off;
  a.B = c.B;
  a.H = 2;
  a.Ki = KK;
  a.Ko = KK;
  a.Kred = KRED;
  a.concat = 0;
  return a;
}

int tucker_tc_fwd(const ckb_step_desc_t& d, Ctx& c) {
  static bool attr = false;
  const size_t smem = sizeof(TkFwdSmem) + 1024;
  if (!attr) {
    if (int rc = set_smem(tucker_tc_fwd_kernel<true>, smem)) return rc;
    if (int rc = set_smem(tucker_tc_fwd_kernel<false>, smem)) return rc;
    attr = true;
  }
  const DenseArgs a = tucker_tc_args(d, c);
  dim3 grid(ceil_div(c.B, ROWS), d.num_folds);
  if ((tc_flags() & 3) == 3)
    tucker_tc_fwd_kernel<true><<<grid, kFwdThreads, smem, c.stream>>>(a);
  else
    tucker_tc_fwd_kernel<false><<<grid, kFwdThreads, smem, c.stream>>>(a);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

int tucker_tc_bwd(const ckb_step_desc_t& d, Ctx& c) {
  static bool attr = false;
  const size_t smem_dx = sizeof(TkDxSmem) + 1024, smem_dw = sizeof(TkDwSmem) + 1024;
  if (!attr) {
    if (int rc = set_smem(tucker_tc_bwd_dx_kernel<true>, smem_dx)) return rc;
    if (int rc = set_smem(tucker_tc_bwd_dx_kernel<false>, smem_dx)) return rc;
    if (int rc = set_smem(tucker_tc_bwd_dw_kernel, smem_dw)) return rc;
    attr = true;
  }
  DenseArgs a = tucker_tc_args(d, c);
  a.gs = GradSrc{c.garena, d.cons_ptr, d.cons_rows, c.B};
  a.gin = c.garena + c.B * d.gin_off;
  float* dW = c.grads[d.slot[0]];
  float* scratch = nullptr;
  const int nblk_alloc = ceil_div(c.B, ROWS) * (ROWS / 32);
  if (dW) {
    if (c.ws_bytes < tucker_tc_ws(d, c.B)) {
      set_error("tucker_bwd: workspace too small (%zu < %zu)", c.ws_bytes, tucker_tc_ws(d, c.B));
      return CKB_ERR_WORKSPACE;
    }
    scratch = (float*)c.ws;
  }
  dim3 grid(ceil_div(c.B, ROWS), d.num_folds);
  if ((tc_flags() & 3) == 3)
    tucker_tc_bwd_dx_kernel<true><<<grid, kDxThreads, smem_dx, c.stream>>>(a, scratch, nblk_alloc);
  else
    tucker_tc_bwd_dx_kernel<false><<<grid, kDxThreads, smem_dx, c.stream>>>(a, scratch, nblk_alloc);
  CKB_LAUNCH_CHECK();
  c.launches++;
  if (dW) {
    dim3 grid2(KRED / NCH, d.num_folds);
    tucker_tc_bwd_dw_kernel<<<grid2, kDwThreads, smem_dw, c.stream>>>(scratch, nblk_alloc,
                                                                    ceil_div(c.B, 32), dW);
    CKB_LAUNCH_CHECK();
    c.launches++;
  }
  return CKB_OK;
}

}  // namespace ckb
