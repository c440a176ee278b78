// Tensor-core (tcgen05 / TMEM) version of the fused sum-product block for the hot shape
// Ki = Ko = 64 (Hadamard arity <= 2), sm_100a only.
//
//   y[b,o] = log( sum_i W[o,i] * exp(u[b,i] - m[b]) ) + m[b],   u = sum_h x_h,  m = max_i u
//
// One CTA owns a fold and walks over 128-sample tiles with a warp-specialised pipeline
// (two CTAs are resident per SM, so the phases of one overlap the other's):
//
//   transform warps : coalesced 16-byte loads of the H input rows (a half-warp per 256-byte row),
//                     u -> max (shuffles) -> e = exp(u - m) -> split e into two tf32 terms
//                     (hi, lo) written as 128B-swizzled UMMA operand tiles; the next tile's rows
//                     are prefetched into L2 meanwhile                              [a_full/empty]
//   MMA thread      : D(128x64, TMEM) = e_hi W_hi^T + e_lo W_hi^T + e_hi W_lo^T
//                     (kind::tf32, three products = fp32-grade accuracy)        [tmem_full/empty]
//   epilogue warps  : tcgen05.ld D -> log -> + m -> staged through shared memory so that every
//                     store instruction writes whole 32-byte sectors of y (two TMEM buffers: the
//                     epilogue of tile t overlaps the transform + MMA of tile t+1)
//
// The 64x64 weight slice of the fold is split (hi, lo) and swizzled into shared memory once per
// CTA.  No intermediate of the block touches HBM: inputs are read once, y is written once.
#include "dense.cuh"
#include "sm100.cuh"

namespace ckb {
using namespace sm100;

namespace {

constexpr int TM = 128;  // samples per tile (UMMA M)
constexpr int KK = 64;   // Ki = Ko
constexpr int kTransformWarps = 8, kEpilogueWarps = 4;
constexpr int kMmaWarp = kTransformWarps + kEpilogueWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;  // 416

__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct __align__(1024) FwdSmem {
  float a_hi[2][TM * 32];  // [k-block][row][32] swizzled               32 KB
  float a_lo[2][TM * 32];  //                                           32 KB
  float w_hi[2][KK * 32];  // [k-block][o][32] swizzled                 16 KB
  float w_lo[2][KK * 32];  //                                           16 KB
  float stage[kEpilogueWarps][32 * 16];  // per-warp 32 rows x 16 columns   8 KB
  float m_buf[2][TM];
  uint64_t a_full, a_empty, tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 2) dense_tc_fwd_kernel(DenseArgs a, int tiles_per_cta, int fast_math) {
  extern __shared__ uint8_t smem_raw[];
  FwdSmem& s = *reinterpret_cast<FwdSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)((a.B + TM - 1) / TM);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(n_tiles_total, t_begin + tiles_per_cta) - t_begin;
  if (n_tiles <= 0) return;

  // ---- one-time setup: barriers, TMEM, weights
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.tmem_full[i], 1);
      mbar_init(&s.tmem_empty[i], kEpilogueWarps * 32);
    }
    mbar_init(&s.a_full, kTransformWarps);
    mbar_init(&s.a_empty, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&s.tmem_base, 256);
  {
    const float* Wf = a.W + (int64_t)f * KK * KK;  // [o][i]
    for (int idx = tid; idx < KK * KK; idx += kThreads) {
      const int o = idx >> 6, i = idx & 63;
      float hi, lo;
      split_tf32(Wf[idx], hi, lo);
      const uint32_t off = swz_off(o, i & 31) >> 2;
      s.w_hi[i >> 5][off] = hi;
      s.w_lo[i >> 5][off] = lo;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TM, KK, 0, 0);
      const uint32_t a_addr[2] = {smem_u32(s.a_hi), smem_u32(s.a_lo)};
      const uint32_t w_addr[2] = {smem_u32(s.w_hi), smem_u32(s.w_lo)};
      for (int it = 0; it < n_tiles; ++it) {
        const int buf = it & 1;
        mbar_wait(&s.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_wait(&s.a_full, it & 1);
        tc_fence_after_sync();
        // The tensor core truncates when it folds a product group into the fp32 accumulator
        // (measured: ~0.6 ulp low per accumulating instruction), so the two small correction
        // products get their own accumulator: only the 8 hi*hi steps touch the large one, and
        // the epilogue adds the two in round-to-nearest fp32.
#pragma unroll
        for (int p = 0; p < 3; ++p) {  // hi*hi | lo*hi, hi*lo
          const uint32_t ab = a_addr[p == 1 ? 1 : 0], wb = w_addr[p == 2 ? 1 : 0];
          const uint32_t d = tmem_base + buf * 128 + (p == 0 ? 0 : KK);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = make_desc(ab + kb * (TM * 128) + ks * 32, 16, 1024);
              const uint64_t db = make_desc(wb + kb * (KK * 128) + ks * 32, 16, 1024);
              mma_tf32(d, da, db, idesc, (p == 2 || kb || ks) ? 1u : 0u);
            }
        }
        mma_commit(&s.a_empty);
        mma_commit(&s.tmem_full[buf]);
      }
    }
  } else if (warp < kTransformWarps) {
    // ================= transform: input rows -> (e_hi, e_lo) operand tiles + row max ==========
    const int l16 = lane & 15, half = lane >> 4;
    const float* row0 = in_row(a, f, 0);
    const float* row1 = a.H == 2 ? in_row(a, f, 1) : nullptr;
    uint8_t* ahi = reinterpret_cast<uint8_t*>(s.a_hi);
    uint8_t* alo = reinterpret_cast<uint8_t*>(s.a_lo);
    // this lane's 16-byte chunk inside a swizzled row: k-block l16/8, chunk l16%8
    const uint32_t kb_off = (uint32_t)(l16 >> 3) * (TM * 128);
    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it & 1;
      const int64_t b0 = (int64_t)(t_begin + it) * TM;
      // issue the loads of this tile before waiting for the operand buffers to drain
      float4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = warp * 16 + 2 * j + half;
        const int64_t b = b0 + r;
        x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < a.B) {
          x[j] = ldg_stream(row0 + b * KK + 4 * l16);
          if (row1) {
            const float4 z = ldg_stream(row1 + b * KK + 4 * l16);
            x[j].x += z.x; x[j].y += z.y; x[j].z += z.z; x[j].w += z.w;
          }
        }
      }
      if (it + 1 < n_tiles) {  // warm L2 with the next tile while this one is processed
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t b = b0 + TM + warp * 16 + 2 * j + half;
          if (b < a.B && (l16 & 7) == 0) {
            prefetch_l2(row0 + b * KK + 4 * l16);
            if (row1) prefetch_l2(row1 + b * KK + 4 * l16);
          }
        }
      }
      mbar_wait(&s.a_empty, (it & 1) ^ 1);
      mbar_wait(&s.tmem_empty[buf], ((it >> 1) & 1) ^ 1);  // m_buf[buf] is free again
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = warp * 16 + 2 * j + half;
        const float4 u = x[j];
        const float m = clamp_max(half_warp_max(fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w))));
        float4 hi, lo;
        if (fast_math & 1) {
          split_tf32(fast_exp(u.x - m), hi.x, lo.x);
          split_tf32(fast_exp(u.y - m), hi.y, lo.y);
          split_tf32(fast_exp(u.z - m), hi.z, lo.z);
          split_tf32(fast_exp(u.w - m), hi.w, lo.w);
        } else {
          split_tf32(expf(u.x - m), hi.x, lo.x);
          split_tf32(expf(u.y - m), hi.y, lo.y);
          split_tf32(expf(u.z - m), hi.z, lo.z);
          split_tf32(expf(u.w - m), hi.w, lo.w);
        }
        const uint32_t off = kb_off + (uint32_t)r * 128u + ((((uint32_t)l16 ^ (uint32_t)r) & 7u) << 4);
        *reinterpret_cast<float4*>(ahi + off) = hi;
        *reinterpret_cast<float4*>(alo + off) = lo;
        if (l16 == 0) s.m_buf[buf][r] = m;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.a_full);
    }
  } else {
    // ================= epilogue: TMEM -> log -> + m -> y =================
    const int q = warp & 3;  // TMEM lane quadrant this warp may read (warps 8..11 -> 0..3)
    float* stg = s.stage[q];
    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it & 1;
      const int64_t b0 = (int64_t)(t_begin + it) * TM + q * 32;
      mbar_wait(&s.tmem_full[buf], (it >> 1) & 1);
      tc_fence_after_sync();
      const float m = s.m_buf[buf][q * 32 + lane];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[16], w[16];
        tmem_ld16(taddr + c * 16, v);
        tmem_ld16(taddr + KK + c * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += w[j];
        // row `lane`, 16 columns -> staging (64-byte rows, chunk swizzled by (row>>1)&3)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 o;
          if (fast_math & 2) {
            o.x = fast_log(v[4 * j]) + m;
            o.y = fast_log(v[4 * j + 1]) + m;
            o.z = fast_log(v[4 * j + 2]) + m;
            o.w = fast_log(v[4 * j + 3]) + m;
          } else {
            o.x = logf(v[4 * j]) + m;
            o.y = logf(v[4 * j + 1]) + m;
            o.z = logf(v[4 * j + 2]) + m;
            o.w = logf(v[4 * j + 3]) + m;
          }
          *reinterpret_cast<float4*>(stg + lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2)) = o;
        }
        __syncwarp();
        // 8 rows x 64 bytes per instruction: whole sectors of y
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = i * 8 + (lane >> 2), ch = lane & 3;
          const float4 o = *reinterpret_cast<const float4*>(stg + row * 16 + ((ch ^ ((row >> 1) & 3)) << 2));
          const int64_t b = b0 + row;
          if (b < a.B)
            *reinterpret_cast<float4*>(a.y + ((int64_t)f * a.B + b) * KK + c * 16 + ch * 4) = o;
        }
        __syncwarp();
      }
      tc_fence_before_sync();
      mbar_arrive(&s.tmem_empty[buf]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

static int g_tc_enabled = -1;
static int g_tc_fast_math = 3;
void set_tensor_cores(int on) { g_tc_enabled = on ? 1 : 0; }
void set_tc_fast_math(int bits) { g_tc_fast_math = bits; }
static bool tc_disabled() {
  if (g_tc_enabled < 0) {
    const char* e = getenv("CKB_DISABLE_TC");
    g_tc_enabled = (e && e[0] == '1') ? 0 : 1;
  }
  return g_tc_enabled == 0;
}

int dense_tc_fwd(const DenseArgs& a, int F, Ctx& c) {
  if (tc_disabled() || a.Ki != KK || a.Ko != KK || a.concat || a.H < 1 || a.H > 2 || a.Kred != KK)
    return 1;
  const size_t smem = sizeof(FwdSmem) + 1024;
  static bool attr = false;
  if (!attr) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_tc_fwd_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int n_tiles = ceil_div(a.B, TM);
  // ~4 CTAs per SM over the launch (2 resident), several tiles per CTA when the batch allows
  int splits = (int)max64(1, min64(n_tiles, ceil_div(4 * kNumSMs, F)));
  const int tiles_per_cta = ceil_div(n_tiles, splits);
  splits = ceil_div(n_tiles, tiles_per_cta);
  dim3 grid(splits, F);
  dense_tc_fwd_kernel<<<grid, kThreads, smem, c.stream>>>(a, tiles_per_cta, g_tc_fast_math);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// ==========================================================================================
// Backward.  With e = exp(u - m), S = exp(y - m) (so the forward product is not recomputed) and
// r[b,o] = g[b,o] / S[b,o]:
//     d/du[b,i]  = e[b,i] * sum_o r[b,o] W[o,i]         GEMM 1  (M = samples, N = i, K = o)
//     d/dW[o,i]  = sum_b r[b,o] e[b,i]                  GEMM 2  (M = o, N = i, K = samples)
// r is contracted over o in GEMM 1 and over the samples in GEMM 2, and tf32 operands must be
// K-major (MN-major tf32 only exists with the 32-byte-atom swizzle), so the transform warps write
// r twice: as [sample][o] tiles for GEMM 1 and, after a 4x4 register transpose across lanes
// (warp shuffles), as [o][sample] tiles for GEMM 2; e is only needed as [i][sample] (GEMM 2 and
// the du epilogue, which reads it column-wise).  tf32 splits:
//   GEMM 1: r_hi W_hi -> main accumulator, r_lo W_hi + r_hi W_lo -> correction accumulator;
//   GEMM 2: ONE M=128 x N=128 instruction stream on the stacked operands [r_hi; r_lo]^T x
//           [e_hi; e_lo]^T gives all four products in separate quadrants of a 128x128
//           accumulator that stays in TMEM for the whole CTA (the lo*lo quadrant is dropped).
// dW leaves the CTA as two partial slabs per batch split which the caller reduces.
// ==========================================================================================
namespace {

constexpr int kBwdTransformWarps = 16, kBwdEpilogueWarps = 8;
constexpr int kBwdMmaWarp = kBwdTransformWarps + kBwdEpilogueWarps;
constexpr int kBwdThreads = (kBwdMmaWarp + 1) * 32;  // 800

struct __align__(1024) BwdSmem {
  float r_hi[2][TM * 32];  // [o-block][sample][32]            GEMM 1 A operand        32 KB
  float r_lo[2][TM * 32];  //                                                           32 KB
  float rT[4][128 * 32];   // [sample-block][hi o 0..63 | lo o 0..63][32 samples]       64 KB
  float eT[4][128 * 32];   // [sample-block][hi i 0..63 | lo i 0..63][32 samples]       64 KB
  float w_hi[2][KK * 32];  // W^T: [o-block][i][32 o's]        GEMM 1 B operand        16 KB
  float w_lo[2][KK * 32];  //                                                           16 KB
  uint64_t ab_full, ab_empty, e_done, d1_full[2], d1_empty[2], d2_full;
  uint32_t tmem_base;
};

constexpr int kMaxCons = 4;  // consumer rows per fold the tensor-core path sums

// 4x4 transpose across the 4 lanes that differ in their two low lane bits: on entry lane j holds
// (row j, cols 0..3); on exit it holds (rows 0..3, col j).
__device__ __forceinline__ float4 transpose4(float4 v, int j) {
  const bool p = j & 1, q = j & 2;
  float s0 = p ? v.x : v.y, s1 = p ? v.z : v.w;
  float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
  if (p) { v.x = r0; v.z = r1; } else { v.y = r0; v.w = r1; }
  s0 = q ? v.x : v.z;
  s1 = q ? v.y : v.w;
  r0 = __shfl_xor_sync(0xffffffffu, s0, 2);
  r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
  if (q) { v.x = r0; v.y = r1; } else { v.z = r0; v.w = r1; }
  return v;
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
  split_tf32(v.x, hi.x, lo.x);
  split_tf32(v.y, hi.y, lo.y);
  split_tf32(v.z, hi.z, lo.z);
  split_tf32(v.w, hi.w, lo.w);
}

__global__ void __launch_bounds__(kBwdThreads, 1)
dense_tc_bwd_kernel(DenseArgs a, int tiles_per_cta, int want_dw, int fast_math) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem& s = *reinterpret_cast<BwdSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)((a.B + TM - 1) / TM);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(n_tiles_total, t_begin + tiles_per_cta) - t_begin;
  if (n_tiles <= 0) return;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.d1_full[i], 1);
      mbar_init(&s.d1_empty[i], kBwdEpilogueWarps * 32);
    }
    mbar_init(&s.ab_full, kBwdTransformWarps);
    mbar_init(&s.ab_empty, 1);
    mbar_init(&s.e_done, kBwdEpilogueWarps * 32);
    mbar_init(&s.d2_full, 1);
    fence_barrier_init();
  }
  if (warp == kBwdMmaWarp) tmem_alloc(&s.tmem_base, 512);
  {
    // W^T image: rows i, K = o
    const float* Wf = a.W + (int64_t)f * KK * KK;  // [o][i]
    for (int idx = tid; idx < KK * KK; idx += kBwdThreads) {
      const int o = idx >> 6, i = idx & 63;
      float hi, lo;
      split_tf32(Wf[idx], hi, lo);
      const uint32_t off = swz_off(i, o & 31) >> 2;
      s.w_hi[o >> 5][off] = hi;
      s.w_lo[o >> 5][off] = lo;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;
  constexpr uint32_t kD2Col = 256;  // D1 buffers: [0,128) and [128,256); D2: [256,384)

  if (warp == kBwdMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_tf32(TM, KK, 0, 0);
      constexpr uint32_t idesc2 = make_idesc_tf32(128, 128, 0, 0);
      const uint32_t r_addr[2] = {smem_u32(s.r_hi), smem_u32(s.r_lo)};
      const uint32_t w_addr[2] = {smem_u32(s.w_hi), smem_u32(s.w_lo)};
      const uint32_t rT_addr = smem_u32(s.rT), eT_addr = smem_u32(s.eT);
      for (int it = 0; it < n_tiles; ++it) {
        const int buf = it & 1;
        mbar_wait(&s.d1_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_wait(&s.ab_full, it & 1);
        tc_fence_after_sync();
        // ---- GEMM 1: T[b,i] = sum_o r[b,o] W[o,i]
#pragma unroll
        for (int p = 0; p < 3; ++p) {  // hi*hi | lo*hi, hi*lo
          const uint32_t ab = r_addr[p == 1 ? 1 : 0], wb = w_addr[p == 2 ? 1 : 0];
          const uint32_t d = tmem_base + buf * 128 + (p == 0 ? 0 : KK);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // 8 o's per step
            const uint64_t da = make_desc(ab + (ks >> 2) * (TM * 128) + (ks & 3) * 32, 16, 1024);
            const uint64_t db = make_desc(wb + (ks >> 2) * (KK * 128) + (ks & 3) * 32, 16, 1024);
            mma_tf32(d, da, db, idesc1, (p == 2 || ks) ? 1u : 0u);
          }
        }
        mma_commit(&s.d1_full[buf]);
        // ---- GEMM 2: dW[o,i] += sum_b r[b,o] e[b,i]   (stacked hi/lo rows, 8 samples per step)
        if (want_dw) {
#pragma unroll
          for (int ks = 0; ks < TM / 8; ++ks) {
            const uint32_t o = (ks >> 2) * (128 * 128) + (ks & 3) * 32;
            mma_tf32(tmem_base + kD2Col, make_desc(rT_addr + o, 16, 1024),
                     make_desc(eT_addr + o, 16, 1024), idesc2, (it || ks) ? 1u : 0u);
          }
        }
        mma_commit(&s.ab_empty);
      }
      mma_commit(&s.d2_full);
    }
  } else if (warp < kBwdTransformWarps) {
    // ================= transform: rows -> r, r^T, e^T operand tiles =================
    const int og = lane >> 2, bsub = lane & 3;  // 16-byte chunk within a 32-column half; row in group
    const float* row0 = in_row(a, f, 0);
    const float* row1 = a.H == 2 ? in_row(a, f, 1) : nullptr;
    const float* yrow = a.y + (int64_t)f * a.B * KK;
    const float* grow[kMaxCons];
    int n_cons = 1;
    if (a.gs.cons_ptr == nullptr) {
      grow[0] = a.gs.garena + (int64_t)f * a.gs.B * KK;
    } else {
      const int c0 = a.gs.cons_ptr[f];
      n_cons = a.gs.cons_ptr[f + 1] - c0;
#pragma unroll
      for (int c = 0; c < kMaxCons; ++c)
        grow[c] = c < n_cons ? a.gs.garena + a.gs.B * a.gs.cons_rows[c0 + c] : nullptr;
    }
    uint8_t* rhi = reinterpret_cast<uint8_t*>(s.r_hi);
    uint8_t* rlo = reinterpret_cast<uint8_t*>(s.r_lo);
    uint8_t* rT = reinterpret_cast<uint8_t*>(s.rT);
    uint8_t* eT = reinterpret_cast<uint8_t*>(s.eT);
    for (int it = 0; it < n_tiles; ++it) {
      const int64_t b0 = (int64_t)(t_begin + it) * TM;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int rbase = warp * 8 + pass * 4;  // 4 consecutive samples handled by this warp
        const int r = rbase + bsub;
        const int64_t b = b0 + r;
        const bool ok = b < a.B;
        float4 xu[2], yv[2], gv[2];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          xu[ch] = yv[ch] = gv[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) {
            const int64_t o = b * KK + 32 * ch + 4 * og;
            xu[ch] = ldg_stream(row0 + o);
            if (row1) {
              const float4 z = ldg_stream(row1 + o);
              xu[ch].x += z.x; xu[ch].y += z.y; xu[ch].z += z.z; xu[ch].w += z.w;
            }
            yv[ch] = ldg_stream(yrow + o);
#pragma unroll
            for (int c = 0; c < kMaxCons; ++c)
              if (c < n_cons) {
                const float4 z = ldg_stream(grow[c] + o);
                gv[ch].x += z.x; gv[ch].y += z.y; gv[ch].z += z.z; gv[ch].w += z.w;
              }
          }
        }
        if (pass == 0) {
          // operand tiles of the previous tile must be drained (both GEMMs + the du epilogue)
          mbar_wait(&s.ab_empty, (it & 1) ^ 1);
          mbar_wait(&s.e_done, (it & 1) ^ 1);
        }
        float m = fmaxf(fmaxf(fmaxf(xu[0].x, xu[0].y), fmaxf(xu[0].z, xu[0].w)),
                        fmaxf(fmaxf(xu[1].x, xu[1].y), fmaxf(xu[1].z, xu[1].w)));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        m = clamp_max(m);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float4 e, rr, hi, lo;
          if (fast_math & 1) {
            e.x = ok ? fast_exp(xu[ch].x - m) : 0.f;
            e.y = ok ? fast_exp(xu[ch].y - m) : 0.f;
            e.z = ok ? fast_exp(xu[ch].z - m) : 0.f;
            e.w = ok ? fast_exp(xu[ch].w - m) : 0.f;
            // m - y = -log S may be positive: fast_exp is exact enough there too (|x| small)
            rr.x = gv[ch].x == 0.f ? 0.f : gv[ch].x * fast_exp(m - yv[ch].x);
            rr.y = gv[ch].y == 0.f ? 0.f : gv[ch].y * fast_exp(m - yv[ch].y);
            rr.z = gv[ch].z == 0.f ? 0.f : gv[ch].z * fast_exp(m - yv[ch].z);
            rr.w = gv[ch].w == 0.f ? 0.f : gv[ch].w * fast_exp(m - yv[ch].w);
          } else {
            e.x = ok ? expf(xu[ch].x - m) : 0.f;
            e.y = ok ? expf(xu[ch].y - m) : 0.f;
            e.z = ok ? expf(xu[ch].z - m) : 0.f;
            e.w = ok ? expf(xu[ch].w - m) : 0.f;
            rr.x = gv[ch].x == 0.f ? 0.f : gv[ch].x * expf(m - yv[ch].x);
            rr.y = gv[ch].y == 0.f ? 0.f : gv[ch].y * expf(m - yv[ch].y);
            rr.z = gv[ch].z == 0.f ? 0.f : gv[ch].z * expf(m - yv[ch].z);
            rr.w = gv[ch].w == 0.f ? 0.f : gv[ch].w * expf(m - yv[ch].w);
          }
          // r as [sample][o]
          split4(rr, hi, lo);
          const uint32_t off = (uint32_t)ch * (TM * 128) + (uint32_t)r * 128u +
                               ((((uint32_t)og ^ (uint32_t)r) & 7u) << 4);
          *reinterpret_cast<float4*>(rhi + off) = hi;
          *reinterpret_cast<float4*>(rlo + off) = lo;
          // r and e as [unit][4 consecutive samples]
          const uint32_t u = 32u * ch + 4u * og + bsub;          // unit this lane owns afterwards
          const uint32_t chunk = ((uint32_t)rbase & 31u) >> 2;   // 16-byte chunk of the 4 samples
          const uint32_t blk = ((uint32_t)rbase >> 5) * (128 * 128);
          const uint32_t off_hi = blk + u * 128u + (((chunk ^ u) & 7u) << 4);
          const uint32_t off_lo = blk + (64u + u) * 128u + (((chunk ^ (64u + u)) & 7u) << 4);
          split4(transpose4(rr, bsub), hi, lo);
          *reinterpret_cast<float4*>(rT + off_hi) = hi;
          *reinterpret_cast<float4*>(rT + off_lo) = lo;
          split4(transpose4(e, bsub), hi, lo);
          *reinterpret_cast<float4*>(eT + off_hi) = hi;
          *reinterpret_cast<float4*>(eT + off_lo) = lo;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.ab_full);
      if (it + 1 < n_tiles && og == 0) {  // warm L2 with the next tile (one lane per 128-byte line)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int64_t b = b0 + TM + warp * 8 + j * 4 + bsub;
          if (b < a.B) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              const int64_t o = b * KK + 32 * ch;
              prefetch_l2(row0 + o);
              if (row1) prefetch_l2(row1 + o);
              prefetch_l2(yrow + o);
              if (n_cons > 0) prefetch_l2(grow[0] + o);
            }
          }
        }
      }
    }
  } else {
    // ================= epilogue: du = e * T -> gin; finally dW partials =================
    const int ew = warp - kBwdTransformWarps;  // 0..7
    const int q = warp & 3;                    // TMEM lane quadrant (warps 16..23 -> 0..3,0..3)
    const int chalf = ew >> 2;                 // which 32 of the 64 columns this warp handles
    const uint8_t* eT = reinterpret_cast<const uint8_t*>(s.eT);
    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it & 1;
      const int64_t b = (int64_t)(t_begin + it) * TM + q * 32 + lane;
      mbar_wait(&s.d1_full[buf], (it >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128 + chalf * 32;
      // e[b][i] = eT[(i)][b] + eT[(64 + i)][b]: sample block q, column `lane`
      const uint32_t ecol = (uint32_t)q * (128 * 128) + ((uint32_t)lane & 3u) * 4u;
      float* dst = a.gin + ((int64_t)f * a.B + b) * KK + chalf * 32;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[16], w[16];
        tmem_ld16(taddr + c * 16, v);
        tmem_ld16(taddr + KK + c * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float o[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t i = (uint32_t)chalf * 32u + c * 16u + j + t;
            const uint32_t ohi = ecol + i * 128u + (((((uint32_t)lane >> 2) ^ i) & 7u) << 4);
            const uint32_t olo = ecol + (64u + i) * 128u + (((((uint32_t)lane >> 2) ^ (64u + i)) & 7u) << 4);
            const float e = *reinterpret_cast<const float*>(eT + ohi) + *reinterpret_cast<const float*>(eT + olo);
            o[t] = e * (v[j + t] + w[j + t]);
          }
          if (b < a.B) *reinterpret_cast<float4*>(dst + c * 16 + j) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&s.d1_empty[buf]);
      mbar_arrive(&s.e_done);
    }
    if (want_dw) {
      // D2 quadrants: rows 0..63 = r_hi^T [e_hi | e_lo], rows 64..127 = r_lo^T [e_hi | (dropped)].
      // dW[o][i] = D2[o][i] + D2[o][64+i] + D2[64+o][i]: the lower half goes through shared
      // memory (the operand tiles are dead once every MMA has completed).
      mbar_wait(&s.d2_full, 0);
      tc_fence_after_sync();
      const int row = q * 32 + lane;  // 0..127
      const int o = row & 63;
      float* xch = s.rT[0];           // [64 columns][64 + 1] exchange buffer (spans rT[0..1])
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kD2Col + chalf * 32;
      float v[32];
      tmem_ld16(taddr, v);
      tmem_ld16(taddr + 16, v + 16);
      if (row < 64) {
        float w[32];
        tmem_ld16(taddr + KK, w);
        tmem_ld16(taddr + KK + 16, w + 16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += w[j];
      } else {
        tmem_ld_wait();
        // stored [column][o] with a padded stride: the 32 lanes (32 rows o) hit 32 banks
#pragma unroll
        for (int j = 0; j < 32; ++j) xch[(chalf * 32 + j) * (KK + 1) + o] = v[j];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kBwdEpilogueWarps * 32) : "memory");
      if (row < 64) {
        float* out = a.dWp + (((int64_t)blockIdx.x * gridDim.y + f) * KK + o) * KK + chalf * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += xch[(chalf * 32 + j) * (KK + 1) + o];
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kBwdMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

void dense_tc_bwd_config(int F, int64_t B, int& splits, int& tiles_per_cta) {
  const int n_tiles = ceil_div(B, TM);
  splits = (int)max64(1, min64(n_tiles, ceil_div(2 * kNumSMs, F)));
  tiles_per_cta = ceil_div(n_tiles, splits);
  splits = ceil_div(n_tiles, tiles_per_cta);
}

bool dense_tc_bwd_ok(const DenseArgs& a) {
  return a.Ki == KK && a.Ko == KK && !a.concat && a.H >= 1 && a.H <= 2 && a.Kred == KK;
}

}  // namespace

size_t dense_tc_bwd_ws(int F, int H, int Ko, int Kred, int64_t B) {
  if (Ko != KK || Kred != KK || H > 2) return 0;
  int splits, tpc;
  dense_tc_bwd_config(F, B, splits, tpc);
  return splits > 1 ? (size_t)splits * F * KK * KK * 4 : 0;
}

int dense_tc_bwd(const DenseArgs& a_in, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes) {
  if (tc_disabled() || !dense_tc_bwd_ok(a_in) || a_in.max_cons > kMaxCons) return 1;
  DenseArgs a = a_in;
  int splits, tpc;
  dense_tc_bwd_config(F, a.B, splits, tpc);
  const size_t n = (size_t)F * KK * KK;
  a.dWp = dW;
  if (dW && splits > 1) {
    if (ws_bytes < splits * n * 4) {
      set_error("dense_tc_bwd: workspace too small (%zu < %zu)", ws_bytes, splits * n * 4);
      return CKB_ERR_WORKSPACE;
    }
    a.dWp = (float*)ws;
  }
  const size_t smem = sizeof(BwdSmem) + 1024;
  static bool attr = false;
  if (!attr) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_tc_bwd_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid(splits, F);
  dense_tc_bwd_kernel<<<grid, kBwdThreads, smem, c.stream>>>(a, tpc, dW ? 1 : 0, g_tc_fast_math);
  CKB_LAUNCH_CHECK();
  c.launches++;
  if (dW && splits > 1) return reduce_partials(a.dWp, dW, (int64_t)n, splits, c);
  return CKB_OK;
}

}  // namespace ckb
