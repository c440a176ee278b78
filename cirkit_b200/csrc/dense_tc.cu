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
          split_tf32(__expf(u.x - m), hi.x, lo.x);
          split_tf32(__expf(u.y - m), hi.y, lo.y);
          split_tf32(__expf(u.z - m), hi.z, lo.z);
          split_tf32(__expf(u.w - m), hi.w, lo.w);
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
            o.x = __logf(v[4 * j]) + m;
            o.y = __logf(v[4 * j + 1]) + m;
            o.z = __logf(v[4 * j + 2]) + m;
            o.w = __logf(v[4 * j + 3]) + m;
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
static int g_tc_fast_math = 0;
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

int dense_tc_bwd(const DenseArgs&, int, float*, Ctx&, char*, size_t) { return 1; }
size_t dense_tc_bwd_ws(int, int, int, int, int64_t) { return 0; }

}  // namespace ckb
