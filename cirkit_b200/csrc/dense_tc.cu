// Tensor-core (tcgen05 / TMEM / bulk-copy) version of the fused sum-product block for the hot
// shape Ki = Ko = 64 (Hadamard arity <= 2), sm_100a only.
//
//   y[b,o] = log( sum_i W[o,i] * exp(u[b,i] - m[b]) ) + m[b],   u = sum_h x_h,  m = max_i u
//
// One CTA owns a fold and walks over 128-sample tiles with a warp-specialised pipeline:
//
//   producer warp   : cp.async.bulk (TMA engine) of the H contiguous 128x64 input blocks of the
//                     tile into a 2-stage shared-memory ring                     [raw_full/empty]
//   transform warps : rows -> u -> max (warp shuffles) -> e = exp(u - m) -> split e into two
//                     tf32 terms (hi, lo) written as 128B-swizzled UMMA operand tiles [a_full/empty]
//   MMA thread      : D(128x64, TMEM) = e_hi W_hi^T + e_lo W_hi^T + e_hi W_lo^T
//                     (kind::tf32, three products = fp32-grade accuracy)        [tmem_full/empty]
//   epilogue warps  : tcgen05.ld D -> log -> + m -> y (two TMEM buffers, so the epilogue of tile t
//                     overlaps the transform + MMA of tile t+1)
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
constexpr int kThreads = 320;
constexpr int kTransformWarps = 4, kEpilogueWarp0 = 4, kProducerWarp = 8, kMmaWarp = 9;

struct __align__(1024) FwdSmem {
  float raw[2][2][TM * KK];  // [stage][h][row][64]                      128 KB
  float a_hi[2][TM * 32];    // [k-block][row][32] swizzled               32 KB
  float a_lo[2][TM * 32];    //                                           32 KB
  float w_hi[2][KK * 32];    // [k-block][o][32] swizzled                 16 KB
  float w_lo[2][KK * 32];    //                                           16 KB
  float m_buf[2][TM];
  uint64_t raw_full[2], raw_empty[2], a_full, a_empty, tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1) dense_tc_fwd_kernel(DenseArgs a, int tiles_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  FwdSmem& s = *reinterpret_cast<FwdSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int splits = gridDim.x;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)((a.B + TM - 1) / TM);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(n_tiles_total, t_begin + tiles_per_cta);
  const int n_tiles = t_end - t_begin;
  (void)splits;
  if (n_tiles <= 0) return;

  // ---- one-time setup: barriers, TMEM, weights
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.raw_full[i], 1);
      mbar_init(&s.raw_empty[i], kTransformWarps);
      mbar_init(&s.tmem_full[i], 1);
      mbar_init(&s.tmem_empty[i], 128);
    }
    mbar_init(&s.a_full, kTransformWarps);
    mbar_init(&s.a_empty, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&s.tmem_base, 128);
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

  if (warp == kProducerWarp) {
    // ================= bulk-copy producer =================
    if (lane == 0) {
      const float* rows[2];
      for (int h = 0; h < a.H; ++h) rows[h] = in_row(a, f, h);
      for (int it = 0; it < n_tiles; ++it) {
        const int st = it & 1;
        const int64_t b0 = (int64_t)(t_begin + it) * TM;
        const uint32_t nrows = (uint32_t)min64(TM, a.B - b0);
        mbar_wait(&s.raw_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&s.raw_full[st], a.H * nrows * KK * 4);
        for (int h = 0; h < a.H; ++h)
          bulk_g2s(s.raw[st][h], rows[h] + b0 * KK, nrows * KK * 4, &s.raw_full[st]);
      }
    }
  } else if (warp == kMmaWarp) {
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
        uint32_t acc = 0;
#pragma unroll
        for (int p = 0; p < 3; ++p) {       // hi*hi, lo*hi, hi*lo
          const uint32_t ab = a_addr[p == 1 ? 1 : 0], wb = w_addr[p == 2 ? 1 : 0];
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = make_desc(ab + kb * (TM * 128) + ks * 32, 16, 1024);
              const uint64_t db = make_desc(wb + kb * (KK * 128) + ks * 32, 16, 1024);
              mma_tf32(tmem_base + buf * KK, da, db, idesc, acc);
              acc = 1;
            }
        }
        mma_commit(&s.a_empty);
        mma_commit(&s.tmem_full[buf]);
      }
    }
  } else if (warp < kTransformWarps) {
    // ================= transform: raw rows -> (e_hi, e_lo) operand tiles + row max =================
    for (int it = 0; it < n_tiles; ++it) {
      const int st = it & 1, buf = it & 1;
      const int64_t b0 = (int64_t)(t_begin + it) * TM;
      const int nrows = (int)min64(TM, a.B - b0);
      mbar_wait(&s.raw_full[st], (it >> 1) & 1);
      mbar_wait(&s.a_empty, (it & 1) ^ 1);
      mbar_wait(&s.tmem_empty[buf], ((it >> 1) & 1) ^ 1);  // m_buf[buf] is free again
      const float* r0 = s.raw[st][0];
      const float* r1 = s.raw[st][1];
      uint8_t* ahi = reinterpret_cast<uint8_t*>(s.a_hi);
      uint8_t* alo = reinterpret_cast<uint8_t*>(s.a_lo);
#pragma unroll 2
      for (int rr = 0; rr < 32; ++rr) {
        const int r = warp * 32 + rr;
        float u0 = 0.f, u1 = 0.f;
        if (r < nrows) {
          u0 = r0[r * KK + lane];
          u1 = r0[r * KK + 32 + lane];
          if (a.H == 2) {
            u0 += r1[r * KK + lane];
            u1 += r1[r * KK + 32 + lane];
          }
        }
        const float m = clamp_max(warp_max(fmaxf(u0, u1)));
        float h0, l0, h1, l1;
        split_tf32(expf(u0 - m), h0, l0);
        split_tf32(expf(u1 - m), h1, l1);
        const uint32_t off = swz_off(r, lane);
        *reinterpret_cast<float*>(ahi + off) = h0;
        *reinterpret_cast<float*>(ahi + TM * 128 + off) = h1;
        *reinterpret_cast<float*>(alo + off) = l0;
        *reinterpret_cast<float*>(alo + TM * 128 + off) = l1;
        if (lane == 0) s.m_buf[buf][r] = m;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s.a_full);
        mbar_arrive(&s.raw_empty[st]);
      }
    }
  } else {
    // ================= epilogue: TMEM -> log -> + m -> y =================
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it & 1;
      const int64_t b0 = (int64_t)(t_begin + it) * TM;
      const int row = q * 32 + lane;
      const int64_t b = b0 + row;
      mbar_wait(&s.tmem_full[buf], (it >> 1) & 1);
      tc_fence_after_sync();
      const float m = s.m_buf[buf][row];
      float* yrow = a.y + ((int64_t)f * a.B + b) * KK;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * KK;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[16];
        tmem_ld16(taddr + c * 16, v);
        tmem_ld_wait();
        if (b < a.B) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o;
            o.x = logf(v[j]) + m;
            o.y = logf(v[j + 1]) + m;
            o.z = logf(v[j + 2]) + m;
            o.w = logf(v[j + 3]) + m;
            *reinterpret_cast<float4*>(yrow + c * 16 + j) = o;
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&s.tmem_empty[buf]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace

static int g_tc_enabled = -1;
void set_tensor_cores(int on) { g_tc_enabled = on ? 1 : 0; }
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
  // enough CTAs for ~4 per SM over the launch, but keep several tiles per CTA when possible
  int splits = (int)max64(1, min64(n_tiles, ceil_div(4 * kNumSMs, F)));
  const int tiles_per_cta = ceil_div(n_tiles, splits);
  splits = ceil_div(n_tiles, tiles_per_cta);
  dim3 grid(splits, F);
  dense_tc_fwd_kernel<<<grid, kThreads, smem, c.stream>>>(a, tiles_per_cta);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

int dense_tc_bwd(const DenseArgs&, int, float*, Ctx&, char*, size_t) { return 1; }
size_t dense_tc_bwd_ws(int, int, int, int, int64_t) { return 0; }

}  // namespace ckb
