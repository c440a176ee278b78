// Tensor-core (tcgen05 / TMEM) versions of the fused sum-product block for the hot shape
// Ki = Ko = 64 (Hadamard arity <= 2), sm_100a only.
//
//   forward   y[b,o] = log( sum_i W[o,i] * exp(u[b,i] - m[b]) ) + m[b],  u = sum_h x_h, m = max_i u
//   backward  r = g / S,  d/du[b,i] = e[b,i] * sum_o r[b,o] W[o,i],  d/dW[o,i] = sum_b r[b,o] e[b,i]
//
// All matrix products run as 3xTF32 on the 5th-generation tensor cores: every fp32 operand x is
// split into hi + lo (hi = x rounded to tf32, lo = x - hi) and hi*hi + lo*hi + hi*lo is
// accumulated in TMEM, which keeps fp32-grade accuracy (the dropped lo*lo term is 2^-22).
// Operand tiles are written by the CUDA cores straight into 128-byte-swizzled shared-memory
// tiles in the K-major UMMA layout, so nothing but the layer's inputs and outputs touches HBM.
#include "dense.cuh"
#include "sm100.cuh"
#include "tc_util.cuh"

namespace ckb {
using namespace sm100;

// Debug timeline: with CKB_OPT_TC_FAST_MATH bit 7 set, CTA (0,0) of the backward kernel records
// clock64() at phase boundaries (read back with ckb_debug_read).
__device__ long long g_dbg[512];

namespace {

constexpr int TM = 128;  // samples per tile (UMMA M)
constexpr int KK = 64;   // Ki = Ko

// Splits the fold's 64x64 weight slice into (hi, lo) swizzled K-major tiles, stacked per k-block
// as 128 rows [hi rows 0..63 | lo rows 0..63] so that ONE N=128 instruction multiplies an A tile
// with both halves (W_hi -> accumulator columns 0..63, W_lo -> 64..127) and an N=64 instruction
// on the same descriptor uses W_hi alone.
// TRANSPOSED = false: rows o, K = i  (forward:  S = e W^T)
// TRANSPOSED = true : rows i, K = o  (backward: T = r W)
constexpr uint32_t kWBlock = 128 * 128;  // bytes per k-block of the stacked weight tile
template <bool TRANSPOSED, int NTHREADS>
__device__ __forceinline__ void stage_weights(const float* Wf, uint32_t w, int tid) {
  // All of this thread's 16-byte pieces are requested before the first one is used: the stores
  // below are volatile asm, so a load inside their loop would be serialised behind them (one
  // L2 / HBM round trip per piece; the CTA's set-up took 5 us that way).
  constexpr int kPieces = KK * KK / 4;
  constexpr int PER = (kPieces + NTHREADS - 1) / NTHREADS;
  float4 v[PER];
#pragma unroll
  for (int n = 0; n < PER; ++n) {
    const int p = tid + n * NTHREADS;
    v[n] = p < kPieces ? __ldg(reinterpret_cast<const float4*>(Wf) + p) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int n = 0; n < PER; ++n) {
    const int p = tid + n * NTHREADS;
    if (p < kPieces) {
      const int o = p >> 4, i0 = (p & 15) * 4;
      float4 hi, lo;
      split4(v[n], hi, lo);
      if (TRANSPOSED) {  // rows i0..i0+3, column o
        const float h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t off = (uint32_t)(o >> 5) * kWBlock + swz_off(i0 + t, o & 31);
          sts32(w + off, h[t]);
          sts32(w + off + KK * 128, l[t]);  // (row + 64) & 7 == row & 7: same swizzle
        }
      } else {  // row o, columns i0..i0+3: one 16-byte chunk
        const uint32_t off = (uint32_t)(i0 >> 5) * kWBlock + swz_off(o, i0 & 31);
        sts128(w + off, hi);
        sts128(w + off + KK * 128, lo);
      }
    }
  }
}

// ==========================================================================================
// Forward.  One CTA owns a fold and walks over 128-sample tiles with a warp-specialised pipeline
// (two CTAs are resident per SM, so the phases of one overlap the other's):
//   transform warps : coalesced 16-byte loads of the H input rows (a half-warp per 256-byte row),
//                     u -> max (shuffles) -> e = exp(u - m) -> (hi, lo) operand tiles; the next
//                     tile's rows are prefetched into L2 meanwhile                  [a_full/empty]
//   MMA thread      : D(128x64, TMEM) = e_hi W_hi^T (+ correction accumulator e_lo W_hi^T +
//                     e_hi W_lo^T)                                              [tmem_full/empty]
//   epilogue warps  : tcgen05.ld D -> log -> + m -> staged through shared memory so that every
//                     store instruction writes whole 32-byte sectors of y (two TMEM buffers: the
//                     epilogue of tile t overlaps the transform + MMA of tile t+1)
// ==========================================================================================
constexpr int kTransformWarps = 8, kEpilogueWarps = 4;
constexpr int kMmaWarp = kTransformWarps + kEpilogueWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;  // 416

struct __align__(1024) FwdSmem {
  float a_hi[2][TM * 32];  // [k-block][row][32] swizzled               32 KB
  float a_lo[2][TM * 32];  //                                           32 KB
  float w[2][128 * 32];    // [k-block][hi o 0..63 | lo o 0..63][32] swizzled     32 KB
  float stage[kEpilogueWarps][32 * 16];  // per-warp 32 rows x 16 columns   8 KB
  float m_buf[2][TM];
  uint64_t a_full, a_empty, tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

template <bool FAST>
__global__ void __launch_bounds__(kThreads, 2) dense_tc_fwd_kernel(DenseArgs a, int tiles_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  FwdSmem& s = *reinterpret_cast<FwdSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)((a.B + TM - 1) / TM);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(n_tiles_total, t_begin + tiles_per_cta) - t_begin;
  if (n_tiles <= 0) return;

  // input row pointers first: their (dependent) index loads overlap the rest of the setup
  const float* row0 = in_row(a, f, 0);
  const float* row1 = a.H == 2 ? in_row(a, f, 1) : nullptr;

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
  // (the weights come out of the parameter ops, several launches back: no dependency on the
  //  preceding grid, whose output -- this layer's input rows -- is read after pdl_wait())
  stage_weights<false, kThreads>(a.W + (int64_t)f * KK * KK, smem_u32(s.w), tid);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;
  pdl_wait();
  pdl_launch_dependents();  // (after the wait: see common.cuh)

  if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    {  // whole warp converged; the instructions are predicated on the elected lane (a single
       // lane in a divergent branch pays ~55 clocks per tcgen05.mma instead of ~25)
      constexpr uint32_t idesc_n128 = make_idesc_tf32(TM, 2 * KK, 0, 0);
      constexpr uint32_t idesc_n64 = make_idesc_tf32(TM, KK, 0, 0);
      const uint64_t d_ahi = make_desc(smem_u32(s.a_hi), 16, 1024);
      const uint64_t d_alo = make_desc(smem_u32(s.a_lo), 16, 1024);
      const uint64_t d_w = make_desc(smem_u32(s.w), 16, 1024);
      for (int it = 0; it < n_tiles; ++it) {
        const int buf = it & 1;
        mbar_wait(&s.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_wait(&s.a_full, it & 1);
        tc_fence_after_sync();
        // The tensor core truncates when it folds a product group into the fp32 accumulator
        // (measured: ~0.6 ulp low per accumulating instruction), so the two small correction
        // products get their own accumulator (columns 64..127): only the 8 hi*hi steps touch the
        // large one, and the epilogue adds the two in round-to-nearest fp32.
        //   e_hi x [W_hi | W_lo]  (N = 128): main | correction
        //   e_lo x  W_hi          (N =  64): correction
        const uint32_t d = tmem_base + buf * 128;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_tf32_warp(d, desc_at(d_ahi, (ks >> 2) * (TM * 128) + (ks & 3) * 32),
                   desc_at(d_w, (ks >> 2) * kWBlock + (ks & 3) * 32), idesc_n128, ks ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_tf32_warp(d + KK, desc_at(d_alo, (ks >> 2) * (TM * 128) + (ks & 3) * 32),
                   desc_at(d_w, (ks >> 2) * kWBlock + (ks & 3) * 32), idesc_n64, 1u);
        mma_commit_warp(&s.a_empty);
        mma_commit_warp(&s.tmem_full[buf]);
      }
    }
  } else if (warp < kTransformWarps) {
    // ================= transform: input rows -> (e_hi, e_lo) operand tiles + row max ==========
    const int l16 = lane & 15, half = lane >> 4;
    const uint32_t ahi = smem_u32(s.a_hi), alo = smem_u32(s.a_lo);
    // this lane's 16-byte chunk inside a swizzled row: k-block l16/8, chunk l16%8
    const uint32_t kb_off = (uint32_t)(l16 >> 3) * (TM * 128);
    // this lane's element of row `warp*16 + half` of the current tile; row j adds 2 rows
    const int64_t e0 = ((int64_t)t_begin * TM + warp * 16 + half) * KK + 4 * l16;
    const float* p0 = row0 + e0;
    const float* p1 = row1 ? row1 + e0 : nullptr;
    int64_t rows_left = a.B - ((int64_t)t_begin * TM + warp * 16 + half);
    // L2 warm-up of the tile after: this warp's 16 rows are 4 KB per input = 32 lines
    const int pf_off = TM * KK + (warp * 16 - (warp * 16 + half)) * KK - 4 * l16 + lane * 32;
    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it & 1;
      // issue the loads of this tile before waiting for the operand buffers to drain
      float4 x[8];
      if (rows_left + warp * 16 + half >= TM) {  // whole tile in range (uniform over the CTA)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = ldg_stream(p0 + 2 * j * KK);
        if (p1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 z = ldg_stream(p1 + 2 * j * KK);
            x[j].x += z.x; x[j].y += z.y; x[j].z += z.z; x[j].w += z.w;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool ok = rows_left > 2 * j;
          x[j] = ldg_stream_if(p0 + 2 * j * KK, ok);
          const float4 z = ldg_stream_if(p1 + 2 * j * KK, ok && p1 != nullptr);
          x[j].x += z.x; x[j].y += z.y; x[j].z += z.z; x[j].w += z.w;
        }
      }
      if (it + 1 < n_tiles && rows_left - TM > 15 - half) {
        prefetch_l2(p0 + pf_off);
        if (p1) prefetch_l2(p1 + pf_off);
      }
      p0 += TM * KK;
      if (p1) p1 += TM * KK;
      rows_left -= TM;
      mbar_wait(&s.a_empty, (it & 1) ^ 1);
      mbar_wait(&s.tmem_empty[buf], ((it >> 1) & 1) ^ 1);  // m_buf[buf] is free again
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = warp * 16 + 2 * j + half;
        const float4 u = x[j];
        const float m = clamp_max(half_warp_max(fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w))));
        float4 e, hi, lo;
        e.x = exp_nonpos<FAST>(u.x - m);
        e.y = exp_nonpos<FAST>(u.y - m);
        e.z = exp_nonpos<FAST>(u.z - m);
        e.w = exp_nonpos<FAST>(u.w - m);
        split4(e, hi, lo);
        const uint32_t off = kb_off + (uint32_t)r * 128u + ((((uint32_t)l16 ^ (uint32_t)r) & 7u) << 4);
        sts128(ahi + off, hi);
        sts128(alo + off, lo);
        if (l16 == 0) s.m_buf[buf][r] = m;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.a_full);
    }
  } else {
    // ================= epilogue: TMEM -> log -> + m -> y =================
    const int q = warp & 3;  // TMEM lane quadrant this warp may read (warps 8..11 -> 0..3)
    const uint32_t stg = smem_u32(s.stage[q]);
    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it & 1;
      const int64_t b0 = (int64_t)(t_begin + it) * TM + q * 32;
      mbar_wait_relaxed(&s.tmem_full[buf], (it >> 1) & 1);
      tc_fence_after_sync();
      const float m = s.m_buf[buf][q * 32 + lane];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[16], w[16];
        tmem_ld16(taddr + c * 16, v);
        tmem_ld16(taddr + KK + c * 16, w);
        tmem_ld_wait();
        // row `lane`, 16 columns -> staging (64-byte rows, chunk swizzled by (row>>1)&3)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 o;
          o.x = log_<FAST>(v[4 * j] + w[4 * j]) + m;
          o.y = log_<FAST>(v[4 * j + 1] + w[4 * j + 1]) + m;
          o.z = log_<FAST>(v[4 * j + 2] + w[4 * j + 2]) + m;
          o.w = log_<FAST>(v[4 * j + 3] + w[4 * j + 3]) + m;
          sts128(stg + (uint32_t)(lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2)) * 4u, o);
        }
        __syncwarp();
        // 8 rows x 64 bytes per instruction: whole sectors of y
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = i * 8 + (lane >> 2), ch = lane & 3;
          const float4 o = lds128(stg + (uint32_t)(row * 16 + ((ch ^ ((row >> 1) & 3)) << 2)) * 4u);
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

// ==========================================================================================
// Backward.  r is contracted over o in GEMM 1 (T = r W) and over the samples in GEMM 2
// (dW = r^T e), and tf32 operands must be K-major (MN-major tf32 only exists with the 32-byte
// atom swizzle), so r is written twice: as [sample][o] tiles for GEMM 1 and, after a 4x4
// register transpose across lanes (warp shuffles), as [o][sample] tiles for GEMM 2; e is only
// needed as [i][sample] (GEMM 2 and the du epilogue, which reads it column-wise).
//   GEMM 1: r_hi W_hi -> main accumulator, r_lo W_hi + r_hi W_lo -> correction accumulator;
//   GEMM 2: ONE M=128 x N=128 instruction stream on the stacked operands [r_hi; r_lo]^T x
//           [e_hi; e_lo]^T gives all four products in separate quadrants of a 128x128
//           accumulator that stays in TMEM for the whole CTA (the lo*lo quadrant is dropped).
// 16 worker warps do both the operand transform of tile t and, once GEMM 1 has finished, its du
// epilogue (each warp: 32 samples x 16 columns), so every warp is busy in both phases and the
// registers left (one CTA per SM) hold the next tile's loads while the tensor core works.
// du is staged through the r tile (dead after GEMM 1) so that stores cover whole sectors.
// ==========================================================================================
constexpr int kWorkers = 16;
constexpr int kBwdThreads = kWorkers * 32;  // 512: 128 registers per thread, no spills
constexpr int kGemm1Warp = 0, kGemm2Warp = 1;  // lane 0 of these workers issues the MMAs

struct __align__(1024) BwdSmem {
  float r_hi[2][TM * 32];  // [o-block][sample][32]            GEMM 1 A operand        32 KB
  float r_lo[2][TM * 32];  //                                                           32 KB
  float rT[4][128 * 32];   // [sample-block][hi o 0..63 | lo o 0..63][32 samples]       64 KB
  float eT[4][128 * 32];   // [sample-block][hi i 0..63 | lo i 0..63][32 samples]       64 KB
  float w[2][128 * 32];    // W^T: [o-block][hi i 0..63 | lo i 0..63][32 o's]   GEMM 1 B    32 KB
  uint64_t ab_full, ab_empty, d1_full, d2_full;
  uint32_t tmem_base;
};

constexpr int kMaxCons = 4;  // consumer rows per fold the tensor-core path sums

struct BwdLoads {
  float4 x0[2][2], x1[2][2], y[2][2], g[2][2];  // [pass][column half]
};

// What a worker thread needs to stream its share of the tiles: pointers to its first element of
// the tile to be loaded next (row warp*8 + bsub, columns 4*og; pass p adds 4 rows, column half
// ch adds 32 columns) and how many rows are left below that row.
struct BwdStream {
  const float *x0, *x1, *y, *g;  // x1 / g may be null (arity 1 / no consumer)
  int64_t rows_left;             // B - (row of pass 0 in the next tile)
};

// MODE 2: the tile is complete, 1: rows must be checked, 0: nothing to load
template <int MODE>
__device__ __forceinline__ void bwd_load_pass(BwdLoads& L, const BwdStream& st, int p) {
  if (MODE == 0) return;
  const bool ok = MODE == 2 || st.rows_left > 4 * p;
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    const int o = p * 4 * KK + ch * 32;
    if (MODE == 2) {
      L.x0[p][ch] = ldg_stream(st.x0 + o);
      if (st.x1) L.x1[p][ch] = ldg_stream(st.x1 + o);
      L.y[p][ch] = ldg_stream(st.y + o);
      if (st.g) L.g[p][ch] = ldg_stream(st.g + o);
    } else {
      L.x0[p][ch] = ldg_stream_if(st.x0 + o, ok);
      L.x1[p][ch] = ldg_stream_if(st.x1 + o, ok && st.x1 != nullptr);
      L.y[p][ch] = ldg_stream_if(st.y + o, ok);
      L.g[p][ch] = ldg_stream_if(st.g + o, ok && st.g != nullptr);
    }
  }
}

// Shared-memory byte offsets of a worker thread's stores (all other terms are immediates).
struct BwdOffsets {
  uint32_t nat[2];  // [pass]: r as [sample][o]
  uint32_t tr[2];   // [pass]: r / e as [unit][4 consecutive samples]
};

// One tile of the operand transform.  Rows that do not exist were loaded as zeros, which makes
// their r rows zero (g = 0), so they drop out of both GEMMs without any further masking.
template <bool FAST, int NEXT>
__device__ __forceinline__ void bwd_transform(BwdLoads& L, const BwdStream& st, const BwdOffsets& off,
                                              uint32_t smem_rhi, int bsub) {
  constexpr uint32_t kRlo = offsetof(BwdSmem, r_lo) - offsetof(BwdSmem, r_hi);
  constexpr uint32_t kRT = offsetof(BwdSmem, rT) - offsetof(BwdSmem, r_hi);
  constexpr uint32_t kET = offsetof(BwdSmem, eT) - offsetof(BwdSmem, r_hi);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float4 xu[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      xu[ch].x = L.x0[p][ch].x + L.x1[p][ch].x;
      xu[ch].y = L.x0[p][ch].y + L.x1[p][ch].y;
      xu[ch].z = L.x0[p][ch].z + L.x1[p][ch].z;
      xu[ch].w = L.x0[p][ch].w + L.x1[p][ch].w;
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
      e.x = exp_nonpos<FAST>(xu[ch].x - m);
      e.y = exp_nonpos<FAST>(xu[ch].y - m);
      e.z = exp_nonpos<FAST>(xu[ch].z - m);
      e.w = exp_nonpos<FAST>(xu[ch].w - m);
      rr.x = L.g[p][ch].x * exp_capped<FAST>(m - L.y[p][ch].x);
      rr.y = L.g[p][ch].y * exp_capped<FAST>(m - L.y[p][ch].y);
      rr.z = L.g[p][ch].z * exp_capped<FAST>(m - L.y[p][ch].z);
      rr.w = L.g[p][ch].w * exp_capped<FAST>(m - L.y[p][ch].w);
      const uint32_t a_nat = smem_rhi + off.nat[p] + ch * (TM * 128);
      const uint32_t a_tr = smem_rhi + off.tr[p] + ch * (32 * 128);
      split4(rr, hi, lo);
      sts128(a_nat, hi);
      sts128(a_nat + kRlo, lo);
      split4(transpose4(rr, bsub), hi, lo);
      sts128(a_tr + kRT, hi);
      sts128(a_tr + kRT + 64 * 128, lo);  // (64 + u) & 7 == u & 7: same swizzle
      split4(transpose4(e, bsub), hi, lo);
      sts128(a_tr + kET, hi);
      sts128(a_tr + kET + 64 * 128, lo);
    }
    // this pass's registers are free again: refill them with the next tile's rows right away,
    // so the requests are spread over the transform and have a whole tile period to land
    bwd_load_pass<NEXT>(L, st, p);
  }
}

template <bool FAST>
__global__ void __launch_bounds__(kBwdThreads, 1)
dense_tc_bwd_kernel(DenseArgs a, int tiles_per_cta, int want_dw, int flags) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem& s = *reinterpret_cast<BwdSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)((a.B + TM - 1) / TM);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(n_tiles_total, t_begin + tiles_per_cta) - t_begin;
  if (n_tiles <= 0) return;
  const bool dbg = (flags & 128) && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
#ifdef CKB_TIMELINE
#define DBG(slot) do { if (dbg) g_dbg[slot] = clock64(); } while (0)
#else
#define DBG(slot) do { (void)dbg; } while (0)
#endif
  if (tid == 0) DBG(0);

  // row pointers first: their dependent index loads overlap the rest of the setup
  const float* row0 = in_row(a, f, 0);
  const float* row1 = a.H == 2 ? in_row(a, f, 1) : nullptr;
  const float* yrow = a.y + (int64_t)f * a.B * KK;
  const float* grow0 = nullptr;
  int n_cons = 1, cons0 = 0;
  if (a.gs.cons_ptr == nullptr) {
    grow0 = a.gs.garena + (int64_t)f * a.gs.B * KK;
  } else {
    cons0 = a.gs.cons_ptr[f];
    n_cons = a.gs.cons_ptr[f + 1] - cons0;
    if (n_cons > 0) grow0 = a.gs.garena + a.gs.B * a.gs.cons_rows[cons0];
  }

  if (tid == 0) {
    mbar_init(&s.ab_full, kWorkers);
    mbar_init(&s.ab_empty, want_dw ? 2 : 1);  // one commit per GEMM
    mbar_init(&s.d1_full, 1);
    mbar_init(&s.d2_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, 256);
  stage_weights<true, kBwdThreads>(a.W + (int64_t)f * KK * KK, smem_u32(s.w), tid);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;
  constexpr uint32_t kD2Col = 128;  // D1: [0,128) (main | correction); D2: [128,256)
  if (tid == 0) DBG(1);

  {
    // ================= workers: operand transform, then du epilogue =================
    const int og = lane >> 2, bsub = lane & 3;  // 16-byte chunk within a 32-column half; row in group
    const int q = warp & 3;                     // TMEM lane quadrant for the epilogue
    const int cg = warp >> 2;                   // epilogue column group: columns 16*cg .. 16*cg+15
    const uint32_t rhi = smem_u32(s.r_hi);
    const uint32_t eT = smem_u32(s.eT);
    const uint32_t stg = rhi + (uint32_t)warp * 2048u;  // 32 rows x 16 columns, inside the dead r tile

    BwdOffsets off;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const uint32_t r = warp * 8 + p * 4 + bsub;           // sample row inside the tile
      off.nat[p] = r * 128u + ((((uint32_t)og ^ r) & 7u) << 4);
      const uint32_t u = 4u * og + bsub;                    // unit (mod 32) this lane owns afterwards
      const uint32_t chunk = ((r & 31u) >> 2);              // 16-byte chunk of the 4 samples
      off.tr[p] = (r >> 5) * (128 * 128) + u * 128u + (((chunk ^ u) & 7u) << 4);
    }
    BwdStream st;
    {
      const int64_t row = (int64_t)t_begin * TM + warp * 8 + bsub;
      const int64_t e0 = row * KK + 4 * og;
      st.x0 = row0 + e0;
      st.x1 = row1 ? row1 + e0 : nullptr;
      st.y = yrow + e0;
      st.g = n_cons > 0 ? grow0 + e0 : nullptr;
      st.rows_left = a.B - row;
    }
    auto advance = [&]() {
      st.x0 += TM * KK;
      if (st.x1) st.x1 += TM * KK;
      st.y += TM * KK;
      if (st.g) st.g += TM * KK;
      st.rows_left -= TM;
    };

    // du stores: rows q*32 + (lane>>2) + 8i of the tile, columns cg*16 + (lane&3)*4 ..+3
    float* pdu = a.gin + ((int64_t)f * a.B + (int64_t)t_begin * TM + q * 32 + (lane >> 2)) * KK +
                 cg * 16 + (lane & 3) * 4;
    int64_t du_left = a.B - ((int64_t)t_begin * TM + q * 32 + (lane >> 2));
    const bool store_du = !(flags & 32);

    BwdLoads L;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
        L.x0[p][ch] = L.x1[p][ch] = L.y[p][ch] = L.g[p][ch] = make_float4(0.f, 0.f, 0.f, 0.f);
    bwd_load_pass<1>(L, st, 0);
    bwd_load_pass<1>(L, st, 1);
    advance();
    for (int it = 0; it < n_tiles; ++it) {
      const int64_t b0 = (int64_t)(t_begin + it) * TM;
      if (warp == 0 && it < 4) DBG(16 + it * 8 + 2);
      // every worker has finished the previous epilogue (reads of eT / the staging area) and both
      // GEMMs of the previous tile have completed: the operand tiles may be overwritten
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory");
      mbar_wait(&s.ab_empty, (it & 1) ^ 1);
      if (warp == 0 && it < 4) DBG(16 + it * 8 + 3);
      if (n_cons > 1) {  // rare (DAG-shaped circuits): add the other consumers' rows to g
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int64_t b = b0 + warp * 8 + p * 4 + bsub;
          for (int c = 1; c < n_cons; ++c) {
            const float* gr = a.gs.garena + a.gs.B * a.gs.cons_rows[cons0 + c] + b * KK + 4 * og;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              const float4 z = ldg_stream_if(gr + 32 * ch, b < a.B);
              L.g[p][ch].x += z.x; L.g[p][ch].y += z.y; L.g[p][ch].z += z.z; L.g[p][ch].w += z.w;
            }
          }
        }
      }
      // The next tile's rows are requested AFTER the proxy fence: fence.proxy.async also waits for
      // the global loads the thread has in flight (measured: -6 % on the backward launches);
      // bit 4 of CKB_OPT_TC_FAST_MATH restores the refill inside the transform for A/B runs.
      const bool late = (flags & 16) == 0;
#ifdef CKB_TIMELINE
      // experiments: bit 6 = never reload (compute on stale registers), bit 3 = no math (the loads
      // are consumed by a dummy reduction)
      if (flags & 8) {
        float acc = 0.f;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int ch = 0; ch < 2; ++ch)
            acc += L.x0[p][ch].x + L.x1[p][ch].y + L.y[p][ch].z + L.g[p][ch].w + L.x0[p][ch].w;
        if (acc == 1234.5f) sts32(rhi, acc);
        if (it + 1 < n_tiles) {
          if (st.rows_left >= TM - warp * 8 - bsub) { bwd_load_pass<2>(L, st, 0); bwd_load_pass<2>(L, st, 1); }
          else { bwd_load_pass<1>(L, st, 0); bwd_load_pass<1>(L, st, 1); }
        }
      } else if (flags & 64) bwd_transform<FAST, 0>(L, st, off, rhi, bsub);
      else
#endif
      if (it + 1 >= n_tiles || late) bwd_transform<FAST, 0>(L, st, off, rhi, bsub);
      else if (st.rows_left >= TM - warp * 8 - bsub) bwd_transform<FAST, 2>(L, st, off, rhi, bsub);
      else bwd_transform<FAST, 1>(L, st, off, rhi, bsub);
      if (!late) advance();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.ab_full);
      if (late) {
        if (it + 1 < n_tiles) {
          if (st.rows_left >= TM - warp * 8 - bsub) { bwd_load_pass<2>(L, st, 0); bwd_load_pass<2>(L, st, 1); }
          else { bwd_load_pass<1>(L, st, 0); bwd_load_pass<1>(L, st, 1); }
        }
        advance();
      }
      if (warp == 0 && it < 4) DBG(16 + it * 8 + 4);
      // ---- MMA issue.  There is no dedicated MMA warp (a 17th warp would cap the kernel at 96
      // registers per thread): lane 0 of worker 0 issues GEMM 1 and lane 0 of worker 1 GEMM 2
      // once every worker has arrived; both workers would be waiting for GEMM 1 anyway.
      if (warp == kGemm1Warp) {
        {
          mbar_wait(&s.ab_full, it & 1);
          tc_fence_after_sync();
          if (it < 4 && lane == 0) DBG(16 + it * 8 + 0);
          // GEMM 1: T[b,i] = sum_o r[b,o] W[o,i]
          //   r_hi x [W_hi | W_lo]  (N = 128): main | correction;  r_lo x W_hi (N = 64): correction
          constexpr uint32_t idesc_n128 = make_idesc_tf32(TM, 2 * KK, 0, 0);
          constexpr uint32_t idesc_n64 = make_idesc_tf32(TM, KK, 0, 0);
          constexpr uint32_t kRlo = offsetof(BwdSmem, r_lo) - offsetof(BwdSmem, r_hi);
          constexpr uint32_t kW = offsetof(BwdSmem, w) - offsetof(BwdSmem, r_hi);
          const uint64_t d_r = make_desc(rhi, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // 8 o's per step
            mma_tf32_warp(tmem_base, desc_at(d_r, (ks >> 2) * (TM * 128) + (ks & 3) * 32),
                     desc_at(d_r, kW + (ks >> 2) * kWBlock + (ks & 3) * 32), idesc_n128, ks ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_tf32_warp(tmem_base + KK, desc_at(d_r, kRlo + (ks >> 2) * (TM * 128) + (ks & 3) * 32),
                     desc_at(d_r, kW + (ks >> 2) * kWBlock + (ks & 3) * 32), idesc_n64, 1u);
          mma_commit_warp(&s.d1_full);
          mma_commit_warp(&s.ab_empty);
          if (it < 4 && lane == 0) DBG(16 + it * 8 + 1);
        }
        __syncwarp();
      } else if (warp == kGemm2Warp && want_dw) {
        {
          mbar_wait(&s.ab_full, it & 1);
          tc_fence_after_sync();
          // GEMM 2: dW[o,i] += sum_b r[b,o] e[b,i]   (stacked hi/lo rows, 8 samples per step)
          constexpr uint32_t idesc2 = make_idesc_tf32(128, 128, 0, 0);
          constexpr uint32_t kRT = offsetof(BwdSmem, rT) - offsetof(BwdSmem, r_hi);
          constexpr uint32_t kETo = offsetof(BwdSmem, eT) - offsetof(BwdSmem, r_hi);
          const uint64_t d_r = make_desc(rhi, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < TM / 8; ++ks) {
            const uint32_t o = (ks >> 2) * (128 * 128) + (ks & 3) * 32;
            mma_tf32_warp(tmem_base + kD2Col, desc_at(d_r, kRT + o), desc_at(d_r, kETo + o), idesc2,
                     (it || ks) ? 1u : 0u);
          }
          mma_commit_warp(&s.ab_empty);
          if (it + 1 == n_tiles) mma_commit_warp(&s.d2_full);
        }
        __syncwarp();
      }
      // ---- du epilogue: rows 32q..32q+31, columns 16cg..16cg+15
      mbar_wait(&s.d1_full, it & 1);
      tc_fence_after_sync();
      if (warp == 0 && it < 4) DBG(16 + it * 8 + 5);
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + cg * 16;
        // e[b][i] = eT[i][b] + eT[64 + i][b]: sample block q, column `lane`
        const uint32_t ecol = eT + (uint32_t)q * (128 * 128) + ((uint32_t)lane & 3u) * 4u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // 8 columns at a time keeps the register footprint small
          float v[8], w[8];
          tmem_ld8(taddr + 8 * h, v);
          tmem_ld8(taddr + KK + 8 * h, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t i = (uint32_t)cg * 16u + 8 * h + 4 * j + t;
              const uint32_t ohi = ecol + i * 128u + (((((uint32_t)lane >> 2) ^ i) & 7u) << 4);
              const float e = lds32(ohi) + lds32(ohi + 64u * 128u);
              o[t] = e * (v[4 * j + t] + w[4 * j + t]);
            }
            sts128(stg + (uint32_t)(lane * 16 + (((2 * h + j) ^ ((lane >> 1) & 3)) << 2)) * 4u,
                   make_float4(o[0], o[1], o[2], o[3]));
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = i * 8 + (lane >> 2), ch = lane & 3;
          const float4 o = lds128(stg + (uint32_t)(row * 16 + ((ch ^ ((row >> 1) & 3)) << 2)) * 4u);
          if (du_left > 8 * i && store_du) *reinterpret_cast<float4*>(pdu + i * 8 * KK) = o;
        }
        pdu += TM * KK;
        du_left -= TM;
      }
      tc_fence_before_sync();
      if (warp == 0 && it < 4) DBG(16 + it * 8 + 6);
    }
    if (want_dw) {
      // D2 quadrants: rows 0..63 = r_hi^T [e_hi | e_lo], rows 64..127 = r_lo^T [e_hi | (dropped)].
      // dW[o][i] = D2[o][i] + D2[o][64+i] + D2[64+o][i]: the lower half goes through shared
      // memory (the operand tiles are dead once every MMA has completed).
      mbar_wait(&s.d2_full, 0);
      tc_fence_after_sync();
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory");  // staging area is idle
      const int row = q * 32 + lane;  // 0..127
      const int o = row & 63;
      const uint32_t xch = smem_u32(s.rT);  // [64 columns][64 + 1] exchange buffer
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kD2Col + cg * 16;
      float v[16];
      tmem_ld16(taddr, v);
      if (row < 64) {
        float w[16];
        tmem_ld16(taddr + KK, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += w[j];
      } else {
        tmem_ld_wait();
        // stored [column][o] with a padded stride: the 32 lanes (32 rows o) hit 32 banks
#pragma unroll
        for (int j = 0; j < 16; ++j) sts32(xch + (uint32_t)((cg * 16 + j) * (KK + 1) + o) * 4u, v[j]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory");
      if (row < 64) {
        float* out = a.dWp + (((int64_t)blockIdx.x * gridDim.y + f) * KK + o) * KK + cg * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += lds32(xch + (uint32_t)((cg * 16 + j) * (KK + 1) + o) * 4u);
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  if (tid == 0) DBG(2);
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) DBG(3);
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
#undef DBG
}


// ==========================================================================================
// Backward, second version (round 2).  What the B200 probes of round 2 established
// (scripts/micro/mn_major_probe.cu, umma_probe2.cu, tmem_shape_probe.cu):
//   * kind::tf32 accepts MN-major A and B operands with the SWIZZLE_128B_BASE32B layout (rows of
//     128 bytes per k, 32-byte chunks xor (k & 3)), so GEMM 2 (dW = r^T e, contraction over the
//     samples) reads r and e as plain [sample][unit] rows: no register transposes;
//   * K-major operands do NOT accept that layout, so GEMM 1 (T = r W, contraction over the units)
//     cannot share the image: its A operand r goes to TENSOR MEMORY instead (tcgen05.st, TS-form
//     MMA), which also takes 128 KB of operand traffic per tile off the shared-memory port;
//   * the 16x256b shape of tcgen05.ld/st is the mma C-fragment mapping: thread t of a warp owns
//     rows t/4 and t/4+8 of a 16-row slab and, in every group of 8 columns, columns 2(t%4), +1.
// One thread <-> element ownership therefore serves everything: 8-byte global loads (a warp
// instruction covers 8 rows x 32 bytes = whole sectors), r -> TMEM, r and e -> the MN-major images
// (conflict-free 8-byte stores), the read-back of T, and the du stores (whole sectors again) --
// e never leaves the registers between the transform and the epilogue, and du is not staged.
//   warp w: TMEM quadrant q = w & 3 (rows 32q..), rows +16 * ((w >> 2) & 1), columns 32 * (w >> 3);
//   the two warps that share rows (w, w ^ 8) exchange their half-row maxima through shared memory.
// ==========================================================================================
struct __align__(1024) Bwd2Smem {
  float r_mn[4][TM * 32];  // [r_hi units 0..31 | r_hi 32..63 | r_lo 0..31 | r_lo 32..63][sample][32]  64 KB
  float e_mn[4][TM * 32];  // same for e                                                            64 KB
  float w[2][128 * 32];    // W^T: [o-block][hi i 0..63 | lo i 0..63][32 o's]   GEMM 1 B             32 KB
  float mbuf[2][2][TM];    // [tile parity][column half][row]: half-row maxima
  uint64_t ab_full, ab_empty, d1_full, d2_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 v;
  asm("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_stream2_if(const float* p, bool ok) {
  float2 v;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
      "@q ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];\n\t}"
      : "=&f"(v.x), "=&f"(v.y)
      : "l"(p), "r"((int)ok));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
// 16 TMEM lanes x 32 columns <-> 16 registers per thread: reg 4n + 2a + c = (row t/4 + 8a,
// column 8n + 2(t%4) + c)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
      "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}

// a thread's share of a tile: [row a][column group n] pairs of consecutive columns
struct Bwd2Loads {
  float2 x0[2][4], x1[2][4], y[2][4], g[2][4];
};
struct Bwd2Stream {
  const float *x0, *x1, *y, *g;  // first element of the thread's share of the NEXT tile to load
  int64_t rows_left;             // B - (row a = 0 of that tile)
};
// MODE 2: the tile is complete, 1: rows must be checked
template <int MODE>
__device__ __forceinline__ void bwd2_load(Bwd2Loads& L, const Bwd2Stream& st) {
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const bool ok = MODE == 2 || st.rows_left > 8 * a;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int o = a * 8 * KK + 8 * n;
      if (MODE == 2) {
        L.x0[a][n] = ldg_stream2(st.x0 + o);
        if (st.x1) L.x1[a][n] = ldg_stream2(st.x1 + o);
        L.y[a][n] = ldg_stream2(st.y + o);
        if (st.g) L.g[a][n] = ldg_stream2(st.g + o);
      } else {
        L.x0[a][n] = ldg_stream2_if(st.x0 + o, ok);
        L.x1[a][n] = ldg_stream2_if(st.x1 + o, ok && st.x1 != nullptr);
        L.y[a][n] = ldg_stream2_if(st.y + o, ok);
        L.g[a][n] = ldg_stream2_if(st.g + o, ok && st.g != nullptr);
      }
    }
  }
}

template <bool FAST>
__global__ void __launch_bounds__(kBwdThreads, 1)
dense_tc_bwd2_kernel(DenseArgs a, int tiles_per_cta, int want_dw, int flags) {
  extern __shared__ uint8_t smem_raw[];
  Bwd2Smem& s = *reinterpret_cast<Bwd2Smem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)((a.B + TM - 1) / TM);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(n_tiles_total, t_begin + tiles_per_cta) - t_begin;
  if (n_tiles <= 0) return;
#ifdef CKB_TIMELINE
  // timeline of warps 0 and 9 of CTA (0, gridDim.y / 2): clock64() at the phase boundaries
  const bool dbg = (flags & 128) && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2 && (tid & 31) == 0 &&
                   ((tid >> 5) == 0 || (tid >> 5) == 9);
  const int dbg_base = (tid >> 5) == 0 ? 16 : 272;
#define DBG2(it, slot) do { if (dbg && (it) < 15) g_dbg[dbg_base + (it) * 16 + (slot)] = clock64(); } while (0)
  if (dbg && (tid >> 5) == 0) g_dbg[0] = clock64();
#else
#define DBG2(it, slot) do { } while (0)
#endif

  const float* row0 = in_row(a, f, 0);
  const float* row1 = a.H == 2 ? in_row(a, f, 1) : nullptr;
  const float* yrow = a.y + (int64_t)f * a.B * KK;
  const float* grow0 = nullptr;
  int n_cons = 1, cons0 = 0;
  if (a.gs.cons_ptr == nullptr) {
    grow0 = a.gs.garena + (int64_t)f * a.gs.B * KK;
  } else {
    cons0 = a.gs.cons_ptr[f];
    n_cons = a.gs.cons_ptr[f + 1] - cons0;
    if (n_cons > 0) grow0 = a.gs.garena + a.gs.B * a.gs.cons_rows[cons0];
  }

  if (tid == 0) {
    mbar_init(&s.ab_full, kWorkers);
    mbar_init(&s.ab_empty, 1);  // GEMM 2 has finished reading the shared-memory images
    mbar_init(&s.d1_full, 1);
    mbar_init(&s.d2_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);
  stage_weights<true, kBwdThreads>(a.W + (int64_t)f * KK * KK, smem_u32(s.w), tid);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;
  // TMEM columns: D1 [0,128) (main | correction), D2 [128,256), r_hi [256,320), r_lo [320,384)
  constexpr uint32_t kD2Col = 128, kRhiCol = 256, kRloCol = 320;

  const int q = warp & 3, half = (warp >> 2) & 1, chalf = warp >> 3;
  const int r8 = lane >> 2, c2 = lane & 3;
  const int trow = q * 32 + half * 16 + r8;  // row a = 0 inside the tile; a = 1 adds 8
  const uint32_t lane_addr = (uint32_t)(q * 32 + half * 16) << 16;
  const uint32_t mn_base = smem_u32(s.r_mn);
  constexpr uint32_t kEmn = offsetof(Bwd2Smem, e_mn) - offsetof(Bwd2Smem, r_mn);
  constexpr uint32_t kLo = 2 * TM * 128;  // hi -> lo inside an image
  // byte offset of (row a, column group n) of this thread inside an MN-major image:
  //   atom chalf, row * 128, 32-byte chunk n ^ (row & 3), 8 bytes per thread
  uint32_t mn_off[2];
#pragma unroll
  for (int aa = 0; aa < 2; ++aa) mn_off[aa] = (uint32_t)chalf * (TM * 128) + (uint32_t)(trow + 8 * aa) * 128u + c2 * 8u;
  const uint32_t rsw = (uint32_t)(trow & 3);  // (row & 3) is the same for a = 0 and a = 1

  Bwd2Stream st;
  {
    const int64_t row = (int64_t)t_begin * TM + trow;
    const int64_t e0 = row * KK + 32 * chalf + 2 * c2;
    st.x0 = row0 + e0;
    st.x1 = row1 ? row1 + e0 : nullptr;
    st.y = yrow + e0;
    st.g = n_cons > 0 ? grow0 + e0 : nullptr;
    st.rows_left = a.B - row;
  }
  auto advance = [&]() {
    st.x0 += TM * KK;
    if (st.x1) st.x1 += TM * KK;
    st.y += TM * KK;
    if (st.g) st.g += TM * KK;
    st.rows_left -= TM;
  };
  float* pdu = a.gin + ((int64_t)f * a.B + (int64_t)t_begin * TM + trow) * KK + 32 * chalf + 2 * c2;
  int64_t du_left = a.B - ((int64_t)t_begin * TM + trow);

  Bwd2Loads L;
#pragma unroll
  for (int aa = 0; aa < 2; ++aa)
#pragma unroll
    for (int n = 0; n < 4; ++n) L.x0[aa][n] = L.x1[aa][n] = L.y[aa][n] = L.g[aa][n] = make_float2(0.f, 0.f);
  bwd2_load<1>(L, st);
  advance();

  for (int it = 0; it < n_tiles; ++it) {
    const int64_t b0 = (int64_t)(t_begin + it) * TM;
    DBG2(it, 0);
    if (n_cons > 1) {  // rare (DAG-shaped circuits): add the other consumers' rows to g
#pragma unroll
      for (int aa = 0; aa < 2; ++aa) {
        const int64_t b = b0 + trow + 8 * aa;
        for (int c = 1; c < n_cons; ++c) {
          const float* gr = a.gs.garena + a.gs.B * a.gs.cons_rows[cons0 + c] + b * KK + 32 * chalf + 2 * c2;
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const float2 z = ldg_stream2_if(gr + 8 * n, b < a.B);
            L.g[aa][n].x += z.x; L.g[aa][n].y += z.y;
          }
        }
      }
    }
    // ---- transform: u, row max (exchanged with the warp that holds the other 32 columns), e, r
    float e[2][4][2];  // stays in registers until the du epilogue
    float rr[2][4][2];
    {
      float mloc[2];
#pragma unroll
      for (int aa = 0; aa < 2; ++aa) {
        float m = -INFINITY;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          e[aa][n][0] = L.x0[aa][n].x + L.x1[aa][n].x;
          e[aa][n][1] = L.x0[aa][n].y + L.x1[aa][n].y;
          m = fmaxf(m, fmaxf(e[aa][n][0], e[aa][n][1]));
        }
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        mloc[aa] = m;
        if (c2 == 0) s.mbuf[it & 1][chalf][trow + 8 * aa] = m;
      }
      DBG2(it, 1);
      asm volatile("bar.sync %0, 64;" ::"r"(2 + (warp & 7)) : "memory");
      DBG2(it, 2);
#pragma unroll
      for (int aa = 0; aa < 2; ++aa) {
        const float m = clamp_max(fmaxf(mloc[aa], s.mbuf[it & 1][chalf ^ 1][trow + 8 * aa]));
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const float ux = e[aa][n][0], uy = e[aa][n][1];
          e[aa][n][0] = exp_nonpos<FAST>(ux - m);
          e[aa][n][1] = exp_nonpos<FAST>(uy - m);
          rr[aa][n][0] = L.g[aa][n].x * exp_capped<FAST>(m - L.y[aa][n].x);
          rr[aa][n][1] = L.g[aa][n].y * exp_capped<FAST>(m - L.y[aa][n].y);
        }
      }
    }
    DBG2(it, 3);
    // GEMM 2 of the previous tile has finished reading the images (GEMM 1 and the epilogue of the
    // previous tile are behind this thread in program order)
    mbar_wait(&s.ab_empty, (it & 1) ^ 1);
    DBG2(it, 4);
    {
      float hi[16], lo[16];
#pragma unroll
      for (int aa = 0; aa < 2; ++aa)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int c = 0; c < 2; ++c) split_tf32(rr[aa][n][c], hi[4 * n + 2 * aa + c], lo[4 * n + 2 * aa + c]);
      tmem_st_16x256b_x4(tmem_base + lane_addr + kRhiCol + 32 * chalf, hi);
      tmem_st_16x256b_x4(tmem_base + lane_addr + kRloCol + 32 * chalf, lo);
#pragma unroll
      for (int aa = 0; aa < 2; ++aa)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const uint32_t o = mn_base + mn_off[aa] + (((uint32_t)n ^ rsw) << 5);
          sts64(o, hi[4 * n + 2 * aa], hi[4 * n + 2 * aa + 1]);
          sts64(o + kLo, lo[4 * n + 2 * aa], lo[4 * n + 2 * aa + 1]);
        }
#pragma unroll
      for (int aa = 0; aa < 2; ++aa)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          float h0, l0, h1, l1;
          split_tf32(e[aa][n][0], h0, l0);
          split_tf32(e[aa][n][1], h1, l1);
          const uint32_t o = mn_base + kEmn + mn_off[aa] + (((uint32_t)n ^ rsw) << 5);
          sts64(o, h0, h1);
          sts64(o + kLo, l0, l1);
        }
    }
    tmem_st_wait();
#ifdef CKB_TIMELINE
    if (!(flags & 2048))
#endif
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.ab_full);
    DBG2(it, 5);
    // next tile's rows: requested after the proxy fence (which waits for loads in flight)
    if (it + 1 < n_tiles && !(flags & 64)) {
      if (st.rows_left >= TM - trow) bwd2_load<2>(L, st);
      else bwd2_load<1>(L, st);
    }
    advance();
    DBG2(it, 6);
    // ---- MMA issue: lane-elected instructions from converged warps 0 (GEMM 1) and 1 (GEMM 2)
    if (warp == kGemm1Warp) {
      mbar_wait(&s.ab_full, it & 1);
      tc_fence_after_sync();
      DBG2(it, 7);
      // GEMM 1: T[b,i] = sum_o r[b,o] W[o,i], A = r from tensor memory
      //   r_hi x [W_hi | W_lo]  (N = 128): main | correction;  r_lo x W_hi (N = 64): correction
      constexpr uint32_t idesc_n128 = make_idesc_tf32(TM, 2 * KK, 0, 0);
      constexpr uint32_t idesc_n64 = make_idesc_tf32(TM, KK, 0, 0);
      const uint64_t d_w = make_desc(smem_u32(s.w), 16, 1024);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        mma_tf32_ts_warp(tmem_base, tmem_base + kRhiCol + 8 * ks,
                         desc_at(d_w, (ks >> 2) * kWBlock + (ks & 3) * 32), idesc_n128, ks ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        mma_tf32_ts_warp(tmem_base + KK, tmem_base + kRloCol + 8 * ks,
                         desc_at(d_w, (ks >> 2) * kWBlock + (ks & 3) * 32), idesc_n64, 1u);
      mma_commit_warp(&s.d1_full);
      DBG2(it, 8);
      if (!want_dw) mma_commit_warp(&s.ab_empty);
      __syncwarp();
    } else if (warp == kGemm2Warp && want_dw) {
      mbar_wait(&s.ab_full, it & 1);
      tc_fence_after_sync();
      // GEMM 2: dW[o,i] += sum_b r[b,o] e[b,i]: both operands MN-major (rows = samples),
      // M = [r_hi | r_lo] units, N = [e_hi | e_lo] units, 8 samples per instruction
      constexpr uint32_t idesc2 = make_idesc_tf32(128, 128, 1, 1);
      const uint64_t d_r = make_desc_mn(mn_base, TM * 128, 512);
      const uint64_t d_e = make_desc_mn(mn_base + kEmn, TM * 128, 512);
#pragma unroll
      for (int ks = 0; ks < TM / 8; ++ks)
        mma_tf32_warp(tmem_base + kD2Col, desc_at(d_r, ks * 1024), desc_at(d_e, ks * 1024), idesc2,
                      (it || ks) ? 1u : 0u);
      mma_commit_warp(&s.ab_empty);
      if (it + 1 == n_tiles) mma_commit_warp(&s.d2_full);
      __syncwarp();
    }
    // ---- du epilogue: du = e * (main + correction), same element ownership
    mbar_wait(&s.d1_full, it & 1);
    tc_fence_after_sync();
    DBG2(it, 9);
    {
      float v[16], w[16];
      tmem_ld_16x256b_x4(tmem_base + lane_addr + 32 * chalf, v);
      tmem_ld_16x256b_x4(tmem_base + lane_addr + KK + 32 * chalf, w);
      tmem_ld_wait();
      DBG2(it, 10);
#pragma unroll
      for (int aa = 0; aa < 2; ++aa)
        if (du_left > 8 * aa && !(flags & 32)) {
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            float2 o;
            o.x = e[aa][n][0] * (v[4 * n + 2 * aa] + w[4 * n + 2 * aa]);
            o.y = e[aa][n][1] * (v[4 * n + 2 * aa + 1] + w[4 * n + 2 * aa + 1]);
            *reinterpret_cast<float2*>(pdu + aa * 8 * KK + 8 * n) = o;
          }
        }
      pdu += TM * KK;
      du_left -= TM;
    }
    tc_fence_before_sync();
    DBG2(it, 11);
  }
#ifdef CKB_TIMELINE
  if (dbg && (tid >> 5) == 0) g_dbg[1] = clock64();
#endif
  if (want_dw) {
    // D2 quadrants: rows 0..63 = r_hi^T [e_hi | e_lo], rows 64..127 = r_lo^T [e_hi | (dropped)].
    // dW[o][i] = D2[o][i] + D2[o][64+i] + D2[64+o][i]: the lower half goes through shared memory
    // (the operand images are dead once every MMA has completed).
    mbar_wait(&s.d2_full, 0);
    tc_fence_after_sync();
    asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory");
    const int cg = warp >> 2;
    const int row = q * 32 + lane;  // 0..127
    const int o = row & 63;
    const uint32_t xch = smem_u32(s.r_mn);  // [64 columns][64 + 1] exchange buffer
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kD2Col + cg * 16;
    float v[16];
    tmem_ld16(taddr, v);
    if (row < 64) {
      float w[16];
      tmem_ld16(taddr + KK, w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += w[j];
    } else {
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) sts32(xch + (uint32_t)((cg * 16 + j) * (KK + 1) + o) * 4u, v[j]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory");
    if (row < 64) {
      float* out = a.dWp + (((int64_t)blockIdx.x * gridDim.y + f) * KK + o) * KK + cg * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += lds32(xch + (uint32_t)((cg * 16 + j) * (KK + 1) + o) * 4u);
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
#ifdef CKB_TIMELINE
  if (dbg && (tid >> 5) == 0) g_dbg[2] = clock64();
#endif
#undef DBG2
}

void dense_tc_bwd_config(int F, int64_t B, int& splits, int& tiles_per_cta) {
  const int n_tiles = ceil_div(B, TM);
  splits = (int)max64(1, min64(n_tiles, ceil_div(2 * kNumSMs, F)));
  tiles_per_cta = ceil_div(n_tiles, splits);
  splits = ceil_div(n_tiles, tiles_per_cta);
}

bool dense_tc_ok(const DenseArgs& a) {
  return a.Ki == KK && a.Ko == KK && !a.concat && a.H >= 1 && a.H <= 2 && a.Kred == KK;
}

}  // namespace

static int g_tc_enabled = -1;
static int g_tc_flags = 3 | 512;  // bit 0: MUFU exp, bit 1: MUFU log (both set: the fast-math kernels), bit 9: tcgen05 kernels for Ki = Ko = 128
void set_tensor_cores(int on) { g_tc_enabled = on ? 1 : 0; }
void set_tc_fast_math(int bits) { g_tc_flags = bits; }
bool tc_disabled() {
  if (g_tc_enabled < 0) {
    const char* e = getenv("CKB_DISABLE_TC");
    g_tc_enabled = (e && e[0] == '1') ? 0 : 1;
  }
  return g_tc_enabled == 0;
}

int tc_flags() { return g_tc_flags; }

int tucker_debug_read(void* dst, size_t bytes);

int bwd3_debug_read(void* dst, size_t bytes);

int debug_read(void* dst, size_t bytes) {
  if (g_tc_flags & 4096) return bwd3_debug_read(dst, bytes);
  if (g_tc_flags & 256) return tucker_debug_read(dst, bytes);
  if (bytes > sizeof(long long) * 512) bytes = sizeof(long long) * 512;
  CKB_CUDA_CHECK(cudaMemcpyFromSymbol(dst, g_dbg, bytes));
  return CKB_OK;
}

int dense_tc_fwd(const DenseArgs& a, int F, Ctx& c) {
  if (tc_disabled() || !dense_tc_ok(a)) return 1;
  const size_t smem = sizeof(FwdSmem) + 1024;
  static PerDeviceOnce attr;
  if (attr.first()) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_tc_fwd_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_tc_fwd_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int n_tiles = ceil_div(a.B, TM);
  // ~4 CTAs per SM over the launch (2 resident), several tiles per CTA when the batch allows
  int splits = (int)max64(1, min64(n_tiles, ceil_div(4 * kNumSMs, F)));
  const int tiles_per_cta = ceil_div(n_tiles, splits);
  splits = ceil_div(n_tiles, tiles_per_cta);
  dim3 grid(splits, F);
  if ((g_tc_flags & 3) == 3)
    CKB_CUDA_CHECK(launch_pdl(dense_tc_fwd_kernel<true>, grid, dim3(kThreads), smem, c.stream, a, tiles_per_cta));
  else
    CKB_CUDA_CHECK(launch_pdl(dense_tc_fwd_kernel<false>, grid, dim3(kThreads), smem, c.stream, a, tiles_per_cta));
  c.launches++;
  return CKB_OK;
}

size_t dense_tc_bwd_ws(int F, int H, int Ko, int Kred, int64_t B) {
  if (Ko != KK || Kred != KK || H > 2) return 0;
  int splits, tpc;
  dense_tc_bwd_config(F, B, splits, tpc);
  const size_t a = splits > 1 ? (size_t)splits * F * KK * KK * 4 : 0;
  const size_t b = dense_tc_bwd3_ws(F, B);
  return a > b ? a : b;
}

int dense_tc_bwd(const DenseArgs& a_in, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes) {
  if (tc_disabled() || !dense_tc_ok(a_in) || a_in.max_cons > kMaxCons) return 1;
  if (dense_tc_bwd3_ok(a_in, a_in.rows64)) return dense_tc_bwd3(a_in, F, dW, c, ws, ws_bytes);
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
  const size_t smem2 = sizeof(Bwd2Smem) + 1024;
  static PerDeviceOnce attr;
  if (attr.first()) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_tc_bwd_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_tc_bwd_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_tc_bwd2_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_tc_bwd2_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  }
  dim3 grid(splits, F);
  if (!(g_tc_flags & 1024)) {  // bit 10: the round-1 kernel (A/B runs)
    if ((g_tc_flags & 3) == 3)
      dense_tc_bwd2_kernel<true><<<grid, kBwdThreads, smem2, c.stream>>>(a, tpc, dW ? 1 : 0, g_tc_flags);
    else
      dense_tc_bwd2_kernel<false><<<grid, kBwdThreads, smem2, c.stream>>>(a, tpc, dW ? 1 : 0, g_tc_flags);
  } else if ((g_tc_flags & 3) == 3)
    dense_tc_bwd_kernel<true><<<grid, kBwdThreads, smem, c.stream>>>(a, tpc, dW ? 1 : 0, g_tc_flags);
  else
    dense_tc_bwd_kernel<false><<<grid, kBwdThreads, smem, c.stream>>>(a, tpc, dW ? 1 : 0, g_tc_flags);
  CKB_LAUNCH_CHECK();
  c.launches++;
  if (dW && splits > 1) return reduce_partials(a.dWp, dW, (int64_t)n, splits, c);
  return CKB_OK;
}

}  // namespace ckb
