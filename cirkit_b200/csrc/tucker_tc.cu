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
// operand e1 (x) e2 is formed in registers and written straight into TMEM (the A operand of the
// MMAs), so only x1, x2, y, g, W and the per-step operand images derived from them touch HBM.
// All products are 3xTF32 (hi*hi + hi*lo + lo*hi, see sm100.cuh); shared-memory operand tiles
// (the weight side) are 128-byte-swizzled K-major.
//
// The tensor core truncates when it folds a product group into the fp32 accumulator (~0.6 ulp low
// per accumulating instruction, measured).  Over the 512 k-steps of a Ki^2 = 4096 reduction that
// would bias log S by ~1e-5 per layer, all with the same sign, so the forward accumulates in TMEM
// only over chunks of 16 k-steps and folds the chunks into fp32 registers with round-to-nearest
// adds; the two small correction products have their own accumulator.
#include "dense.cuh"
#include "sm100.cuh"
#include "tc_util.cuh"

namespace ckb {
using namespace sm100;

// Debug timeline (built with -DCKB_TIMELINE): CTA (0,0) of the forward kernel records clock64()
// at pipeline events of its first k-blocks; read back with ckb_debug_read when bit 8 of
// CKB_OPT_TC_FAST_MATH is set.
__device__ long long g_dbg_tk[1024];
int tucker_debug_read(void* dst, size_t bytes) {
  if (bytes > sizeof(long long) * 1024) bytes = sizeof(long long) * 1024;
  CKB_CUDA_CHECK(cudaMemcpyFromSymbol(dst, g_dbg_tk, bytes));
  return CKB_OK;
}
#ifdef CKB_TIMELINE
#define TKDBG(kb, slot)                                                                    \
  do {                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (kb) < 24) g_dbg_tk[16 + (kb) * 16 + (slot)] = clock64(); \
  } while (0)
#define TKMARK(slot)                                                                       \
  do {                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 0) g_dbg_tk[slot] = clock64();                    \
  } while (0)
#define DXDBG(i, slot)                                                                     \
  do {                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (i) >= 8 && (i) < 32) g_dbg_tk[512 + ((i) - 8) * 16 + (slot)] = clock64(); \
  } while (0)
#else
#define TKDBG(kb, slot) do { } while (0)
#define TKMARK(slot) do { } while (0)
#define DXDBG(i, slot) do { } while (0)
#endif

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

// TMEM budget and shared-memory traffic shape all three kernels: a tf32 tcgen05.mma reads 32 bytes
// of every operand row per instruction, so with both operands in shared memory (v1 of this file)
// the 128 B/clk of the shared-memory crossbar, not the tensor pipe, set the pace (27 % pipe
// utilisation measured).  The operand that is generated on the fly (the Kronecker rows, r, P^T)
// is therefore written to TMEM with tcgen05.st and consumed as the A operand from there; only the
// weight-side tiles travel through shared memory.
//
// ==========================================================================================
// Forward.  CTA = (fold, 128 samples).  The reduction runs over 64 ring slots, one per index i of
// the first input (64 reduction indices (i, j = 0..63) = 8 k-steps per slot).
//   warps 0-7  : A producers; thread = (sample row, column half): e1[b,i] * e2[b,j] for its 32 j's
//                -> (hi, lo) -> TMEM slot (tcgen05.st)
//   warp 8     : weight loader: one 32 KB bulk copy (TMA engine) per slot of the pre-split image
//                tucker_split_w_kernel wrote (splitting W inside every CTA, 16 per fold, made the
//                weight producers the critical path: the generic->async proxy fence after their
//                shared-memory stores also waits for the global loads they have in flight)
//   warps 10-11: MMA issue.  Measured on B200: ~150 clk per mbarrier wait, ~80 per commit and
//                ~50 per tcgen05.mma when one lane issues from a divergent branch, so the whole
//                warp stays converged and the instructions are predicated on the elected lane
//                (~25-35 clk per MMA), A and W share ONE full/empty barrier pair per slot, a
//                k-step is two instructions: a_hi x [W_hi | W_lo] (N = 128: main | correction)
//                and a_lo x W_hi (N = 64: correction), and the two warps take alternate chunks
//   warps 12-15: chunk accumulation (TMEM -> fp32 registers every 16 k-steps) and the log epilogue
// TMEM columns: (main | correction) accumulator pair x2 [0,256), A ring 2 x (hi 64 | lo 64) [256,512).
// ==========================================================================================
constexpr int kFwdThreads = 512;
constexpr int kNS = 2;            // ring depth
constexpr int kChunkSlots = 2;    // slots (of 8 k-steps) per TMEM chunk
constexpr uint32_t kColA = 256;

struct __align__(1024) TkFwdSmem {
  float w[kNS][2][128 * 32];  // [slot][k-block][hi o 0..63 | lo o 0..63][32] swizzled   64 KB
  float msum[TM];
  uint64_t full[kNS], empty[kNS], mfull[2], mempty[2], turn[2];
  uint32_t tmem_base;
};

template <bool FAST>
__global__ void __launch_bounds__(kFwdThreads, 1)
tucker_tc_fwd_kernel(DenseArgs a, const float* __restrict__ wimg) {
  extern __shared__ uint8_t smem_raw[];
  TkFwdSmem& s = *reinterpret_cast<TkFwdSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int64_t b0 = (int64_t)blockIdx.x * TM;
  if (tid == 0) TKMARK(0);

  if (tid == 0) {
    for (int i = 0; i < kNS; ++i) {
      mbar_init(&s.full[i], 8 + 1);  // 8 A-producer warps + the weight copy (arrive.expect_tx)
      mbar_init(&s.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.mfull[i], 1);
      mbar_init(&s.mempty[i], 128);
      mbar_init(&s.turn[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;
  if (tid == 0) TKMARK(1);

  if (warp < 8) {
    // ================= A producers =================
    const int q = warp & 3, h = warp >> 2;
    const int row = q * 32 + lane;
    const int64_t b = b0 + row;
    const bool valid = b < a.B;
    const float* x1 = in_row(a, f, 0) + (valid ? b : 0) * KK;
    const float* x2 = in_row(a, f, 1) + (valid ? b : 0) * KK;
    float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      m1 = fmaxf(m1, max4(ldg_nc(x1 + 4 * c)));
      m2 = fmaxf(m2, max4(ldg_nc(x2 + 4 * c)));
    }
    m1 = clamp_max(m1);
    m2 = clamp_max(m2);
    float e2[32];  // [jh][16]: columns jh*32 + h*16 + c
#pragma unroll
    for (int jh = 0; jh < 2; ++jh)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 v = ldg_nc(x2 + jh * 32 + h * 16 + 4 * c);
        e2[jh * 16 + 4 * c] = valid ? exp_nonpos<FAST>(v.x - m2) : 0.f;
        e2[jh * 16 + 4 * c + 1] = valid ? exp_nonpos<FAST>(v.y - m2) : 0.f;
        e2[jh * 16 + 4 * c + 2] = valid ? exp_nonpos<FAST>(v.z - m2) : 0.f;
        e2[jh * 16 + 4 * c + 3] = valid ? exp_nonpos<FAST>(v.w - m2) : 0.f;
      }
    if (h == 0) s.msum[row] = fmaxf(m1 + m2, -FLT_MAX);
    if (tid == 0) TKMARK(2);
    const uint32_t abase = tmem_base + ((uint32_t)(q * 32) << 16) + kColA + h * 16;
    float x1n = __ldg(x1);
    for (int i = 0; i < KK; ++i) {
      const float a1 = exp_nonpos<FAST>(x1n - m1);
      if (i + 1 < KK) x1n = __ldg(x1 + i + 1);
      const uint32_t sl = (uint32_t)i & (kNS - 1);
      // the products are ready before the slot is: only the TMEM stores sit behind the wait
      float hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) split_tf32(a1 * e2[c], hi[c], lo[c]);
      if (tid == 0) TKDBG(i, 3);
      mbar_wait(&s.empty[sl], ((i / kNS) & 1) ^ 1);
      tc_fence_after_sync();
      if (tid == 0) TKDBG(i, 4);
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        tmem_st16(abase + sl * 128 + jh * 32, hi + jh * 16);
        tmem_st16(abase + sl * 128 + 64 + jh * 32, lo + jh * 16);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.full[sl]);
      if (tid == 0) TKDBG(i, 5);
    }
  } else if (warp == 8) {
    // ================= weight loader: one bulk copy (TMA engine) per slot =================
    // wimg holds, per fold and slot i, the two stacked [W_hi | W_lo] swizzled k-block tiles exactly
    // as the MMA reads them (tucker_split_w_kernel), so a slot is 32 contiguous KB
    if (lane == 0) {
      const float* src = wimg + (int64_t)f * KK * (2 * 128 * 32);
      for (int i = 0; i < KK; ++i) {
        const uint32_t sl = (uint32_t)i & (kNS - 1);
        mbar_wait(&s.empty[sl], ((i / kNS) & 1) ^ 1);
        mbar_arrive_expect_tx(&s.full[sl], 2 * kTile);
        bulk_g2s(&s.w[sl][0][0], src + (int64_t)i * (2 * 128 * 32), 2 * kTile, &s.full[sl]);
      }
    }
  } else if (warp == 10 || warp == 11) {
    // ================= MMA issuers (whole warp converged, instructions on the elected lane) =====
    // Two issuer warps take alternate chunks (each chunk has its own accumulator pair), so the
    // ~150-clock barrier waits and ~80-clock commits of one overlap the MMA issue of the other.
    constexpr uint32_t idesc_n128 = make_idesc_tf32(TM, 2 * KK, 0, 0);
    constexpr uint32_t idesc_n64 = make_idesc_tf32(TM, KK, 0, 0);
    const uint64_t d_w = make_desc(smem_u32(s.w), 16, 1024);
    const int buf = warp - 10;
    const uint32_t d = tmem_base + buf * 128;
    for (int chunk = buf; chunk < KK / kChunkSlots; chunk += 2) {
      // A parity wait can only tell two consecutive phases of a barrier apart, and the slots'
      // previous phases were consumed by the other issuer: wait until it has seen all of them.
      const int n = chunk >> 1;
      if (chunk > 0) mbar_wait(&s.turn[buf], buf ? (n & 1) : ((n - 1) & 1));
      mbar_wait(&s.mempty[buf], (n & 1) ^ 1);
#pragma unroll
      for (int u = 0; u < kChunkSlots; ++u) {
        const int i = chunk * kChunkSlots + u;
        const uint32_t sl = (uint32_t)i & (kNS - 1);
        if (lane == 0) TKDBG(i, 0);
        mbar_wait(&s.full[sl], (i / kNS) & 1);
        // hand the ring over: the other issuer may start waiting for the next chunk's slots
        if (u == kChunkSlots - 1 && lane == 0) mbar_arrive(&s.turn[buf ^ 1]);
        tc_fence_after_sync();
        if (lane == 0) TKDBG(i, 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t a_hi = tmem_base + kColA + sl * 128 + ks * 8, a_lo = a_hi + 64;
          const uint64_t b_w = desc_at(d_w, (sl * 2 + (ks >> 2)) * kTile + (ks & 3) * 32);
          mma_tf32_ts_warp(d, a_hi, b_w, idesc_n128, (u == 0 && ks == 0) ? 0u : 1u);
          mma_tf32_ts_warp(d + KK, a_lo, b_w, idesc_n64, 1u);
        }
        if (lane == 0) TKDBG(i, 10);
        mma_commit_warp(&s.empty[sl]);
        if (u == kChunkSlots - 1) mma_commit_warp(&s.mfull[buf]);
        if (lane == 0) TKDBG(i, 2);
      }
    }
  } else if (warp >= 12) {
    // ================= chunk accumulation + epilogue =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float acc[KK];
#pragma unroll
    for (int j = 0; j < KK; ++j) acc[j] = 0.f;
    constexpr int kChunks = KK / kChunkSlots;
    for (int c = 0; c < kChunks; ++c) {
      const int buf = c & 1;
      mbar_wait_relaxed(&s.mfull[buf], (c >> 1) & 1);
      tc_fence_after_sync();
      if (tid == 12 * 32) TKDBG(c * kChunkSlots + kChunkSlots - 1, 7);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float v[16], w[16];
        tmem_ld16(lane_base + buf * 128 + cc * 16, v);
        tmem_ld16(lane_base + buf * 128 + KK + cc * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[cc * 16 + j] += v[j] + w[j];
      }
      tc_fence_before_sync();
      mbar_arrive(&s.mempty[buf]);
    }
    if (tid == 12 * 32) TKMARK(3);
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
  if (tid == 0) TKMARK(4);
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
  if (tid == 0) TKMARK(5);
}

// ==========================================================================================
// Backward, part 0: e1 = exp(x1 - m1), e2 = exp(x2 - m2) and r = g / S = g exp(m1 + m2 - y) for
// 32 samples of a fold per CTA, written as the scratch block parts 1 and 2 read: stacked
// [r_hi | r_lo] rows, e1 rows, e2 rows, each [unit][32 samples], 16-byte chunks xor-swizzled.
// HBM-bound streaming kernel (coalesced 16-byte loads, half-warp row maxima, transpose through
// shared memory).
// ==========================================================================================
constexpr int kBlkFloats = 128 * 32 + 64 * 32 + 64 * 32;  // 8192 floats = 32 KB
constexpr int kBlkE1 = 128 * 32, kBlkE2 = 128 * 32 + 64 * 32;
__device__ __forceinline__ int blk_off(int unit, int bcol) {
  return unit * 32 + ((((bcol >> 2) ^ unit) & 7) << 2) + (bcol & 3);
}

template <bool FAST>
__global__ void __launch_bounds__(256) tucker_prep_kernel(DenseArgs a, float* scratch, int nblk) {
  __shared__ float T[4][KK][33];  // r_hi, r_lo, e1, e2 as [unit][sample]
  const int f = blockIdx.y, tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * 32;
  const float* x1 = in_row(a, f, 0);
  const float* x2 = in_row(a, f, 1);
  const float* yf = a.y + (int64_t)f * a.B * KK;
  const int cbeg = a.gs.cons_ptr[f], cend = a.gs.cons_ptr[f + 1];
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const int idx4 = tid + n * 256;
    const int row = idx4 >> 4, c4 = (idx4 & 15) * 4;  // 16 lanes per sample row
    const int64_t b = b0 + row;
    const bool valid = b < a.B;
    const int64_t e = (valid ? b : 0) * KK + c4;
    const float4 v1 = ldg_stream(x1 + e), v2 = ldg_stream(x2 + e), yv = ldg_stream(yf + e);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = cbeg; k < cend; ++k) {
      const float4 z = ldg_stream(a.gs.garena + a.gs.B * a.gs.cons_rows[k] + e);
      g.x += z.x; g.y += z.y; g.z += z.z; g.w += z.w;
    }
    const float m1 = clamp_max(half_warp_max(max4(v1)));
    const float m2 = clamp_max(half_warp_max(max4(v2)));
    const float ms = fmaxf(m1 + m2, -FLT_MAX);
    const float z = valid ? 1.f : 0.f;
    const float e1[4] = {z * exp_nonpos<FAST>(v1.x - m1), z * exp_nonpos<FAST>(v1.y - m1),
                         z * exp_nonpos<FAST>(v1.z - m1), z * exp_nonpos<FAST>(v1.w - m1)};
    const float e2[4] = {z * exp_nonpos<FAST>(v2.x - m2), z * exp_nonpos<FAST>(v2.y - m2),
                         z * exp_nonpos<FAST>(v2.z - m2), z * exp_nonpos<FAST>(v2.w - m2)};
    const float r[4] = {z * g.x * exp_capped<FAST>(ms - yv.x), z * g.y * exp_capped<FAST>(ms - yv.y),
                        z * g.z * exp_capped<FAST>(ms - yv.z), z * g.w * exp_capped<FAST>(ms - yv.w)};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float hi, lo;
      split_tf32(r[t], hi, lo);
      T[0][c4 + t][row] = hi;
      T[1][c4 + t][row] = lo;
      T[2][c4 + t][row] = e1[t];
      T[3][c4 + t][row] = e2[t];
    }
  }
  __syncthreads();
  float* blk = scratch + ((int64_t)f * nblk + blockIdx.x) * kBlkFloats;
  const int warp = tid >> 5, lane = tid & 31;
  for (int u = warp; u < 4 * KK; u += 8) {  // 256 rows of 32 samples: r_hi 64, r_lo 64, e1 64, e2 64
    const int part = u >> 6, unit = u & 63;
    // scratch row index inside its section: r rows are stacked (hi 0..63, lo 64..127)
    const int sec = part < 2 ? 0 : (part == 2 ? kBlkE1 : kBlkE2);
    const int urow = part == 1 ? 64 + unit : unit;
    blk[sec + blk_off(urow, lane)] = T[part][unit][lane];
  }
}

// ==========================================================================================
// Backward, part 1 (d/dx1, d/dx2 and the operands of part 2).  CTA = (fold, 256 samples).
// 16 worker warps; thread = (sample row, column half).  Prologue: row maxes, e2, r = g/S, written
// as (hi, lo) to TMEM (the A operand of every GEMM of this CTA); the transposed, split r and the
// raw e1, e2 also go to a scratch block per 32 samples, already in the swizzled image part 2
// multiplies from.  Main loop over i: the 64x64 slice W[:, i, :] is staged transposed ([j][o], so
// that o is the K axis), GEMM T_i[b, j] = sum_o r[b,o] W[o,i,j] (K = 64; the three 3xTF32
// products share one accumulator: a 2^-22 relative bias is irrelevant for gradients), and the
// workers fold T_i into d/dx1[b,i] (dot with e2) and d/dx2[b,:] (axpy with e1[b,i]).
// TMEM columns: T accumulators [tile][buffer] x 64 -> [0,256); r (hi | lo) per tile -> [256,512).
// ==========================================================================================
constexpr int kDxWorkers = 16;
constexpr int kDxNW = 4;  // weight ring depth (32 KB slots)
constexpr int kDxThreads = (kDxWorkers + 3) * 32;  // 608: warps 16 / 17 issue the MMAs of M tile 0 / 1,
                                                   // warp 18 runs the weight copies

// scratch block of 32 samples: see tucker_prep_kernel

struct __align__(1024) TkDxSmem {
  float w[kDxNW][2][128 * 32];  // [slot][o half][hi j 0..63 | lo j 0..63][32 o]   128 KB
  float stage[KK * ROWS];       // e1 as [i][row]                                         64 KB
  float part[2][ROWS];
  uint64_t wfull[kDxNW], wempty[kDxNW], tfull[2][2], tempty[2][2];  // t*: [M tile][buffer]
  uint32_t tmem_base;
};

template <bool FAST>
__global__ void __launch_bounds__(kDxThreads, 1)
tucker_tc_bwd_dx_kernel(DenseArgs a, float* scratch, int nblk, const float* __restrict__ wimg) {
  extern __shared__ uint8_t smem_raw[];
  TkDxSmem& s = *reinterpret_cast<TkDxSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int64_t b0 = (int64_t)blockIdx.x * ROWS;
  if (tid == 0) TKMARK(8);

  if (tid == 0) {
    for (int i = 0; i < kDxNW; ++i) {
      mbar_init(&s.wfull[i], 1);
      mbar_init(&s.wempty[i], 2);  // one commit per tile issuer
    }
    for (int i = 0; i < 2; ++i)
      for (int t = 0; t < 2; ++t) {
        mbar_init(&s.tfull[t][i], 1);
        mbar_init(&s.tempty[t][i], (kDxWorkers / 2) * 32);
      }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;
  if (tid == 0) TKMARK(9);

  if (warp < kDxWorkers) {
    const int q = warp & 3, t = (warp >> 2) & 1, ch = warp >> 3;
    const int p = t * TM + q * 32 + lane;  // row inside the CTA
    const int64_t b = b0 + p;
    const bool valid = b < a.B;
    const int64_t bsafe = valid ? b : 0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // r (split), e1 and e2 of this CTA's samples were written by tucker_prep_kernel, transposed
    // ([unit][32 samples] rows), so lanes = consecutive samples read them coalesced.  (A prologue
    // that derived them here from x1, x2, y, g took 30 us of a 100 us CTA, serial on every SM.)
    const float* blk = scratch + ((int64_t)f * nblk + (b0 + p) / 32) * kBlkFloats;
    float* stg = s.stage;  // e1 as [i][row]
    float e2[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) e2[c] = __ldg(blk + kBlkE2 + blk_off(32 * ch + c, lane));
    {
      const uint32_t rbase = lane_base + 256 + t * 128 + ch * 32;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int o = 32 * ch + hh * 16 + c;
          hi[c] = __ldg(blk + blk_off(o, lane));
          lo[c] = __ldg(blk + blk_off(64 + o, lane));
        }
        tmem_st16(rbase + hh * 16, hi);
        tmem_st16(rbase + 64 + hh * 16, lo);
      }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) stg[(32 * ch + c) * ROWS + p] = __ldg(blk + kBlkE1 + blk_off(32 * ch + c, lane));
    tmem_st_wait();
    tc_fence_before_sync();

    if (tid == 0) TKMARK(10);
    // every worker's r is in TMEM before the first MMA
    asm volatile("bar.sync 9, %0;" ::"n"(kDxThreads) : "memory");
    if (tid == 0) TKMARK(11);

    float acc2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc2[j] = 0.f;
    float* gin1 = a.gin + (((int64_t)f * 2 + 0) * a.B + bsafe) * KK;
    for (int i = 0; i < KK; ++i) {
      const float e1i = stg[i * ROWS + p];
      const int buf = i & 1;
      if (tid == 0) DXDBG(i, 5);
      if (tid == 8 * 32) DXDBG(i, 9);
      mbar_wait(&s.tfull[t][buf], (i >> 1) & 1);
      tc_fence_after_sync();
      if (tid == 0) DXDBG(i, 6);
      if (tid == 8 * 32) DXDBG(i, 10);
      const uint32_t taddr = lane_base + t * 128 + buf * 64 + ch * 32;
      float dot = 0.f;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float v[8];
        tmem_ld8(taddr + h * 8, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dot = fmaf(v[j], e2[h * 8 + j], dot);
          acc2[h * 8 + j] = fmaf(v[j], e1i, acc2[h * 8 + j]);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&s.tempty[t][buf]);
      if (tid == 0) DXDBG(i, 7);
      if (tid == 8 * 32) DXDBG(i, 11);
      if (ch == 1) s.part[buf][p] = dot;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 7)) : "memory");
      if (ch == 0 && valid) gin1[i] = e1i * (dot + s.part[buf][p]);
      if (tid == 0) DXDBG(i, 8);
      if (tid == 8 * 32) DXDBG(i, 12);
    }
    if (tid == 0) TKMARK(12);
    if (valid) {
      float* gin2 = a.gin + (((int64_t)f * 2 + 1) * a.B + b) * KK + 32 * ch;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(gin2 + 4 * c) =
            make_float4(e2[4 * c] * acc2[4 * c], e2[4 * c + 1] * acc2[4 * c + 1],
                        e2[4 * c + 2] * acc2[4 * c + 2], e2[4 * c + 3] * acc2[4 * c + 3]);
    }
  } else if (warp == kDxWorkers + 2) {
    // ================= weight loader =================
    // the transposed, split slice W[:, i, :]^T ([hi j | lo j] rows x o) comes as ONE 32 KB bulk
    // copy per i from the image tucker_split_wt_kernel wrote
    asm volatile("bar.sync 9, %0;" ::"n"(kDxThreads) : "memory");
    if (lane == 0) {
      const float* wsrc = wimg + (int64_t)f * KK * (2 * 128 * 32);
      for (int i = 0; i < KK; ++i) {
        const uint32_t st = (uint32_t)i & (kDxNW - 1);
        mbar_wait(&s.wempty[st], ((i / kDxNW) & 1) ^ 1);
        mbar_arrive_expect_tx(&s.wfull[st], 2 * kTile);
        bulk_g2s(&s.w[st][0][0], wsrc + (int64_t)i * (2 * 128 * 32), 2 * kTile, &s.wfull[st]);
      }
    }
  } else {
    // ================= MMA issuers: warp 16 -> M tile 0, warp 17 -> M tile 1 =================
    // whole warp converged, instructions on the elected lane (see the forward kernel)
    const int t = warp - kDxWorkers;
    constexpr uint32_t idesc_n64 = make_idesc_tf32(TM, KK, 0, 0);
    const uint64_t d_w = make_desc(smem_u32(s.w), 16, 1024);
    asm volatile("bar.sync 9, %0;" ::"n"(kDxThreads) : "memory");  // r is in TMEM
    tc_fence_after_sync();
    for (int i = 0; i < KK; ++i) {
      const uint32_t st = i & (kDxNW - 1), buf = i & 1;
      if (lane == 0 && t == 0) DXDBG(i, 0);
      mbar_wait(&s.tempty[t][buf], ((i >> 1) & 1) ^ 1);
      if (lane == 0 && t == 0) DXDBG(i, 1);
      mbar_wait(&s.wfull[st], (i / kDxNW) & 1);
      tc_fence_after_sync();
      if (lane == 0 && t == 0) DXDBG(i, 2);
      const uint32_t d = tmem_base + t * 128 + buf * 64;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t a_hi = tmem_base + 256 + t * 128 + ks * 8, a_lo = a_hi + 64;
        const uint64_t b_hi = desc_at(d_w, (st * 2 + (ks >> 2)) * kTile + (ks & 3) * 32);
        const uint64_t b_lo = desc_at(d_w, (st * 2 + (ks >> 2)) * kTile + 64 * 128 + (ks & 3) * 32);
        mma_tf32_ts_warp(d, a_hi, b_hi, idesc_n64, ks ? 1u : 0u);
        mma_tf32_ts_warp(d, a_hi, b_lo, idesc_n64, 1u);
        mma_tf32_ts_warp(d, a_lo, b_hi, idesc_n64, 1u);
      }
      if (lane == 0 && t == 0) DXDBG(i, 3);
      mma_commit_warp(&s.wempty[st]);
      mma_commit_warp(&s.tfull[t][buf]);
      if (lane == 0 && t == 0) DXDBG(i, 4);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) TKMARK(13);
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================================
// Backward, part 2 (d/dW).  CTA = (fold, 256 consecutive reduction indices n = (i, j): four i's),
// loops over ALL samples in blocks of 32 (the K axis of this GEMM):
//   dW^T[n, o] = sum_b P[b,n] r[b,o],  P[b,(i,j)] = e1[b,i] e2[b,j]
// A = P^T (two M tiles of 128 n; one producer thread per n forms the 32 products of a block from
// the raw e1 / e2 rows and writes (hi, lo) to a TMEM ring), B = stacked [r_hi^T ; r_lo^T]
// (128 x 32 per block, bulk-copied as part 1 wrote it).  One accumulator per tile takes all three
// 3xTF32 products.  TMEM columns: dW^T accumulators [0,128), A ring [stage][tile] x 64 [128,512).
// ==========================================================================================
constexpr int kDwProdWarps = 8;
constexpr int kDwTmaWarp = 8, kDwMmaWarp = 9;  // warps 9 and 10 issue the MMAs of M tile 0 / 1
constexpr int kDwThreads = 11 * 32;
constexpr int NCH = 256;  // reduction indices per CTA
constexpr int kNR = 4;    // raw ring (shared memory)
constexpr int kNP = 3;    // P ring (TMEM)

struct __align__(1024) TkDwSmem {
  float rstack[kNR][128 * 32];  // 64 KB
  float e2raw[kNR][64 * 32];    // 32 KB
  float e1raw[kNR][4 * 32];     //  2 KB
  uint64_t raw_full[kNR], raw_empty[kNR], p_full[kNP], p_empty[kNP], done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kDwThreads, 1)
tucker_tc_bwd_dw_kernel(const float* __restrict__ scratch, int nblk_alloc, int nblk, float* dW) {
  extern __shared__ uint8_t smem_raw[];
  TkDwSmem& s = *reinterpret_cast<TkDwSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int i0 = 4 * blockIdx.x, n0 = NCH * blockIdx.x;

  if (tid == 0) {
    for (int i = 0; i < kNR; ++i) {
      mbar_init(&s.raw_full[i], 1);
      mbar_init(&s.raw_empty[i], 2);  // one commit per tile issuer
    }
    for (int i = 0; i < kNP; ++i) {
      mbar_init(&s.p_full[i], kDwProdWarps);
      mbar_init(&s.p_empty[i], 2);
    }
    mbar_init(&s.done, 2);
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
        const int sr = kb & (kNR - 1);
        mbar_wait(&s.raw_empty[sr], ((kb / kNR) & 1) ^ 1);
        const float* blk = fblk + (int64_t)kb * kBlkFloats;
        mbar_arrive_expect_tx(&s.raw_full[sr], 128 * 128 + 64 * 128 + 4 * 128);
        bulk_g2s(s.rstack[sr], blk, 128 * 128, &s.raw_full[sr]);
        bulk_g2s(s.e2raw[sr], blk + kBlkE2, 64 * 128, &s.raw_full[sr]);
        bulk_g2s(s.e1raw[sr], blk + kBlkE1 + i0 * 32, 4 * 128, &s.raw_full[sr]);
      }
    }
  } else if (warp >= kDwMmaWarp) {
    // whole warp converged, instructions on the elected lane (see the forward kernel); one issuer
    // warp per M tile so that their barrier waits and commits overlap
    constexpr uint32_t idesc_n64 = make_idesc_tf32(TM, KK, 0, 0);
    const uint64_t d_r = make_desc(smem_u32(s.rstack), 16, 1024);
    const int tt = warp - kDwMmaWarp;
    const uint32_t d = tmem_base + tt * 64;
    int sp = 0, php = 0;
    for (int kb = 0; kb < nblk; ++kb) {
      const uint32_t sr = kb & (kNR - 1);
      mbar_wait(&s.raw_full[sr], (kb / kNR) & 1);
      mbar_wait(&s.p_full[sp], php);
      tc_fence_after_sync();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t a_hi = tmem_base + 128 + (sp * 2 + tt) * 64 + ks * 8, a_lo = a_hi + 32;
        const uint64_t b_hi = desc_at(d_r, sr * kTile + ks * 32);
        const uint64_t b_lo = desc_at(d_r, sr * kTile + 64 * 128 + ks * 32);
        mma_tf32_ts_warp(d, a_hi, b_hi, idesc_n64, (kb | ks) ? 1u : 0u);
        mma_tf32_ts_warp(d, a_hi, b_lo, idesc_n64, 1u);
        mma_tf32_ts_warp(d, a_lo, b_hi, idesc_n64, 1u);
      }
      mma_commit_warp(&s.raw_empty[sr]);
      mma_commit_warp(&s.p_empty[sp]);
      if (++sp == kNP) { sp = 0; php ^= 1; }
    }
    mma_commit_warp(&s.done);
  } else {
    // ================= producers: P^T rows into TMEM =================
    const int q = warp & 3, t = warp >> 2;
    const int n = tid;  // row of the P^T operand: (i0 + il, j); n = t*128 + q*32 + lane
    const int il = n >> 6, j = n & 63;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t e1row = (uint32_t)il * 128u, e1key = (uint32_t)(i0 + il) & 7u;
    const uint32_t e2row = (uint32_t)j * 128u, e2key = (uint32_t)j & 7u;
    const uint32_t e1s = smem_u32(s.e1raw), e2s = smem_u32(s.e2raw);
    int sp = 0, php = 0;
    for (int kb = 0; kb < nblk; ++kb) {
      const uint32_t sr = kb & (kNR - 1);
      mbar_wait(&s.raw_full[sr], (kb / kNR) & 1);
      mbar_wait(&s.p_empty[sp], php ^ 1);
      tc_fence_after_sync();
      const uint32_t abase = lane_base + 128 + (sp * 2 + t) * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t cc = (uint32_t)(hh * 4 + c);
          const float4 u = lds128(e1s + sr * (4 * 128) + e1row + (((cc ^ e1key) & 7u) << 4));
          const float4 v = lds128(e2s + sr * (64 * 128) + e2row + (((cc ^ e2key) & 7u) << 4));
          split_tf32(u.x * v.x, hi[4 * c], lo[4 * c]);
          split_tf32(u.y * v.y, hi[4 * c + 1], lo[4 * c + 1]);
          split_tf32(u.z * v.z, hi[4 * c + 2], lo[4 * c + 2]);
          split_tf32(u.w * v.w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        tmem_st16(abase + hh * 16, hi);
        tmem_st16(abase + 32 + hh * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.p_full[sp]);
      if (++sp == kNP) { sp = 0; php ^= 1; }
    }
    // ================= epilogue: dW[o][n0 + n] = D[n][o] =================
    mbar_wait_relaxed(&s.done, 0);
    tc_fence_after_sync();
    float* out = dW + (int64_t)f * KK * KRED + n0 + n;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      float v[16];
      tmem_ld16(lane_base + t * 64 + cc * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) out[(int64_t)(cc * 16 + jj) * KRED] = v[jj];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kDwMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// W (F, 64, 4096) -> per (fold, slot i, k-block): stacked [W_hi rows o | W_lo rows o][32 j] tiles,
// 128-byte swizzled: the shared-memory image of the forward's B operand.
__global__ void tucker_split_w_kernel(const float* __restrict__ W, float* __restrict__ img, int64_t n16) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n16;
       idx += (int64_t)gridDim.x * blockDim.x) {
    // idx enumerates 16-byte pieces of W in memory order: (f, o, i, kbi, c)
    const int c = (int)(idx & 7), kbi = (int)((idx >> 3) & 1), i = (int)((idx >> 4) & 63);
    const int o = (int)((idx >> 10) & 63);
    const int64_t f = idx >> 16;
    const float4 v = ldg_stream(W + idx * 4);
    float4 hi, lo;
    split4(v, hi, lo);
    float* dst = img + ((f * KK + i) * 2 + kbi) * (128 * 32) + o * 32 + (((c ^ o) & 7) << 2);
    *reinterpret_cast<float4*>(dst) = hi;
    *reinterpret_cast<float4*>(dst + 64 * 32) = lo;
  }
}

// W (F, 64, 4096) -> per (fold, i, o half): stacked [W_hi[o, i, j] rows j | W_lo rows j][32 o] tiles,
// 128-byte swizzled: the shared-memory image of the backward's B operand (o is the K axis there).
__global__ void __launch_bounds__(256) tucker_split_wt_kernel(const float* __restrict__ W,
                                                             float* __restrict__ img) {
  __shared__ float T[KK][KK + 1];
  const int i = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  const float* src = W + (int64_t)f * KK * KRED + i * KK;
  for (int idx = tid; idx < KK * KK / 4; idx += 256) {
    const int o = idx >> 4, c = idx & 15;
    const float4 v = ldg_stream(src + (int64_t)o * KRED + 4 * c);
    T[o][4 * c] = v.x; T[o][4 * c + 1] = v.y; T[o][4 * c + 2] = v.z; T[o][4 * c + 3] = v.w;
  }
  __syncthreads();
  float* dst = img + ((int64_t)f * KK + i) * (2 * 128 * 32);
  for (int idx = tid; idx < 2 * KK * 8; idx += 256) {
    const int kb = idx >> 9, j = (idx >> 3) & 63, c = idx & 7;
    const int o = kb * 32 + c * 4;
    float4 v = make_float4(T[o][j], T[o + 1][j], T[o + 2][j], T[o + 3][j]), hi, lo;
    split4(v, hi, lo);
    float* p = dst + kb * (128 * 32) + j * 32 + (((c ^ j) & 7) << 2);
    *reinterpret_cast<float4*>(p) = hi;
    *reinterpret_cast<float4*>(p + 64 * 32) = lo;
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

static size_t tucker_tc_img_bytes(const ckb_step_desc_t& d) {
  return (size_t)d.num_folds * KK * 2 * 128 * 32 * 4;  // split weight image, 2 MB per fold
}
static size_t tucker_tc_scratch_bytes(const ckb_step_desc_t& d, int64_t B) {
  const int64_t nblk = (B + ROWS - 1) / ROWS * (ROWS / 32);
  return (size_t)d.num_folds * nblk * kBlkFloats * 4;
}
static size_t tucker_tc_fwd_ws(const ckb_step_desc_t& d) { return tucker_tc_img_bytes(d); }
static size_t tucker_tc_bwd_ws(const ckb_step_desc_t& d, int64_t B) {
  return tucker_tc_img_bytes(d) + tucker_tc_scratch_bytes(d, B);  // image first, then the scratch
}
size_t tucker_tc_ws(const ckb_step_desc_t& d, int64_t B) {
  const size_t a = tucker_tc_bwd_ws(d, B), b = tucker_tc_fwd_ws(d);
  return a > b ? a : b;
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
  static PerDeviceOnce attr;
  const size_t smem = sizeof(TkFwdSmem) + 1024;
  if (attr.first()) {
    if (int rc = set_smem(tucker_tc_fwd_kernel<true>, smem)) return rc;
    if (int rc = set_smem(tucker_tc_fwd_kernel<false>, smem)) return rc;
  }
  const DenseArgs a = tucker_tc_args(d, c);
  if (c.ws_bytes < tucker_tc_fwd_ws(d)) {
    set_error("tucker_fwd: workspace too small (%zu < %zu)", c.ws_bytes, tucker_tc_fwd_ws(d));
    return CKB_ERR_WORKSPACE;
  }
  float* wimg = (float*)c.ws;
  const int64_t n16 = (int64_t)d.num_folds * KK * KRED / 4;
  tucker_split_w_kernel<<<(int)min64((n16 + 255) / 256, 16 * kNumSMs), 256, 0, c.stream>>>(a.W, wimg, n16);
  CKB_LAUNCH_CHECK();
  c.launches++;
  dim3 grid(ceil_div(c.B, TM), d.num_folds);
  if ((tc_flags() & 3) == 3)
    tucker_tc_fwd_kernel<true><<<grid, kFwdThreads, smem, c.stream>>>(a, wimg);
  else
    tucker_tc_fwd_kernel<false><<<grid, kFwdThreads, smem, c.stream>>>(a, wimg);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

int tucker_tc_bwd(const ckb_step_desc_t& d, Ctx& c) {
  static PerDeviceOnce attr;
  const size_t smem_dx = sizeof(TkDxSmem) + 1024, smem_dw = sizeof(TkDwSmem) + 1024;
  if (attr.first()) {
    if (int rc = set_smem(tucker_tc_bwd_dx_kernel<true>, smem_dx)) return rc;
    if (int rc = set_smem(tucker_tc_bwd_dx_kernel<false>, smem_dx)) return rc;
    if (int rc = set_smem(tucker_tc_bwd_dw_kernel, smem_dw)) return rc;
  }
  DenseArgs a = tucker_tc_args(d, c);
  a.gs = GradSrc{c.garena, d.cons_ptr, d.cons_rows, c.B};
  a.gin = c.garena + c.B * d.gin_off;
  float* dW = c.grads[d.slot[0]];
  const int nblk_alloc = ceil_div(c.B, ROWS) * (ROWS / 32);
  const size_t need = tucker_tc_bwd_ws(d, c.B);
  if (c.ws_bytes < need) {
    set_error("tucker_bwd: workspace too small (%zu < %zu)", c.ws_bytes, need);
    return CKB_ERR_WORKSPACE;
  }
  float* wimg = (float*)c.ws;
  float* scratch = (float*)(c.ws + tucker_tc_img_bytes(d));
  {
    dim3 gridp(nblk_alloc, d.num_folds);
    if ((tc_flags() & 3) == 3) tucker_prep_kernel<true><<<gridp, 256, 0, c.stream>>>(a, scratch, nblk_alloc);
    else tucker_prep_kernel<false><<<gridp, 256, 0, c.stream>>>(a, scratch, nblk_alloc);
    CKB_LAUNCH_CHECK();
    c.launches++;
  }
  tucker_split_wt_kernel<<<dim3(KK, d.num_folds), 256, 0, c.stream>>>(a.W, wimg);
  CKB_LAUNCH_CHECK();
  c.launches++;
  dim3 grid(ceil_div(c.B, ROWS), d.num_folds);
  if ((tc_flags() & 3) == 3)
    tucker_tc_bwd_dx_kernel<true><<<grid, kDxThreads, smem_dx, c.stream>>>(a, scratch, nblk_alloc, wimg);
  else
    tucker_tc_bwd_dx_kernel<false><<<grid, kDxThreads, smem_dx, c.stream>>>(a, scratch, nblk_alloc, wimg);
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
