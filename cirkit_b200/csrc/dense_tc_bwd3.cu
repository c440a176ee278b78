// Backward of the fused sum-product block, Ki = Ko = 64 -- third version (round 2): TMA-fed.
//
//   r = g * exp(m - y),  e = exp(u - m),  T = r W,  du = e .* T,  dW += r^T e        (3xTF32)
//
// What the round-2 timelines of the register-fed kernels showed (profiles/r02_timeline_bwd2.txt):
// with every thread issuing its own global loads and stores the LSU is the bottleneck -- 32 LDG.64
// take 2 000-5 000 clocks to ISSUE, the du stores 1 000-2 400, and every phase of the CTA waits on
// the same queue; the arithmetic of a 128-sample tile is 6 500 clocks without any memory traffic.
// Here the memory traffic does not go through the LSU at all:
//   * a producer warp streams x0, x1, y, g through the TMA engine (cp.async.bulk.tensor.2d, boxes
//     of 32 rows x 128 bytes, SWIZZLE_128B_ATOM_32B) into a 3-slot ring, always 2 slots (64 KB)
//     ahead of the workers;
//   * du leaves through a swizzled staging tile and two TMA tensor stores per tile;
//   * the workers read / write shared memory only (conflict-free 8-byte accesses: the swizzle is
//     the same 32-byte-chunk xor (row & 3) as the MN-major UMMA images).
// Tiles are 64 samples (shared memory: ring 96 KB + MN images 64 KB + W 32 KB + du 16 KB):
//   GEMM 1 (T = r W): M = 64, A = r in TENSOR MEMORY (TS form; rows 16q.. live in lanes 32q..32q+15,
//           the same lanes the M = 64 accumulator uses), B = W^T image, two accumulator buffers;
//   GEMM 2 (dW += r^T e): M = N = 128 (hi | lo stacked), both operands MN-major images, K = 64.
// Element ownership is the 16x256b TMEM fragment for everything (loads from the ring, r -> TMEM,
// images, read-back of T, du): warp w = (quadrant q = w & 3, column group cg = w >> 2, 16 columns),
// lane = (r8 = lane >> 2, c2 = lane & 3): rows 16q + r8 (+8), columns 16cg + 8n + 2c2 (+1).
// Software pipeline: the du epilogue of tile t-1 runs after the operands of tile t are published,
// so GEMM 1's latency is hidden behind it.
//
// Preconditions (else dense_tc.cu's register-fed kernel runs): B % 64 == 0, every fold has exactly
// one consumer row, all row offsets are multiples of 64 floats (CKB_STEP_ROWS64).
#include <cuda.h>

#include "dense.cuh"
#include "sm100.cuh"
#include "tc_util.cuh"

namespace ckb {
using namespace sm100;

namespace {

constexpr int TM3 = 64;  // samples per tile
constexpr int KK = 64;
constexpr int kStages = 3;
constexpr int kStageRows = 32;
constexpr int kWorkers3 = 16;
constexpr int kMmaWarp3 = 16, kLoadWarp3 = 17, kStoreWarp3 = 18;
constexpr int kThreads3 = 19 * 32;
constexpr uint32_t kWBlock3 = 128 * 128;

struct __align__(1024) Bwd3Smem {
  float raw[kStages][4][2][kStageRows * 32];  // [slot][x0,x1,y,g][column half][row][32]  swizzled  96 KB
  float r_mn[4][TM3 * 32];  // [r_hi 0..31 | r_hi 32..63 | r_lo 0..31 | r_lo 32..63][sample][32]     32 KB
  float e_mn[4][TM3 * 32];  //                                                                       32 KB
  float w[2][128 * 32];     // W^T: [o-block][hi i 0..63 | lo i 0..63][32 o's]                       32 KB
  float du[2][TM3 * 32];    // [column half][sample][32] swizzled: TMA store source                 16 KB
  float mbuf[2][4][TM3];    // [tile parity][column group][row]: quarter-row maxima
  uint64_t full[kStages], empty[kStages];
  uint64_t ab_full, img_free, d1_full, d1_empty, du_full, du_empty, d2_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm),
               "r"(c0), "r"(c1), "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
// 16 TMEM lanes x 16 columns <-> 8 registers per thread: reg 4n + 2a + c = (row t/4 + 8a,
// column 8n + 2(t%4) + c)
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
      : "memory");
}

// W^T image: rows i, K = o, stacked per k-block of 32 o's as [hi rows 0..63 | lo rows 0..63]
template <int NTHREADS>
__device__ __forceinline__ void stage_wt(const float* Wf, uint32_t w, int tid) {
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
      const float h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t off = (uint32_t)(o >> 5) * kWBlock3 + swz_off(i0 + t, o & 31);
        sts32(w + off, h[t]);
        sts32(w + off + KK * 128, l[t]);
      }
    }
  }
}

__device__ unsigned long long g_bwd3_dbg[8];

template <bool FAST>
__global__ void __launch_bounds__(kThreads3, 1)
dense_tc_bwd3_kernel(DenseArgs a, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmG,
                     const __grid_constant__ CUtensorMap tmDU, int tiles_per_cta, int want_dw, int flags) {
  extern __shared__ uint8_t smem_raw[];
  Bwd3Smem& s = *reinterpret_cast<Bwd3Smem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int n_tiles_total = (int)(a.B / TM3);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(n_tiles_total, t_begin + tiles_per_cta) - t_begin;
  if (n_tiles <= 0) return;

  // The producer lane sets up its own ring barriers and puts the first kStages slots in flight
  // BEFORE the weight image is staged: the HBM latency of the first tile (~2 us) then overlaps the
  // CTA's set-up instead of following it (a CTA lives for 10-70 us, the set-up used to be ~8).
  int64_t x0_row = 0, x1_row = 0, y_row = 0, g_row = 0;
  int preissued = 0;
  const uint32_t slot_bytes = (a.H == 2 ? 4u : 3u) * 2u * kStageRows * 128u;
  auto issue_slot = [&](int i) {
    const int slot = i % kStages;
    mbar_arrive_expect_tx(&s.full[slot], slot_bytes);
    const int row = (t_begin * TM3) + i * kStageRows;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      tma_load_2d(s.raw[slot][0][ch], &tmX, 32 * ch, (int)(x0_row + row), &s.full[slot]);
      if (a.H == 2) tma_load_2d(s.raw[slot][1][ch], &tmX, 32 * ch, (int)(x1_row + row), &s.full[slot]);
      tma_load_2d(s.raw[slot][2][ch], &tmY, 32 * ch, (int)(y_row + row), &s.full[slot]);
      tma_load_2d(s.raw[slot][3][ch], &tmG, 32 * ch, (int)(g_row + row), &s.full[slot]);
    }
  };
  if (warp == kLoadWarp3 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], kWorkers3);  // every worker passes every slot (see the ring protocol below)
    }
    fence_barrier_init();
    pdl_wait();  // y, g (and the arenas behind x) are outputs of earlier grids of the stream
    // rows are addressed as rows of 64 floats relative to the base of each tensor map
    x0_row = a.in_rows ? (a.B * a.in_rows[f * a.H]) / KK : (int64_t)f * a.B;
    x1_row = a.H == 2 ? (a.B * a.in_rows[f * a.H + 1]) / KK : 0;
    y_row = (int64_t)f * a.B;
    g_row = a.gs.cons_ptr ? (a.gs.B * a.gs.cons_rows[a.gs.cons_ptr[f]]) / KK : (int64_t)f * a.gs.B;
    preissued = min(kStages, 2 * n_tiles);
    for (int i = 0; i < preissued; ++i) issue_slot(i);
  }
  if (tid == 0) {
    mbar_init(&s.ab_full, kWorkers3);
    mbar_init(&s.img_free, 1);
    mbar_init(&s.d1_full, 1);
    mbar_init(&s.d1_empty, kWorkers3);
    mbar_init(&s.du_full, kWorkers3);
    mbar_init(&s.du_empty, 1);
    mbar_init(&s.d2_full, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp3) tmem_alloc(&s.tmem_base, 512);  // 384 used (allocations are powers of two)
  stage_wt<kThreads3>(a.W + (int64_t)f * KK * KK, smem_u32(s.w), tid);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s.tmem_base;
  pdl_wait();  // everything below may touch global memory the preceding grid wrote (du, dW slabs)
  pdl_launch_dependents();  // (after the wait: see common.cuh)
  // TMEM columns: D1 [0,128) (main | correction), D2 [128,256), r_hi [256,320), r_lo [320,384)
  constexpr uint32_t kD2Col = 128, kRhiCol = 256, kRloCol = 320;

  if (warp == kLoadWarp3) {
    // ================= producer: TMA loads, 32 rows x 4 arrays per ring slot =================
    if (lane == 0) {
      for (int i = preissued; i < 2 * n_tiles; ++i) {
        mbar_wait_relaxed(&s.empty[i % kStages], ((i / kStages) & 1) ^ 1);
        issue_slot(i);
      }
    }
  } else if (warp == kStoreWarp3) {
    // ================= du: TMA stores of the staged tile =================
    if (lane == 0 && !(flags & 16384)) {
      const int64_t du_row = (int64_t)f * a.B + (int64_t)t_begin * TM3;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait_relaxed(&s.du_full, t & 1);
        tma_store_2d(&tmDU, 0, (int)(du_row + (int64_t)t * TM3), s.du[0]);
        tma_store_2d(&tmDU, 32, (int)(du_row + (int64_t)t * TM3), s.du[1]);
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(&s.du_empty);
      }
      bulk_wait<0>();
    }
  } else if (warp == kMmaWarp3) {
    // ================= MMA issuer (converged warp, lane-elected instructions) =================
    constexpr uint32_t idesc_n128 = make_idesc_tf32(TM3, 2 * KK, 0, 0);
    constexpr uint32_t idesc_n64 = make_idesc_tf32(TM3, KK, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_tf32(128, 128, 1, 1);
    const uint64_t d_w = make_desc(smem_u32(s.w), 16, 1024);
    const uint64_t d_r = make_desc_mn(smem_u32(s.r_mn), TM3 * 128, 512);
    const uint64_t d_e = make_desc_mn(smem_u32(s.e_mn), TM3 * 128, 512);
    for (int t = 0; t < n_tiles; ++t) {
      const uint32_t d1 = tmem_base;
      mbar_wait(&s.ab_full, t & 1);
      // every worker has read T of the previous tile out of D1.  (Measured on the B200: a
      // tcgen05.ld that overlaps an M = 64 tcgen05.mma writing OTHER columns occasionally returns
      // garbage -- a second D1 buffer that would let the two overlap was dropped for that reason;
      // the read-back takes ~150 clocks, the rest of the epilogue still overlaps GEMM 1.)
      if (t >= 1) mbar_wait(&s.d1_empty, (t - 1) & 1);
      tc_fence_after_sync();
      // GEMM 1: r_hi x [W_hi | W_lo] (N = 128): main | correction;  r_lo x W_hi (N = 64): correction
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        mma_tf32_ts_warp(d1, tmem_base + kRhiCol + 8 * ks, desc_at(d_w, (ks >> 2) * kWBlock3 + (ks & 3) * 32),
                         idesc_n128, ks ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        mma_tf32_ts_warp(d1 + KK, tmem_base + kRloCol + 8 * ks,
                         desc_at(d_w, (ks >> 2) * kWBlock3 + (ks & 3) * 32), idesc_n64, 1u);
      mma_commit_warp(&s.d1_full);
      if (want_dw) {
        // GEMM 2: dW[o,i] += sum_b r[b,o] e[b,i], 8 samples per instruction
#pragma unroll
        for (int ks = 0; ks < TM3 / 8; ++ks)
          mma_tf32_warp(tmem_base + kD2Col, desc_at(d_r, ks * 1024), desc_at(d_e, ks * 1024), idesc2,
                        (t || ks) ? 1u : 0u);
      }
      mma_commit_warp(&s.img_free);
      if (t + 1 == n_tiles) mma_commit_warp(&s.d2_full);
    }
  } else {
    // ================= workers =================
    const int q = warp & 3, cg = warp >> 2;
    const int r8 = lane >> 2, c2 = lane & 3;
    const int trow = 16 * q + r8;  // row a = 0 inside the tile; a = 1 adds 8
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;  // M = 64: rows 16q.. live in lanes 32q..32q+15
    const uint32_t rsw = (uint32_t)(trow & 3);
    const uint32_t chunk0 = 2u * (cg & 1);  // 32-byte chunk of column group n = 0 inside a 128-byte row
    // byte offset of (row a, n = 0) in a [column half][row][128 B] swizzled block, before the chunk xor
    //   ring slot: rows 0..31 of the slot = tile rows 32h..32h+31, h = q >> 1
    uint32_t raw_off[2], img_off[2];
#pragma unroll
    for (int aa = 0; aa < 2; ++aa) {
      raw_off[aa] = (uint32_t)(cg >> 1) * (kStageRows * 128) + (uint32_t)((trow + 8 * aa) & 31) * 128u + c2 * 8u;
      img_off[aa] = (uint32_t)(cg >> 1) * (TM3 * 128) + (uint32_t)(trow + 8 * aa) * 128u + c2 * 8u;
    }
    const uint32_t raw_base = smem_u32(s.raw);
    const uint32_t rmn = smem_u32(s.r_mn), emn = smem_u32(s.e_mn), dus = smem_u32(s.du);
    constexpr uint32_t kArr = 2 * kStageRows * 128;  // bytes per array in a slot
    constexpr uint32_t kSlot = 4 * kArr;
    constexpr uint32_t kLo = 2 * TM3 * 128;          // hi -> lo inside an image
    const bool two = a.H == 2;

    float e_prev[2][2][2];
    for (int t = 0; t <= n_tiles; ++t) {
      float e[2][2][2], hi[8], lo[8];
      if (t < n_tiles) {
        // ---- this tile's rows from the ring.  Protocol: EVERY worker waits for BOTH 32-row slots of
        // the tile, in order, and releases both -- it reads only the one that holds its rows.  An
        // mbarrier parity wait tells "phase k complete" from "not yet" only for a waiter that is at
        // most one phase away: with consumers-only waits a warp of one half could ask for use k+1
        // of a slot while use k (the other half's rows, still in flight) was incomplete, and the
        // parity of k+1 then reads as "complete" (observed on the B200 as rare garbage and hangs).
        float2 x[2][2], yv[2][2], gv[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 2 * t + h;
          const int slot = i % kStages;
          mbar_wait(&s.full[slot], (i / kStages) & 1);
          if (h == (q >> 1)) {
            const uint32_t sb = raw_base + slot * kSlot;
#pragma unroll
            for (int aa = 0; aa < 2; ++aa)
#pragma unroll
              for (int n = 0; n < 2; ++n) {
                const uint32_t o = sb + raw_off[aa] + (((chunk0 + n) ^ rsw) << 5);
                x[aa][n] = lds64(o);
                if (two) {
                  const float2 z = lds64(o + kArr);
                  x[aa][n].x += z.x;
                  x[aa][n].y += z.y;
                }
                yv[aa][n] = lds64(o + 2 * kArr);
                gv[aa][n] = lds64(o + 3 * kArr);
              }
            // the loads have landed in registers before the slot is released
            if (__float_as_uint(x[0][0].x + x[0][1].x + x[1][0].x + x[1][1].x + yv[0][0].x + yv[0][1].x +
                                yv[1][0].x + yv[1][1].x + gv[0][0].x + gv[0][1].x + gv[1][0].x +
                                gv[1][1].x) == 0x7fc12345u)
              s.mbuf[0][0][0] = 0.f;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&s.empty[slot]);
        }
#ifdef CKB_BWD3_CHECK
        if (flags & 4096) {  // debug: the ring against direct global loads of the same elements
          const float* row0 = in_row(a, f, 0);
          const float* row1 = a.H == 2 ? in_row(a, f, 1) : nullptr;
          const float* yrow = a.y + (int64_t)f * a.B * KK;
          const float* grow = a.gs.cons_ptr ? a.gs.garena + a.gs.B * a.gs.cons_rows[a.gs.cons_ptr[f]]
                                            : a.gs.garena + (int64_t)f * a.gs.B * KK;
#pragma unroll
          for (int aa = 0; aa < 2; ++aa)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
              const int64_t e0 = ((int64_t)(t_begin + t) * TM3 + trow + 8 * aa) * KK + 16 * cg + 8 * n + 2 * c2;
              float ux = row0[e0], uy = row0[e0 + 1];
              if (row1) { ux += row1[e0]; uy += row1[e0 + 1]; }
              if (ux != x[aa][n].x || uy != x[aa][n].y) atomicAdd(&g_bwd3_dbg[0], 1ull);
              if (yrow[e0] != yv[aa][n].x || yrow[e0 + 1] != yv[aa][n].y) atomicAdd(&g_bwd3_dbg[1], 1ull);
              if (grow[e0] != gv[aa][n].x || grow[e0 + 1] != gv[aa][n].y) atomicAdd(&g_bwd3_dbg[2], 1ull);
            }
          if (tid == 0) atomicAdd(&g_bwd3_dbg[7], 1ull);
        }
#endif
        // ---- row max over the 64 columns: 4 lanes, then the 4 warps of this quadrant
        float mloc[2];
#pragma unroll
        for (int aa = 0; aa < 2; ++aa) {
          float m = fmaxf(fmaxf(x[aa][0].x, x[aa][0].y), fmaxf(x[aa][1].x, x[aa][1].y));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          mloc[aa] = m;
          if (c2 == 0) s.mbuf[t & 1][cg][trow + 8 * aa] = m;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
        float rr[2][2][2];
#pragma unroll
        for (int aa = 0; aa < 2; ++aa) {
          const int r = trow + 8 * aa;
          float m = fmaxf(fmaxf(s.mbuf[t & 1][0][r], s.mbuf[t & 1][1][r]),
                          fmaxf(s.mbuf[t & 1][2][r], s.mbuf[t & 1][3][r]));
          m = clamp_max(fmaxf(m, mloc[aa]));
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            e[aa][n][0] = exp_nonpos<FAST>(x[aa][n].x - m);
            e[aa][n][1] = exp_nonpos<FAST>(x[aa][n].y - m);
            rr[aa][n][0] = gv[aa][n].x * exp_capped<FAST>(m - yv[aa][n].x);
            rr[aa][n][1] = gv[aa][n].y * exp_capped<FAST>(m - yv[aa][n].y);
          }
        }
#pragma unroll
        for (int aa = 0; aa < 2; ++aa)
#pragma unroll
          for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int c = 0; c < 2; ++c) split_tf32(rr[aa][n][c], hi[4 * n + 2 * aa + c], lo[4 * n + 2 * aa + c]);
        // GEMM 1 of the previous tile has consumed r (tensor memory), GEMM 2 the images
        if (t >= 1) {
          mbar_wait(&s.d1_full, (t - 1) & 1);
          mbar_wait(&s.img_free, (t - 1) & 1);
          tc_fence_after_sync();
        }
        tmem_st_16x256b_x2(tmem_base + lane_addr + kRhiCol + 16 * cg, hi);
        tmem_st_16x256b_x2(tmem_base + lane_addr + kRloCol + 16 * cg, lo);
#pragma unroll
        for (int aa = 0; aa < 2; ++aa)
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            const uint32_t o = img_off[aa] + (((chunk0 + n) ^ rsw) << 5);
            sts64(rmn + o, hi[4 * n + 2 * aa], hi[4 * n + 2 * aa + 1]);
            sts64(rmn + o + kLo, lo[4 * n + 2 * aa], lo[4 * n + 2 * aa + 1]);
            float h0, l0, h1, l1;
            split_tf32(e[aa][n][0], h0, l0);
            split_tf32(e[aa][n][1], h1, l1);
            sts64(emn + o, h0, h1);
            sts64(emn + o + kLo, l0, l1);
          }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.ab_full);
      }
      const bool serial = (flags & 8192) != 0;
      if (serial && t < n_tiles) {
#pragma unroll
        for (int aa = 0; aa < 2; ++aa)
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            e_prev[aa][n][0] = e[aa][n][0];
            e_prev[aa][n][1] = e[aa][n][1];
          }
      }
      if (serial ? t < n_tiles : t >= 1) {
        // ---- du epilogue of tile t-1: du = e * (main + correction) -> staging tile -> TMA store
        const int tp = serial ? t : t - 1;
        if (serial || t == n_tiles) {  // (inside the loop the wait happened before the operands were overwritten)
          mbar_wait(&s.d1_full, tp & 1);
          tc_fence_after_sync();
        }
        float v[8], w[8];
        const uint32_t d1 = tmem_base + lane_addr + 16 * cg;
        tmem_ld_16x256b_x2(d1, v);
        tmem_ld_16x256b_x2(d1 + KK, w);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.d1_empty);
        if (flags & 16384) {  // debug variant: du straight from the registers
          float* pdu = a.gin + ((int64_t)f * a.B + (int64_t)(t_begin + tp) * TM3 + trow) * KK + 16 * cg + 2 * c2;
#pragma unroll
          for (int aa = 0; aa < 2; ++aa)
#pragma unroll
            for (int n = 0; n < 2; ++n)
              *reinterpret_cast<float2*>(pdu + aa * 8 * KK + 8 * n) =
                  make_float2(e_prev[aa][n][0] * (v[4 * n + 2 * aa] + w[4 * n + 2 * aa]),
                              e_prev[aa][n][1] * (v[4 * n + 2 * aa + 1] + w[4 * n + 2 * aa + 1]));
        } else {
        if (tp >= 1) mbar_wait(&s.du_empty, (tp - 1) & 1);  // the previous tile has left the staging area
#pragma unroll
        for (int aa = 0; aa < 2; ++aa)
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            const uint32_t o = img_off[aa] + (((chunk0 + n) ^ rsw) << 5);
            sts64(dus + o, e_prev[aa][n][0] * (v[4 * n + 2 * aa] + w[4 * n + 2 * aa]),
                  e_prev[aa][n][1] * (v[4 * n + 2 * aa + 1] + w[4 * n + 2 * aa + 1]));
          }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.du_full);
        }
      }
      if (serial && t == n_tiles - 1) break;
#pragma unroll
      for (int aa = 0; aa < 2; ++aa)
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          e_prev[aa][n][0] = e[aa][n][0];
          e_prev[aa][n][1] = e[aa][n][1];
        }
    }
    if (want_dw) {
      // D2 quadrants: rows 0..63 = r_hi^T [e_hi | e_lo], rows 64..127 = r_lo^T [e_hi | (dropped)].
      // dW[o][i] = D2[o][i] + D2[o][64+i] + D2[64+o][i]: the lower half goes through shared memory
      // (the images are dead once every MMA has completed).
      mbar_wait(&s.d2_full, 0);
      tc_fence_after_sync();
      const int row = q * 32 + lane;  // 0..127
      const int o = row & 63;
      const uint32_t xch = smem_u32(s.r_mn);  // [64 columns][64 + 1] exchange buffer (16.6 KB)
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
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers3 * 32) : "memory");
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
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp3) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// [rows][64] fp32 matrix at `base`, boxes of box_rows x 32 columns, 32-byte-atom 128-byte swizzle.
// A descriptor depends on (base, box_rows) only and the arenas come back at the same addresses step
// after step (caching allocator): the last 64 encodings are kept per host thread -- the driver
// call costs ~1.5 us and a step needs 40 of them.
bool make_map(CUtensorMap* tm, const void* base, int box_rows) {
  struct Entry {
    const void* base;
    int box_rows;
    CUtensorMap map;
  };
  static thread_local Entry cache[64];
  static thread_local int used = 0, next = 0;
  for (int i = 0; i < used; ++i)
    if (cache[i].base == base && cache[i].box_rows == box_rows) {
      *tm = cache[i].map;
      return true;
    }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {64, (cuuint64_t)1 << 31};
  const cuuint64_t strides[1] = {256};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  if (fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  Entry& e = cache[next];
  e.base = base;
  e.box_rows = box_rows;
  e.map = *tm;
  next = (next + 1) % 64;
  if (used < 64) ++used;
  return true;
}

}  // namespace

int bwd3_debug_read(void* dst, size_t bytes) {
  if (bytes > sizeof(unsigned long long) * 8) bytes = sizeof(unsigned long long) * 8;
  CKB_CUDA_CHECK(cudaMemcpyFromSymbol(dst, g_bwd3_dbg, bytes));
  return CKB_OK;
}

void dense_tc_bwd3_config(int F, int64_t B, int& splits, int& tiles_per_cta) {
  const int n_tiles = (int)(B / TM3);
  // about two CTAs per SM over the launch; a CTA's fixed cost (weight image, TMEM, pipeline fill)
  // is worth about two tiles, so the small top levels are cut down to 2 tiles per CTA
  splits = (int)max64(1, min64(n_tiles / 2, ceil_div(2 * kNumSMs, F)));
  tiles_per_cta = ceil_div(n_tiles, splits);
  splits = ceil_div(n_tiles, tiles_per_cta);
}

bool dense_tc_bwd3_ok(const DenseArgs& a, int rows64) {
  return rows64 && a.B % TM3 == 0 && a.B >= TM3 && a.max_cons == 1 && a.Ki == KK && a.Ko == KK &&
         a.Kred == KK && !a.concat && a.H >= 1 && a.H <= 2 && (tc_flags() & 3) == 3 &&
         !(tc_flags() & 2048) && encode_fn() != nullptr &&
         ((uintptr_t)a.arena % 256) == 0 && ((uintptr_t)a.y % 256) == 0 && ((uintptr_t)a.gs.garena % 256) == 0 &&
         ((uintptr_t)a.gin % 256) == 0;
}

size_t dense_tc_bwd3_ws(int F, int64_t B) {
  if (B % TM3 != 0 || B < TM3) return 0;
  int splits, tpc;
  dense_tc_bwd3_config(F, B, splits, tpc);
  return splits > 1 ? (size_t)splits * F * KK * KK * 4 : 0;
}

int dense_tc_bwd3(const DenseArgs& a_in, int F, float* dW, Ctx& c, char* ws, size_t ws_bytes) {
  DenseArgs a = a_in;
  int splits, tpc;
  dense_tc_bwd3_config(F, a.B, splits, tpc);
  const size_t n = (size_t)F * KK * KK;
  a.dWp = dW;
  if (dW && splits > 1) {
    if (ws_bytes < splits * n * 4) {
      set_error("dense_tc_bwd3: workspace too small (%zu < %zu)", ws_bytes, splits * n * 4);
      return CKB_ERR_WORKSPACE;
    }
    a.dWp = (float*)ws;
  }
  CUtensorMap tmX, tmY, tmG, tmDU;
  if (!make_map(&tmX, a.arena, kStageRows) || !make_map(&tmY, a.y, kStageRows) ||
      !make_map(&tmG, a.gs.garena, kStageRows) || !make_map(&tmDU, a.gin, TM3)) {
    set_error("dense_tc_bwd3: cuTensorMapEncodeTiled failed");
    return CKB_ERR_CUDA;
  }
  const size_t smem = sizeof(Bwd3Smem) + 1024;
  static PerDeviceOnce attr;
  if (attr.first())
    CKB_CUDA_CHECK(cudaFuncSetAttribute(dense_tc_bwd3_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(splits, F);
  CKB_CUDA_CHECK(launch_pdl(dense_tc_bwd3_kernel<true>, grid, dim3(kThreads3), smem, c.stream, a, tmX, tmY, tmG,
                            tmDU, tpc, dW ? 1 : 0, tc_flags()));
  c.launches++;
  if (dW && splits > 1) return reduce_partials(a.dWp, dW, (int64_t)n, splits, c);
  return CKB_OK;
}

}  // namespace ckb
