// Thin inline-PTX wrappers for the sm_100a features the tensor-core kernels use:
// mbarrier, bulk async copies (TMA engine, 1-D), tcgen05.mma / tcgen05.ld / TMEM allocation.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptors" (the same fields
// CUTLASS' cute/arch/mma_sm100_desc.hpp names).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ckb {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// try_wait blocks in hardware for a short, implementation-defined time and returns whether the
// phase has completed.  (Measured on B200: passing a long suspend-time hint makes waiters wake
// up several microseconds late when the arrival comes from tcgen05.commit, so the latency
// critical waits poll without a hint; waits that may be slow back off with nanosleep so that
// they do not take issue slots from the warps doing the work.)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spins until the phase with the given parity completes; traps instead of hanging forever
// (a dead-locked pipeline must surface as a launch error, not as a stuck GPU).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(128);
    if (++spins > (1u << 24)) __trap();
  }
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- bulk copy global -> shared
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- TMEM allocation
// One full warp calls these.  ncols: power of two >= 32.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor for a 128-byte-swizzled tile whose rows (128 bytes = 32 fp32)
// are stored densely: 8-row groups are `sbo` bytes apart, 32-element column blocks `lbo` bytes
// apart.  Every operand in this library is K-major.  (Whether tf32 also accepts the same image as
// the MN-major view of the transpose -- which would remove the register transposes of the
// backward kernels -- is what scripts/micro/mn_major_probe.cu is there to find out.)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// MN-major kind::tf32 operand (M or N index contiguous, one 128-byte row of 32 elements per k):
// layout type SWIZZLE_128B_BASE32B -- 32-byte chunks xor (k & 3); `atom_bytes` = distance between
// consecutive groups of 32 M/N elements (LBO), `kgroup_bytes` = distance between groups of 4 k-rows
// (SBO, 512 for dense rows).  A k-step of 8 rows advances the start address by 1024 bytes.
// Established on the B200 by scripts/micro/mn_major_probe.cu and umma_probe2.cu; K-major operands
// do not accept this layout type ("misaligned address").
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t atom_bytes,
                                                 uint32_t kgroup_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((atom_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((kgroup_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}

// The descriptor of the tile `off_bytes` further on in shared memory (same LBO / SBO / swizzle):
// a 32-bit add on the low word.  Shared addresses are below 2^18, so the 14-bit start-address
// field cannot carry into its neighbours.
__device__ __forceinline__ uint64_t desc_at(uint64_t base, uint32_t off_bytes) {
  const uint32_t lo = (uint32_t)base + (off_bytes >> 4);
  return (base & 0xFFFFFFFF00000000ull) | lo;
}

// Instruction descriptor, kind::tf32, fp32 accumulate.  a_mn / b_mn: 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand read from TMEM (lanes = rows, one 32-bit column per K element).
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Whole-warp variants: every lane of a converged warp executes the statement with warp-uniform
// operands and the instruction itself is predicated on the elected lane, which lets the compiler
// keep descriptors in uniform registers instead of serialising a divergent single-lane branch.
__device__ __forceinline__ void mma_tf32_ts_warp(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
          smem_u32(bar))
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- TMEM -> registers
// 32 lanes x 32 bit, 16 consecutive columns: thread `lane` of the warp receives
// D[lane_base + lane][col .. col+15].
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// registers -> TMEM: thread `lane` of the warp writes 16 consecutive columns of its lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
      "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- fp32 -> (tf32 hi, tf32 lo)
// x = hi + lo exactly; hi is x rounded to tf32 (10 explicit mantissa bits, ties away from zero:
// add half a tf32 ulp to the bit pattern and clear the low 13 bits -- two integer ALU ops, where
// cvt.rna.tf32.f32 would go through the quarter-rate conversion pipe), lo = x - hi is exact in
// fp32.  The tensor core reads only the upper 19 bits of each operand, i.e. it truncates lo to
// tf32; lo is symmetric around zero, so that truncation is unbiased and costs 2^-22 |x|.
// Three tf32 products hi*hi + lo*hi + hi*lo then carry fp32-grade accuracy.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

// exp(x) on the MUFU for finite x: e = 2^t with t = fl(x*log2e), times (1 + d) where
// d = x - t*ln2 is the part of the exponent the rounding of t lost (Cody-Waite, two fma's with
// ln2 split into fl(ln2) + tail), so the relative error is that of ex2.approx (2^-22) instead
// of growing with |x|.  5 instructions instead of ~30 for expf.  The caller clamps x into the
// finite range first (-104 flushes to 0, 88 stays finite).
__device__ __forceinline__ float fast_exp_finite(float x) {
  const float t = x * 1.4426950408889634f;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  float d = fmaf(t, -0.6931471805599453f, x);
  d = fmaf(t, 1.9046542e-9f, d);  // fl(ln2) - ln2
  return fmaf(e, d, e);
}
__device__ __forceinline__ float fast_exp(float x) {
  return fast_exp_finite(fminf(fmaxf(x, -104.f), 88.f));
}
// log(x) on the MUFU: lg2.approx * ln2 (absolute error ~2^-22 * |log2 x| + 2^-24).  NOT the .ftz
// form: a sum of products may legitimately be subnormal (the reference's edge case
// tests/backend/torch/test_semiring.py:41-61: weight 1e-38 -> log(1e-38) = -87.5, finite), and the
// tensor core keeps subnormal operands and products (scripts/micro/umma_probe2.cu, test 4).
__device__ __forceinline__ float fast_log(float x) {
  float l;
  asm("lg2.approx.f32 %0, %1;" : "=f"(l) : "f"(x));
  return l * 0.6931471805599453f;
}

// Explicit shared-state-space accesses with 32-bit addresses (the operand tiles are addressed
// through computed byte offsets; generic pointers would cost 64-bit address math and LD/ST
// instead of LDS/STS).
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// Byte offset of element (row, col) inside one 128B-swizzled block of [rows][32 fp32].
__device__ __forceinline__ uint32_t swz_off(uint32_t row, uint32_t col) {
  return row * 128u + ((((col >> 2) ^ row) & 7u) << 4) + ((col & 3u) << 2);
}

}  // namespace sm100
}  // namespace ckb
