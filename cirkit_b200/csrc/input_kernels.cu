// Input-layer kernels: evidence transpose, table (Categorical / Embedding), Gaussian, constant.
#include "common.cuh"

namespace ckb {

// ------------------------------------------------------------------------------------------
// x (B, D) -> xT (D, B).  The reference re-gathers x for every input layer
// (`x[..., scope_idx].permute(1, 0, 2)`, circuits.py:66); here it is transposed and narrowed
// once so that every later read of "variable v of samples b..b+31" is one coalesced line.
// ------------------------------------------------------------------------------------------
template <typename T, typename O>
__global__ void transpose_kernel(const T* __restrict__ x, int64_t B, int D, int64_t ld,
                                 O* __restrict__ xT) {
  __shared__ O tile[32][33];
  const int64_t b0 = (int64_t)blockIdx.x * 32;
  const int d0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int64_t b = b0 + j;
    const int d = d0 + threadIdx.x;
    if (b < B && d < D) tile[j][threadIdx.x] = (O)x[b * ld + d];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int d = d0 + j;
    const int64_t b = b0 + threadIdx.x;
    if (b < B && d < D) xT[(int64_t)d * B + b] = tile[threadIdx.x][j];
  }
}

template <typename T, typename O>
static int launch_transpose(const void* x, int64_t B, int D, int64_t ld, void* xT, cudaStream_t s) {
  if (B == 0 || D == 0) return CKB_OK;
  dim3 grid(ceil_div(B, 32), ceil_div(D, 32)), block(32, 8);
  transpose_kernel<T, O><<<grid, block, 0, s>>>((const T*)x, B, D, ld, (O*)xT);
  CKB_LAUNCH_CHECK();
  return CKB_OK;
}

int transpose_input(const void* x, int dtype, int64_t B, int D, int64_t ld, void* xT, cudaStream_t s) {
  switch (dtype) {
    case CKB_U8: return launch_transpose<uint8_t, int32_t>(x, B, D, ld, xT, s);
    case CKB_I16: return launch_transpose<int16_t, int32_t>(x, B, D, ld, xT, s);
    case CKB_I32: return launch_transpose<int32_t, int32_t>(x, B, D, ld, xT, s);
    case CKB_I64: return launch_transpose<int64_t, int32_t>(x, B, D, ld, xT, s);
    case CKB_F32: return launch_transpose<float, float>(x, B, D, ld, xT, s);
    case CKB_F64: return launch_transpose<double, float>(x, B, D, ld, xT, s);
  }
  set_error("ckb_transpose_input: unknown dtype %d", dtype);
  return CKB_ERR_INVALID;
}

int transpose_mask(const uint8_t* m, int64_t rows, int D, uint8_t* mT, cudaStream_t s) {
  return launch_transpose<uint8_t, uint8_t>(m, rows, D, D, mT, s);
}

// ------------------------------------------------------------------------------------------
// Table lookup: y[f,b,:] = T[f, x[b,var_f], :]  (T is the (F,V,K) log-table a parameter op
// produced).  Pure gather: one sample reads K contiguous floats.
// ------------------------------------------------------------------------------------------
__global__ void table_fwd_kernel(const float* __restrict__ T, const int32_t* __restrict__ scope_var,
                                 const void* __restrict__ xT, int x_is_float,
                                 const uint8_t* __restrict__ maskT, int64_t mask_ld,
                                 const float* __restrict__ integ, float* __restrict__ y, int64_t B,
                                 int K, int V) {
  const int f = blockIdx.y;
  const int var = scope_var[f];
  const float* Tf = T + (int64_t)f * V * K;
  float* yf = y + (int64_t)f * B * K;
  const float* integ_f = integ ? integ + (int64_t)f * K : nullptr;
  if ((K & 3) == 0) {
    // a group of G = min(32, K/4 rounded up to a power of two) lanes copies one K-float row
    const int K4 = K >> 2;
    int G = 1;
    while (G < K4 && G < 32) G <<= 1;
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const int rows_per_warp = 32 / G;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b0 = warp_global * rows_per_warp; b0 < B; b0 += n_warps * rows_per_warp) {
      const int64_t b = b0 + lane / G;
      if (b >= B) continue;
      const float4* src;
      if (read_mask(maskT, mask_ld, var, b)) {
        src = reinterpret_cast<const float4*>(integ_f);
      } else {
        int v = read_state(xT, x_is_float, (int64_t)var * B + b);
        v = min(max(v, 0), V - 1);
        src = reinterpret_cast<const float4*>(Tf + (int64_t)v * K);
      }
      float4* dst = reinterpret_cast<float4*>(yf + b * K);
      for (int k4 = gl; k4 < K4; k4 += G)
        dst[k4] = src ? src[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    const int64_t total = B * K;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
      const int64_t b = idx / K;
      const int k = (int)(idx - b * K);
      float val;
      if (read_mask(maskT, mask_ld, var, b)) {
        val = integ_f ? integ_f[k] : 0.f;
      } else {
        int v = read_state(xT, x_is_float, (int64_t)var * B + b);
        v = min(max(v, 0), V - 1);
        val = Tf[(int64_t)v * K + k];
      }
      yf[idx] = val;
    }
  }
}

int table_fwd(const ckb_step_desc_t& d, Ctx& c) {
  const int K = d.k_out;
  int bx;
  if (K % 4 == 0) {
    int G = 1;
    while (G < K / 4 && G < 32) G <<= 1;
    const int64_t rows_per_block = 8 * (32 / G);      // 256 threads
    bx = (int)min64(ceil_div(c.B, rows_per_block * 4), 2 * kNumSMs);  // >= 4 rows per lane group
  } else {
    bx = (int)min64(ceil_div(c.B * K, 256), 4 * kNumSMs);
  }
  dim3 grid(max(bx, 1), d.num_folds);
  table_fwd_kernel<<<grid, 256, 0, c.stream>>>(
      c.tensors[d.slot[0]], d.scope_var, c.xT, c.x_is_float, c.maskT, c.mask_ld,
      d.int_slot >= 0 ? c.tensors[d.int_slot] : nullptr, c.arena + c.B * d.out_off, c.B, K,
      d.num_states);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// u[f,b,:] = sum_h T2[fold(f,h), x[b, var(fold(f,h))], :] -- the Hadamard product (log space) of
// H table rows, gathered in one pass (CKB_STEP_TABLE_INPUT).  A group of G lanes owns a row, a
// warp works on U batches of rows at once: all state loads, then all table loads (L2: the table
// is V*K*4 bytes per fold), then the stores.
constexpr int kPairU = 4, kPairMaxH = 4;
__global__ void table_pair_gather_kernel(const float* __restrict__ T2, const int32_t* __restrict__ scope_var,
                                         const int64_t* __restrict__ folds, const void* __restrict__ xT,
                                         int x_is_float, float* __restrict__ u, int64_t B, int K, int V,
                                         int H) {
  const int f = blockIdx.y;
  const int K4 = K >> 2;
  int G = 1;
  while (G < K4 && G < 32) G <<= 1;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), rpw = 32 / G;
  const float* Th[kPairMaxH];
  int var[kPairMaxH];
#pragma unroll
  for (int h = 0; h < kPairMaxH; ++h) {
    const int64_t tf = h < H ? folds[(int64_t)f * H + h] : 0;
    Th[h] = T2 + tf * V * K;
    var[h] = h < H ? scope_var[tf] : 0;
  }
  float* uf = u + (int64_t)f * B * K;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t step = (int64_t)rpw * kPairU;
  for (int64_t b0 = warp_global * step; b0 < B; b0 += n_warps * step) {
    int v[kPairU][kPairMaxH];
#pragma unroll
    for (int q = 0; q < kPairU; ++q) {
      const int64_t b = b0 + q * rpw + lane / G;
#pragma unroll
      for (int h = 0; h < kPairMaxH; ++h)
        v[q][h] = (h < H && b < B) ? min(max(read_state(xT, x_is_float, (int64_t)var[h] * B + b), 0), V - 1) : 0;
    }
    for (int k4 = gl; k4 < K4; k4 += G) {
      float4 acc[kPairU];
#pragma unroll
      for (int q = 0; q < kPairU; ++q) {
        acc[q] = __ldg(reinterpret_cast<const float4*>(Th[0] + (int64_t)v[q][0] * K) + k4);
#pragma unroll
        for (int h = 1; h < kPairMaxH; ++h)
          if (h < H) {
            const float4 z = __ldg(reinterpret_cast<const float4*>(Th[h] + (int64_t)v[q][h] * K) + k4);
            acc[q].x += z.x; acc[q].y += z.y; acc[q].z += z.z; acc[q].w += z.w;
          }
      }
#pragma unroll
      for (int q = 0; q < kPairU; ++q) {
        const int64_t b = b0 + q * rpw + lane / G;
        if (b < B) __stcs(reinterpret_cast<float4*>(uf + b * K) + k4, acc[q]);
      }
    }
  }
}

// Same result through shared memory: a CTA owns (fold, tile of KT <= 32 units), stages the H table
// slices it gathers from -- H*V*KT floats, 64 KB for two 256-state tables -- once, and then serves
// every sample of the fold from there.  The table rows are read from L2 / HBM once per CTA instead
// of once per sample (2*B*K*4 bytes per fold at the far-die L2 rate, which is what bounded the
// kernel above: 125 us for the 392 x 2048 north-star level); what remains is the u stream.
// States are fetched 256 samples at a time, coalesced, one chunk ahead (double-buffered).
constexpr int kStageRows = 256;
__global__ void __launch_bounds__(256) table_pair_gather_smem_kernel(
    const float* __restrict__ T2, const int32_t* __restrict__ scope_var, const int64_t* __restrict__ folds,
    const void* __restrict__ xT, int x_is_float, float* __restrict__ u, int64_t B, int K, int V, int H,
    int KT) {
  extern __shared__ __align__(16) float tab[];                     // [H][V][KT]
  int* st = reinterpret_cast<int*>(tab + (size_t)H * V * KT);      // [2][H][kStageRows] row offsets
  const int f = blockIdx.y, k0 = blockIdx.x * KT;
  const int tid = threadIdx.x;
  const int KT4 = KT >> 2;  // float4 chunks per staged row: 8 for KT = 32
  int var[kPairMaxH];
  int64_t tfs[kPairMaxH];
#pragma unroll
  for (int h = 0; h < kPairMaxH; ++h) {  // (index tables: static, not written inside the step)
    tfs[h] = h < H ? folds[(int64_t)f * H + h] : 0;
    var[h] = h < H ? scope_var[tfs[h]] : 0;
  }
  pdl_wait();
  pdl_launch_dependents();
#pragma unroll
  for (int h = 0; h < kPairMaxH; ++h) {
    const int64_t tf = tfs[h];
    if (h < H) {
      const float* src = T2 + tf * V * K + k0;
      for (int i = tid; i < V * KT4; i += 256) {
        const int v = i / KT4, c = i - v * KT4;
        reinterpret_cast<float4*>(tab)[((size_t)h * V + v) * KT4 + c] =
            __ldg(reinterpret_cast<const float4*>(src + (int64_t)v * K) + c);
      }
    }
  }
  auto fetch = [&](int64_t b0, int* dst) {  // thread = sample b0 + tid: its H staged-row offsets
    int r[kPairMaxH];
#pragma unroll
    for (int h = 0; h < kPairMaxH; ++h) {
      r[h] = 0;
      if (h < H && b0 + tid < B)
        r[h] = (h * V + min(max(read_state(xT, x_is_float, (int64_t)var[h] * B + b0 + tid), 0), V - 1)) * KT4;
    }
#pragma unroll
    for (int h = 0; h < kPairMaxH; ++h)
      if (h < H) dst[h * kStageRows + tid] = r[h];
  };
  fetch(0, st);
  __syncthreads();
  const int rows_per_pass = 256 / KT4;  // 32 for KT = 32
  const int gl = tid % KT4, rl = tid / KT4;
  float* uf = u + (int64_t)f * B * K + k0;
  int buf = 0;
  for (int64_t b0 = 0; b0 < B; b0 += kStageRows, buf ^= 1) {
    const int* cur = st + buf * kPairMaxH * kStageRows;
    if (b0 + kStageRows < B) fetch(b0 + kStageRows, st + (buf ^ 1) * kPairMaxH * kStageRows);
#pragma unroll 4
    for (int r0 = 0; r0 < kStageRows; r0 += rows_per_pass) {
      const int r = r0 + rl;
      const int64_t b = b0 + r;
      if (b >= B) break;
      float4 acc = reinterpret_cast<const float4*>(tab)[cur[r] + gl];
#pragma unroll
      for (int h = 1; h < kPairMaxH; ++h)
        if (h < H) {
          const float4 z = reinterpret_cast<const float4*>(tab)[cur[h * kStageRows + r] + gl];
          acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
        }
      __stcs(reinterpret_cast<float4*>(uf + b * K) + gl, acc);
    }
    __syncthreads();  // the next chunk's offsets are in place, this chunk's are free
  }
}

int table_pair_gather(const ckb_step_desc_t& d, Ctx& c, float* u) {
  if (d.slot[1] < 0 || d.scope_var == nullptr || d.in_rows == nullptr || d.num_states <= 0 ||
      d.arity > kPairMaxH || (d.k_in & 3) || c.xT == nullptr) {
    set_error("table-input step: needs T2, table variables, fold table, V, arity <= %d, Ki %% 4 == 0 and x",
              kPairMaxH);
    return CKB_ERR_INVALID;
  }
  // shared-memory version: unit tiles of 32 (or all K <= 32 units), the staged slices within 72 KB
  // so that three CTAs share an SM
  const int KT = d.k_in % 32 == 0 ? 32 : (d.k_in <= 32 ? d.k_in : 0);
  const size_t smem = KT ? (size_t)d.arity * d.num_states * KT * 4 + 2 * kPairMaxH * kStageRows * 4 : 0;
  if (KT && 256 % (KT / 4) == 0 && smem <= 72 * 1024 && c.B >= 2 * d.num_states) {
    static PerDeviceOnce attr;
    if (attr.first())
      CKB_CUDA_CHECK(cudaFuncSetAttribute(table_pair_gather_smem_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    dim3 grid(d.k_in / KT, d.num_folds);
    CKB_CUDA_CHECK(launch_pdl(table_pair_gather_smem_kernel, grid, dim3(256), smem, c.stream,
                              (const float*)c.tensors[d.slot[1]], d.scope_var, d.in_rows, c.xT, c.x_is_float, u,
                              c.B, d.k_in, d.num_states, d.arity, KT));
    c.launches++;
    return CKB_OK;
  }
  int G = 1;
  while (G < d.k_in / 4 && G < 32) G <<= 1;
  const int64_t rows_per_block = 8 * (32 / G) * kPairU;
  dim3 grid((int)max64(1, min64(ceil_div(c.B, rows_per_block), 2 * kNumSMs)), d.num_folds);
  table_pair_gather_kernel<<<grid, 256, 0, c.stream>>>(c.tensors[d.slot[1]], d.scope_var, d.in_rows, c.xT,
                                                       c.x_is_float, u, c.B, d.k_in, d.num_states, d.arity);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// Backward of the lookup: dT[f,v,:] = sum over samples with x=v of g[f,b,:]  (the reference gets
// this from autograd as an `index_put_`, 26 % of its CPU step -- SURVEY §3(b)).
//
// One CTA owns a (fold, unit tile, batch split).  It first buckets its samples by state in shared
// memory -- integer counts, an exclusive scan, and a *stable* fill (one warp walks the samples in
// order and ranks equal states inside each group of 32 with match_any), so every bucket lists its
// samples in ascending order -- and then each warp sums the gradient rows of whole buckets in
// registers and writes each table row once.  No floating-point atomics: the result is
// deterministic, and every g row is read exactly once as one contiguous segment.
constexpr int kTableBwdThreads = 1024;  // most warps a CTA of this kernel has (sizes its shared memory)
constexpr int kTableMaxCons = 8;

// THREADS = 1024: one CTA per SM, for chunks of up to 32768 samples; 512: three CTAs per SM, so the
// bucket sort of one overlaps the row sums (the HBM stream) of the others -- the default for the
// chunk sizes of a training batch.  Same bucket lists, same summation order: bit-identical results.
template <int NT, int THREADS>  // NT units per lane: a CTA covers 32*NT units
__global__ void __launch_bounds__(THREADS, THREADS == 1024 ? 1 : 3)
table_bwd_kernel(GradSrc gs, const int32_t* __restrict__ scope_var, const void* __restrict__ xT,
                 int x_is_float, const uint8_t* __restrict__ maskT, int64_t mask_ld,
                 float* __restrict__ out, int64_t B, int K, int V, int64_t chunk) {
  extern __shared__ int smem_i[];
  int* cnt = smem_i;                                     // [V]  bucket sizes, then fill cursors
  int* start = cnt + V;                                  // [V+1] bucket offsets
  uint16_t* xs = reinterpret_cast<uint16_t*>(start + V + 1);  // [chunk] state of every sample
  uint16_t* list = xs + chunk;                           // [chunk] sample ids grouped by state
  __shared__ const float* grows[kTableMaxCons];
  __shared__ int n_cons_s;
  pdl_wait();
  pdl_launch_dependents();
  const int f = blockIdx.y;
  const int k0 = blockIdx.z * (32 * NT);
  const int var = scope_var[f];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int64_t b_begin = (int64_t)blockIdx.x * chunk;
  const int n = (int)(min64(B, b_begin + chunk) - b_begin);

  if (tid == 0) {
    // rows of the gradient arena this fold sums (its consumers), resolved once
    if (gs.cons_ptr == nullptr) {
      n_cons_s = 1;
      grows[0] = gs.garena + (int64_t)f * gs.B * K;
    } else {
      const int c0 = gs.cons_ptr[f];
      const int nc = gs.cons_ptr[f + 1] - c0;
      n_cons_s = nc <= kTableMaxCons ? nc : -1;  // -1: fall back to the generic pull
      for (int c = 0; c < nc && c < kTableMaxCons; ++c) grows[c] = gs.garena + gs.B * gs.cons_rows[c0 + c];
    }
  }
  // Stable bucket sort of the samples by state, every warp on its own contiguous segment:
  //   (1) per-warp histogram wcnt[warp][v] (match_any groups, no atomics),
  //   (2) bucket offsets = exclusive scan over states of the totals; per-warp cursors inside a
  //       bucket = prefix over the warps (lower segments first, so the order is the sample order),
  //   (3) every warp fills its segment through its cursors.
  int* wcnt = reinterpret_cast<int*>(list + chunk);      // [nwarps][V]
  const int seg = ((n + nwarps - 1) / nwarps + 31) & ~31;  // samples per warp, multiple of 32
  const int i_begin = warp * seg, i_end = min(n, i_begin + seg);
  for (int i = tid; i < nwarps * V; i += blockDim.x) wcnt[i] = 0;
  __syncthreads();
  for (int i0 = i_begin; i0 < i_end; i0 += 32) {
    const int i = i0 + lane;
    int v = 0xFFFF;
    if (i < i_end) {
      const int64_t b = b_begin + i;
      if (!read_mask(maskT, mask_ld, var, b))
        v = min(max(read_state(xT, x_is_float, (int64_t)var * B + b), 0), V - 1);
      xs[i] = (uint16_t)v;
    }
    const bool valid = v != 0xFFFF;
    const unsigned same = __match_any_sync(0xffffffffu, valid ? v : 0x10000 + lane);
    if (valid && lane == __ffs(same) - 1) wcnt[warp * V + v] += __popc(same);
    __syncwarp();
  }
  __syncthreads();
  for (int v = tid; v < V; v += blockDim.x) {  // totals; wcnt becomes the prefix over the warps
    int run = 0;
    for (int w = 0; w < nwarps; ++w) {
      const int c = wcnt[w * V + v];
      wcnt[w * V + v] = run;
      run += c;
    }
    cnt[v] = run;
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the bucket sizes
    int carry = 0;
    for (int i0 = 0; i0 < V; i0 += 32) {
      const int i = i0 + lane;
      const int c = i < V ? cnt[i] : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (i < V) start[i] = carry + incl - c;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) start[V] = carry;
  }
  __syncthreads();
  for (int i0 = i_begin; i0 < i_end; i0 += 32) {
    const int i = i0 + lane;
    const int xv = i < i_end ? xs[i] : 0xFFFF;
    const bool valid = xv != 0xFFFF;
    const unsigned same = __match_any_sync(0xffffffffu, valid ? xv : 0x10000 + lane);
    const int rank = __popc(same & ((1u << lane) - 1u));
    int base = 0;
    if (valid) base = start[xv] + wcnt[warp * V + xv];
    __syncwarp();
    if (valid) {
      list[base + rank] = (uint16_t)i;
      if (rank == __popc(same) - 1) wcnt[warp * V + xv] += rank + 1;  // last lane of the group
    }
    __syncwarp();
  }
  __syncthreads();
  // sum whole buckets: warp per state, lanes over units, 8 rows (8*NT loads per lane) in flight
  float* o = out + ((int64_t)blockIdx.x * gridDim.y + f) * V * K;
  const int n_cons = n_cons_s;
  for (int v = warp; v < V; v += nwarps) {
    const int s0 = start[v], s1 = start[v + 1];
    float acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.f;
    if (NT == 2 && K == 64 && n_cons >= 0) {
      // K = 64 (the CTA covers the whole row): a lane owns columns 2*lane, 2*lane + 1 -- one 8-byte
      // load per row and lane, 32-bit offsets inside the fold's block, whole groups of 4 rows
      // without predicates.  Every column is still summed in bucket (= sample) order: same bits
      // as the generic path below, at a seventh of its instructions (it was issue-bound).
      float2 a2 = make_float2(0.f, 0.f);
      for (int c = 0; c < n_cons; ++c) {
        const float2* g0 = reinterpret_cast<const float2*>(grows[c] + b_begin * 64) + lane;
        int j = s0;
        for (; j + 4 <= s1; j += 4) {
          float2 g[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) g[u] = __ldg(g0 + (int)list[j + u] * 32);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            a2.x += g[u].x;
            a2.y += g[u].y;
          }
        }
        float2 g[3];
#pragma unroll
        for (int u = 0; u < 3; ++u)
          g[u] = j + u < s1 ? __ldg(g0 + (int)list[j + u] * 32) : make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (j + u < s1) {
            a2.x += g[u].x;
            a2.y += g[u].y;
          }
      }
      reinterpret_cast<float2*>(o + (int64_t)v * 64)[lane] = a2;
      continue;
    }
    if (n_cons >= 0) {
      for (int c = 0; c < n_cons; ++c) {
        const float* g0 = grows[c] + k0 + lane;
        for (int j0 = s0; j0 < s1; j0 += 8) {
          float g[8][NT];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = j0 + u;
            const bool ok = j < s1;
            const float* row = g0 + (b_begin + (ok ? list[j] : 0)) * K;
#pragma unroll
            for (int t = 0; t < NT; ++t)
              g[u][t] = (ok && k0 + lane + 32 * t < K) ? __ldg(row + 32 * t) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[t] += g[u][t];
        }
      }
    } else {
      for (int j = s0; j < s1; ++j) {
        const int64_t b = b_begin + list[j];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int k = k0 + lane + 32 * t;
          if (k < K) acc[t] += pull_grad(gs, f, b, K, k);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int k = k0 + lane + 32 * t;
      if (k < K) o[(int64_t)v * K + k] = acc[t];
    }
  }
}

// Gradient of the value an integrated variable contributes (logsumexp of unnormalised logits,
// layers/input.py:414-421): dI[f,k] = sum of g[f,b,k] over the samples whose variable is masked.
// One CTA per fold, fixed summation order (threads: 32 units x 8 batch slices).
__global__ void masked_gsum_kernel(GradSrc gs, const int32_t* __restrict__ scope_var,
                                   const uint8_t* __restrict__ maskT, int64_t mask_ld,
                                   float* __restrict__ dint, int64_t B, int K) {
  __shared__ float red[8][33];
  const int f = blockIdx.x;
  const int var = scope_var[f];
  const int lane = threadIdx.x, slice = threadIdx.y;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    float a = 0.f;
    if (k < K)
      for (int64_t b = slice; b < B; b += 8)
        if (read_mask(maskT, mask_ld, var, b)) a += pull_grad(gs, f, b, K, k);
    red[slice][lane] = a;
    __syncthreads();
    if (slice == 0 && k < K) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += red[j][lane];
      dint[(int64_t)f * K + k] = s;
    }
    __syncthreads();
  }
}

static int table_bwd_nt(int K) { return K <= 32 ? 1 : (K <= 64 ? 2 : 4); }

static void table_bwd_config(const ckb_step_desc_t& d, int64_t B, int& splits, int64_t& chunk) {
  // a CTA buckets at most 32768 samples (uint16 states + ids, 128 KB of shared memory)
  splits = ceil_div(B, 32768);
  const int ktiles = ceil_div(d.k_out, 32 * table_bwd_nt(d.k_out));
  const int64_t want = ceil_div(2 * kNumSMs, (int64_t)d.num_folds * ktiles);
  splits = (int)max64(splits, min64(want, ceil_div(B, 1024)));
  chunk = ceil_div(B, splits);
  splits = ceil_div(B, chunk);
}

size_t table_bwd_ws(const ckb_step_desc_t& d, int64_t B) {
  int splits;
  int64_t chunk;
  table_bwd_config(d, B, splits, chunk);
  return splits > 1 ? (size_t)splits * d.num_folds * d.num_states * d.k_out * 4 : 0;
}

int table_bwd(const ckb_step_desc_t& d, Ctx& c) {
  float* dT = c.grads[d.slot[0]];
  if (d.int_slot >= 0 && c.grads[d.int_slot] != nullptr) {
    float* dint = c.grads[d.int_slot];
    if (c.maskT == nullptr) {
      CKB_CUDA_CHECK(cudaMemsetAsync(dint, 0, (size_t)d.num_folds * d.k_out * 4, c.stream));
    } else {
      GradSrc gs{c.garena, d.cons_ptr, d.cons_rows, c.B};
      masked_gsum_kernel<<<d.num_folds, dim3(32, 8), 0, c.stream>>>(gs, d.scope_var, c.maskT,
                                                                   c.mask_ld, dint, c.B, d.k_out);
      CKB_LAUNCH_CHECK();
      c.launches++;
    }
  }
  if (dT == nullptr) return CKB_OK;
  int splits;
  int64_t chunk;
  table_bwd_config(d, c.B, splits, chunk);
  const int V = d.num_states;
  const int threads = chunk <= 8192 ? 512 : kTableBwdThreads;
  const size_t smem = (size_t)(2 * V + 1) * 4 + (size_t)chunk * 4 + (size_t)(threads / 32) * V * 4 + 16;
  if (smem > 200 * 1024) {
    set_error("table_bwd: %d states do not fit shared memory", V);
    return CKB_ERR_UNSUPPORTED;
  }
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    CKB_CUDA_CHECK(cudaFuncSetAttribute(table_bwd_kernel<1, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(table_bwd_kernel<2, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(table_bwd_kernel<4, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(table_bwd_kernel<1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(table_bwd_kernel<2, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKB_CUDA_CHECK(cudaFuncSetAttribute(table_bwd_kernel<4, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const size_t n = (size_t)d.num_folds * V * d.k_out;
  float* out = dT;
  if (splits > 1) {
    if (c.ws_bytes < (size_t)splits * n * 4) {
      set_error("table_bwd: workspace too small");
      return CKB_ERR_WORKSPACE;
    }
    out = (float*)c.ws;
  }
  GradSrc gs{c.garena, d.cons_ptr, d.cons_rows, c.B};
  const int nt = table_bwd_nt(d.k_out);
  dim3 grid(splits, d.num_folds, ceil_div(d.k_out, 32 * nt));
  auto kern = threads == 512
                  ? (nt == 1 ? table_bwd_kernel<1, 512> : (nt == 2 ? table_bwd_kernel<2, 512> : table_bwd_kernel<4, 512>))
                  : (nt == 1 ? table_bwd_kernel<1, 1024> : (nt == 2 ? table_bwd_kernel<2, 1024> : table_bwd_kernel<4, 1024>));
  CKB_CUDA_CHECK(launch_pdl(kern, grid, dim3(threads), smem, c.stream, gs, d.scope_var, c.xT, c.x_is_float,
                            c.maskT, c.mask_ld, out, c.B, d.k_out, V, chunk));
  c.launches++;
  if (splits > 1) return reduce_partials(out, dT, (int64_t)n, splits, c);
  return CKB_OK;
}

// ------------------------------------------------------------------------------------------
// Gaussian log-density, layers/input.py:661-670 (torch.distributions.Normal.log_prob):
//   y = -(x-mu)^2 / (2 sigma^2) - log(sigma) - log(sqrt(2 pi)) (+ log_partition)
// ------------------------------------------------------------------------------------------
__global__ void gaussian_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ stddev,
                                    const float* __restrict__ logp, const int32_t* __restrict__ scope_var,
                                    const void* __restrict__ xT, int x_is_float,
                                    const uint8_t* __restrict__ maskT, int64_t mask_ld,
                                    float* __restrict__ y, int64_t B, int K) {
  const int f = blockIdx.y;
  const int var = scope_var[f];
  float* yf = y + (int64_t)f * B * K;
  const int64_t total = B * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / K;
    const int k = (int)(idx - b * K);
    const float lp = logp ? logp[(int64_t)f * K + k] : 0.f;
    float val;
    if (read_mask(maskT, mask_ld, var, b)) {
      val = lp;  // integrate(): log-partition (0 when normalised), layers/input.py:672-678
    } else {
      const float x = read_value(xT, x_is_float, (int64_t)var * B + b);
      const float mu = mean[(int64_t)f * K + k], sd = stddev[(int64_t)f * K + k];
      const float dlt = x - mu;
      val = -(dlt * dlt) / (2.f * sd * sd) - logf(sd) - 0.91893853320467274178f + lp;
    }
    yf[idx] = val;
  }
}

int gaussian_fwd(const ckb_step_desc_t& d, Ctx& c) {
  const int K = d.k_out;
  const int bx = (int)min64(ceil_div(c.B * K, 256), 4 * kNumSMs);
  dim3 grid(max(bx, 1), d.num_folds);
  gaussian_fwd_kernel<<<grid, 256, 0, c.stream>>>(
      c.tensors[d.slot[0]], c.tensors[d.slot[1]], d.slot[2] >= 0 ? c.tensors[d.slot[2]] : nullptr,
      d.scope_var, c.xT, c.x_is_float, c.maskT, c.mask_ld, c.arena + c.B * d.out_off, c.B, K);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// d/dmu = (x-mu)/sigma^2, d/dsigma = ((x-mu)^2 - sigma^2)/sigma^3, d/dlogp = 1; reduced over
// the batch by one CTA per fold (threads: 32 units x 8 batch slices).
__global__ void gaussian_bwd_kernel(GradSrc gs, const float* __restrict__ mean,
                                    const float* __restrict__ stddev,
                                    const int32_t* __restrict__ scope_var, const void* __restrict__ xT,
                                    int x_is_float, const uint8_t* __restrict__ maskT, int64_t mask_ld,
                                    float* __restrict__ dmean, float* __restrict__ dstd,
                                    float* __restrict__ dlogp, int64_t B, int K) {
  __shared__ float red[3][8][33];
  const int f = blockIdx.x;
  const int var = scope_var[f];
  const int lane = threadIdx.x, slice = threadIdx.y;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    float a_mu = 0.f, a_sd = 0.f, a_lp = 0.f;
    if (k < K) {
      const float mu = mean[(int64_t)f * K + k], sd = stddev[(int64_t)f * K + k];
      const float inv_var = 1.f / (sd * sd);
      for (int64_t b = slice; b < B; b += 8) {
        const float g = pull_grad(gs, f, b, K, k);
        a_lp += g;
        if (!read_mask(maskT, mask_ld, var, b)) {
          const float dlt = read_value(xT, x_is_float, (int64_t)var * B + b) - mu;
          a_mu += g * dlt * inv_var;
          a_sd += g * (dlt * dlt * inv_var - 1.f) / sd;
        }
      }
    }
    red[0][slice][lane] = a_mu;
    red[1][slice][lane] = a_sd;
    red[2][slice][lane] = a_lp;
    __syncthreads();
    if (slice == 0 && k < K) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s0 += red[0][j][lane];
        s1 += red[1][j][lane];
        s2 += red[2][j][lane];
      }
      if (dmean) dmean[(int64_t)f * K + k] = s0;
      if (dstd) dstd[(int64_t)f * K + k] = s1;
      if (dlogp) dlogp[(int64_t)f * K + k] = s2;
    }
    __syncthreads();
  }
}

int gaussian_bwd(const ckb_step_desc_t& d, Ctx& c) {
  float* dmean = c.grads[d.slot[0]];
  float* dstd = c.grads[d.slot[1]];
  float* dlogp = d.slot[2] >= 0 ? c.grads[d.slot[2]] : nullptr;
  if (!dmean && !dstd && !dlogp) return CKB_OK;
  GradSrc gs{c.garena, d.cons_ptr, d.cons_rows, c.B};
  gaussian_bwd_kernel<<<d.num_folds, dim3(32, 8), 0, c.stream>>>(
      gs, c.tensors[d.slot[0]], c.tensors[d.slot[1]], d.scope_var, c.xT, c.x_is_float, c.maskT,
      c.mask_ld, dmean, dstd, dlogp, c.B, d.k_out);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// ------------------------------------------------------------------------------------------
// Constant layer, layers/input.py:739-743: broadcast a (F,K) log-space value over the batch.
// ------------------------------------------------------------------------------------------
__global__ void constant_fwd_kernel(const float* __restrict__ value, float* __restrict__ y,
                                    int64_t B, int K) {
  const int f = blockIdx.y;
  float* yf = y + (int64_t)f * B * K;
  const int64_t total = B * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    yf[idx] = value[(int64_t)f * K + (int)(idx % K)];
}

int constant_fwd(const ckb_step_desc_t& d, Ctx& c) {
  const int bx = (int)min64(ceil_div(c.B * d.k_out, 256), 4 * kNumSMs);
  dim3 grid(max(bx, 1), d.num_folds);
  constant_fwd_kernel<<<grid, 256, 0, c.stream>>>(c.tensors[d.slot[0]], c.arena + c.B * d.out_off,
                                                  c.B, d.k_out);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

__global__ void constant_bwd_kernel(GradSrc gs, float* __restrict__ dvalue, int64_t B, int K) {
  __shared__ float red[8][33];
  const int f = blockIdx.x;
  const int lane = threadIdx.x, slice = threadIdx.y;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    float a = 0.f;
    if (k < K)
      for (int64_t b = slice; b < B; b += 8) a += pull_grad(gs, f, b, K, k);
    red[slice][lane] = a;
    __syncthreads();
    if (slice == 0 && k < K) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += red[j][lane];
      dvalue[(int64_t)f * K + k] = s;
    }
    __syncthreads();
  }
}

int constant_bwd(const ckb_step_desc_t& d, Ctx& c) {
  float* dv = c.grads[d.slot[0]];
  if (!dv) return CKB_OK;
  GradSrc gs{c.garena, d.cons_ptr, d.cons_rows, c.B};
  constant_bwd_kernel<<<d.num_folds, dim3(32, 8), 0, c.stream>>>(gs, dv, c.B, d.k_out);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// ------------------------------------------------------------------------------------------
// EXTERNAL: an input layer the caller evaluated (per-step PyTorch fallback for layer kinds
// without a kernel).  Forward: (F, B, K) activations -> arena block; backward: the gradient of
// every output element, gathered from its consumers, back to the caller.
// ------------------------------------------------------------------------------------------
__global__ void external_bwd_kernel(GradSrc gs, float* __restrict__ gout, int64_t B, int K) {
  const int f = blockIdx.y;
  const int64_t total = B * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / K;
    gout[(int64_t)f * total + idx] = pull_grad(gs, f, b, K, (int)(idx - b * K));
  }
}

int external_fwd(const ckb_step_desc_t& d, Ctx& c) {
  const size_t bytes = (size_t)d.num_folds * c.B * d.k_out * 4;
  CKB_CUDA_CHECK(cudaMemcpyAsync(c.arena + c.B * d.out_off, c.tensors[d.slot[0]], bytes,
                                 cudaMemcpyDeviceToDevice, c.stream));
  c.launches++;
  return CKB_OK;
}

int external_bwd(const ckb_step_desc_t& d, Ctx& c) {
  float* g = c.grads[d.slot[0]];
  if (!g) return CKB_OK;
  GradSrc gs{c.garena, d.cons_ptr, d.cons_rows, c.B};
  const int bx = (int)min64(ceil_div(c.B * d.k_out, 256), 4 * kNumSMs);
  external_bwd_kernel<<<dim3(max(bx, 1), d.num_folds), 256, 0, c.stream>>>(gs, g, c.B, d.k_out);
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                       int64_t n, int splits) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    // fixed left-to-right order (deterministic), but 16 independent loads in flight at a time
    float s = 0.f;
    int j = 0;
    for (; j + 16 <= splits; j += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = __ldg(partial + (int64_t)(j + u) * n + i);
#pragma unroll
      for (int u = 0; u < 16; ++u) s += v[u];
    }
    for (; j < splits; ++j) s += __ldg(partial + (int64_t)j * n + i);
    out[i] = s;
  }
}
// Few outputs, many slabs (the Ko = 1 root layer: 64 weights, 256 slabs): a warp per output, lanes
// over the slabs (each lane sums its slabs in order, then the shuffle tree: a fixed order), so the
// sum is ~8 dependent loads deep instead of 256.
__global__ void reduce_partials_wide_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                            int64_t n, int splits) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += n_warps) {
    float s = 0.f;
    for (int j = lane; j < splits; j += 32) s += __ldg(partial + (int64_t)j * n + i);
    s = warp_sum(s);
    if (lane == 0) out[i] = s;
  }
}
int reduce_partials(const float* partial, float* out, int64_t n, int splits, Ctx& c) {
  if (splits >= 64 && n <= 16384) {
    const int bw = (int)min64(ceil_div(n, 8), 8 * kNumSMs);
    CKB_CUDA_CHECK(launch_pdl(reduce_partials_wide_kernel, dim3(max(bw, 1)), dim3(256), 0, c.stream, partial, out, n, splits));
    c.launches++;
    return CKB_OK;
  }
  const int bx = (int)min64(ceil_div(n, 256), 8 * kNumSMs);
  CKB_CUDA_CHECK(launch_pdl(reduce_partials_kernel, dim3(max(bx, 1)), dim3(256), 0, c.stream, partial, out, n, splits));
  CKB_LAUNCH_CHECK();
  c.launches++;
  return CKB_OK;
}

}  // namespace ckb
