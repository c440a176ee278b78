/*
 * cirkit_b200 -- C ABI of the B200-native (sm_100a) circuit evaluator.
 *
 * The reference (april-tools/cirkit) has no FFI of its own: its "backend" is the Python class
 * TorchCircuit whose forward pass loops over folded layers and calls stock PyTorch ops
 * (cirkit/backend/torch/circuits.py:242-278, graph/modules.py:303-335).  This header declares
 * the entry points a native backend for that path binds instead; INTEGRATION.md shows the
 * ctypes stub on the reference side.  Conventions:
 *
 *   - C linkage, plain pointers and sizes, no torch / C++ types in any signature;
 *   - every function returns 0 on success and a negative ckb_status on failure; the message of
 *     the last failure on the calling thread is available from ckb_last_error();
 *   - no hidden device allocations: activations, gradients, parameters and scratch are
 *     caller-provided device buffers (ckb_plan_workspace_bytes tells how much scratch);
 *   - all device work is enqueued on the caller's stream (a cudaStream_t passed as void*),
 *     nothing synchronises;
 *   - no ownership transfer: the plan copies the small descriptors it is given and keeps the
 *     DEVICE pointers found inside them (index tables), which must outlive the plan;
 *   - thread-compatible: no mutable globals besides the thread-local error string.
 *
 * Memory model (see cirkit_b200/plan.py:PlanLayout).  All offsets are "per sample" float
 * offsets: the block of step s lives at arena + batch * out_off[s] and is laid out
 * (fold, batch, unit), so a plan is independent of the batch size.
 */
#ifndef CIRKIT_B200_H
#define CIRKIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CKB_VERSION 1

typedef enum {
  CKB_OK = 0,
  CKB_ERR_INVALID = -1,     /* bad argument / inconsistent descriptor          */
  CKB_ERR_UNSUPPORTED = -2, /* shape or layer kind without a kernel            */
  CKB_ERR_CUDA = -3,        /* a CUDA runtime call or launch failed            */
  CKB_ERR_WORKSPACE = -4    /* caller-provided workspace too small             */
} ckb_status;

/* Folded layer kinds.  Each restates one reference layer (file:line in cirkit/backend/torch): */
typedef enum {
  CKB_STEP_TABLE = 0,     /* Categorical layers/input.py:399-412 and Embedding :258-266:
                             y[f,b,k] = T[f, x[b,var_f], k], T = log-table (F,V,K)              */
  CKB_STEP_GAUSSIAN = 1,  /* layers/input.py:661-670                                            */
  CKB_STEP_CONSTANT = 2,  /* layers/input.py:739-743 (value already mapped to log space)        */
  CKB_STEP_DENSE = 3,     /* TorchSumLayer layers/inner.py:266-273 (flags=CKB_DENSE_CONCAT) and
                             TorchCPTLayer layers/optimized.py:171-178 (flags=0), both through
                             LSESumSemiring.apply_reduce semiring.py:382-408                    */
  CKB_STEP_MIXING = 4,    /* TorchSumLayer fed by TorchMixingWeightParameter
                             (parameters/nodes.py:847-862), computed on the (F,K,H) weights     */
  CKB_STEP_HADAMARD = 5,  /* layers/inner.py:126-127                                            */
  CKB_STEP_KRONECKER = 6, /* layers/inner.py:178-187 (arity 2)                                  */
  CKB_STEP_TUCKER = 7,    /* layers/optimized.py:89-103 (arity 2)                               */
  CKB_STEP_TENSORDOT = 9, /* TorchTensorDotLayer layers/optimized.py:205-300 (product circuits, e.g. the
                             partition function of a squared circuit): the Ki = Kj*Kq inputs of a
                             sample are Kq interleaved vectors, each contracted with W (F,Kk,Kj):
                             y[q*Kk + k] = lse_j W[k,j] x[j*Kq + q].  num_states carries Kq.
                             'complex-lse-sum' only (CKB_STEP_COMPLEX).                          */
  CKB_STEP_EXTERNAL = 10, /* an input layer evaluated by the CALLER (a layer kind without a kernel here,
                             e.g. Binomial / Polynomial / Evidence: layers/input.py:437, :815, :746):
                             slot[0] holds its (F, B, Ko) output; forward copies it into the arena,
                             backward gathers d(loss)/d(output) into grads[slot[0]] (same shape)   */
  CKB_STEP_TABLE_DENSE = 8 /* a TABLE layer consumed fold-by-fold by an arity-1 DENSE layer, fused:
                             the dense block is applied to the V rows of the (F,V,Ki) table once
                             per step (batch-independent), giving a (F,V,Ko) table T2, and the
                             batch only gathers rows of T2.  Same values as running
                             layers/input.py:399-412 then layers/inner.py:266-273 per sample.
                             slots {T, W, T2}; out_off / consumers are those of the DENSE layer. */
} ckb_step_kind;

#define CKB_DENSE_CONCAT 1 /* reduce over the concatenation of the H inputs instead of their sum */
#define CKB_STEP_ROWS64 2  /* promise: out_off, gin_off and every entry of in_rows / cons_rows are
                              multiples of 64 floats (lets the TMA-fed kernels address the arenas
                              as matrices of 64-float rows; cirkit_b200.plan.build_layout aligns so) */

#define CKB_STEP_COMPLEX 4 /* 'complex-lse-sum' semiring (semiring.py:410-476): the step's arena blocks,
                              gradient blocks and weights hold interleaved (re, im) float pairs;
                              offsets stay in units per sample.  Kinds: TABLE (Embedding
                              layers/input.py:258-266, weights (F,K,V) complex), DENSE without
                              CONCAT, TUCKER (arity 2), HADAMARD, CONSTANT (log-space value)       */
#define CKB_STEP_TABLE_INPUT 16 /* on a DENSE step without CONCAT: its H inputs per fold are rows of
                              the (F', V, Ki) table T2 of a TABLE_DENSE step, selected by the
                              evidence -- slot[1] = T2, scope_var = the variables of the F' table
                              folds, in_rows (F*H) = TABLE FOLD indices (not arena offsets),
                              num_states = V.  Forward: u[f,b,:] = sum_h T2[fold_h, x[b,var_h], :]
                              is gathered once into the arena block at aux_off (the block of the
                              layer that is no longer materialised) and the layer runs on u with
                              arity 1; backward re-reads u.  Half the traffic of gathering the
                              rows and re-reading them (Hadamard of table rows,
                              layers/optimized.py:171-178 over layers/input.py:399-412)           */
#define CKB_STEP_NO_GATHER 32 /* on a TABLE_DENSE step: forward only builds T2; its consumer gathers
                              (CKB_STEP_TABLE_INPUT).  Backward unchanged.                        */
#define CKB_STEP_REAL_TABLE 8 /* with CKB_STEP_COMPLEX on a TABLE step: the table is the REAL (F,V,K)
                              log-table of a Categorical layer, cast to complex (semiring.py:511-514) */

typedef struct {
  int32_t kind;       /* ckb_step_kind                                                        */
  int32_t num_folds;  /* F                                                                    */
  int32_t arity;      /* H                                                                    */
  int32_t k_in;       /* Ki: units of every gathered input row                                */
  int32_t k_out;      /* Ko                                                                   */
  int32_t flags;
  int32_t num_states; /* V (CKB_STEP_TABLE)                                                   */
  int32_t gin_h;      /* input-gradient rows per fold: 1 (shared by all inputs) or H          */
  int64_t out_off;    /* activation arena: per-sample float offset of the (F,B,Ko) block      */
  int64_t gin_off;    /* gradient arena: per-sample float offset of the (F,gin_h,B,Ki) block  */
  const int64_t* in_rows;   /* device (F*H): per-sample float offsets of the gathered rows    */
  const int32_t* scope_var; /* device (F): variable read by every fold (input layers)         */
  const int32_t* cons_ptr;  /* device (F+1): CSR over the gradient rows each output row sums  */
  const int64_t* cons_rows; /* device (nnz): per-sample float offsets into the gradient arena */
  int32_t slot[4];    /* tensor-table slots of the effective parameters, -1 = absent:
                           TABLE: {T}   GAUSSIAN: {mean, stddev, log_partition}
                           CONSTANT: {value}   DENSE/TUCKER: {W (F,Ko,Kred)}   MIXING: {w (F,K,H)} */
  int32_t int_slot;   /* slot of the (F,Ko) values an integrated variable yields, -1 = zeros   */
  int32_t max_consumers; /* largest cons_ptr[f+1]-cons_ptr[f] (lets kernels pick a fast path)   */
  int64_t aux_off;    /* CKB_STEP_TABLE_INPUT: per-sample float offset of the (F,B,Ki) block u   */
} ckb_step_desc_t;

/* Parameter re-parameterisation ops, run before the layers (forward) and after them (backward).
 * They restate cirkit/backend/torch/parameters/nodes.py and fuse the layout changes the
 * kernels want (tables are stored (F,V,K) so a sample reads K contiguous floats). */
typedef enum {
  CKB_POP_SOFTMAX = 0,       /* nodes.py:764-772 over the last axis: (rows, cols)              */
  CKB_POP_LOG_SOFTMAX_T = 1, /* log(softmax(src)) transposed: src (rows=F, aux=K, cols=V) -> (F,V,K) */
  CKB_POP_LOG_T = 2,         /* log(src) transposed, same shapes                               */
  CKB_POP_COPY_T = 3,        /* transpose only                                                 */
  CKB_POP_SCALED_SIGMOID = 4,/* nodes.py:682-699: sigmoid(x)*(b-a)+a, elementwise              */
  CKB_POP_LOG = 5,           /* elementwise log                                                */
  CKB_POP_LSE_ROWS = 6       /* logsumexp over the last axis: (rows, cols) -> (rows): the value an
                                integrated variable of an unnormalised Categorical yields
                                (layers/input.py:414-421); backward ADDS softmax * g to d(src) */
  ,CKB_POP_CONJ = 7          /* complex conjugate of (rows * cols) complex numbers, nodes.py:742-746;
                                backward: conj of the gradient                                  */
} ckb_param_op_kind;

typedef struct {
  int32_t kind; /* ckb_param_op_kind */
  int32_t src;  /* tensor-table slot read  */
  int32_t dst;  /* tensor-table slot written */
  int32_t cols;
  int64_t rows;
  int32_t aux;
  float a, b;
} ckb_param_op_t;

typedef struct ckb_plan ckb_plan_t;

/* element types accepted by ckb_transpose_input */
typedef enum { CKB_U8 = 0, CKB_I32 = 1, CKB_I64 = 2, CKB_F32 = 3, CKB_F64 = 4, CKB_I16 = 5 } ckb_dtype;

int ckb_version(void);
const char* ckb_last_error(void);

/* Build a plan from n_steps step descriptors and n_ops parameter ops over a table of n_slots
 * tensors.  Replaces: TorchCircuit.__init__ building its address book
 * (circuits.py:73-119, graph/modules.py:262-265). */
int ckb_plan_create(const ckb_step_desc_t* steps, int32_t n_steps, const ckb_param_op_t* ops,
                    int32_t n_ops, int32_t n_slots, ckb_plan_t** out);
void ckb_plan_destroy(ckb_plan_t* plan);

/* Scratch bytes ckb_plan_forward / ckb_plan_backward need for a given batch size. */
size_t ckb_plan_workspace_bytes(const ckb_plan_t* plan, int64_t batch);

/* x (batch, num_vars) row-major with leading dimension ld (elements) -> xT (num_vars, batch):
 * int32 for integer dtypes, float32 for floating dtypes.  Replaces the per-layer
 * `x[..., scope_idx].permute(1, 0, 2)` gather of circuits.py:66. */
int ckb_transpose_input(const void* x, int32_t dtype, int64_t batch, int32_t num_vars, int64_t ld,
                        void* xT, void* stream);

/* mask (rows, num_vars) bool bytes -> maskT (num_vars, rows) bytes (IntegrateQuery, queries.py:48-109) */
int ckb_transpose_mask(const uint8_t* mask, int64_t rows, int32_t num_vars, uint8_t* maskT,
                       void* stream);

#define CKB_RUN_PARAM_OPS 1 /* forward: evaluate the parameter ops; backward: back-propagate them */
#define CKB_USE_GRAPHS 2    /* a call whose arguments repeat exactly (pointers, batch, range, flags,
                               table contents) is captured into a CUDA graph the second time it is
                               seen and replayed from then on: one launch instead of ~20 (host time of
                               small-batch loops).  Every external / per-call tensor must keep its
                               address for the replay to be taken; otherwise the call runs eagerly. */

/* One forward pass over steps [step_begin, step_end).  Replaces TorchDiAcyclicGraph.evaluate
 * (graph/modules.py:303-335) + LayerAddressBook.lookup (circuits.py:30-71) + every layer's
 * forward.  tensors: HOST array of n_slots device pointers.  maskT (num_vars, mask_rows) may be
 * NULL; mask_rows is batch, or 1 when a single mask row is broadcast over the batch. */
int ckb_plan_forward(ckb_plan_t* plan, int32_t step_begin, int32_t step_end, int64_t batch,
                     const void* xT, int32_t x_is_float, const uint8_t* maskT, int64_t mask_rows,
                     float* const* tensors, float* arena, void* workspace, size_t workspace_bytes,
                     int32_t flags, void* stream);

/* Reverse pass over the same steps.  Replaces autograd's walk over the reference's op graph
 * (SURVEY §3(c)).  The caller has written d(loss)/d(output) into the output-gradient block of
 * garena.  grads: HOST array of n_slots device pointers (NULL = gradient not wanted); every
 * requested gradient is overwritten, not accumulated. */
int ckb_plan_backward(ckb_plan_t* plan, int32_t step_begin, int32_t step_end, int64_t batch,
                      const void* xT, int32_t x_is_float, const uint8_t* maskT, int64_t mask_rows,
                      float* const* tensors, float* const* grads, const float* arena,
                      float* garena, void* workspace, size_t workspace_bytes, int32_t flags,
                      void* stream);

/* The parameter ops [op_begin, op_end) of the plan alone (the ops CKB_RUN_PARAM_OPS runs as a
 * whole): backward == 0 evaluates them (TorchParameter.forward, parameters/parameter.py:180-188),
 * backward != 0 back-propagates them (grads as for ckb_plan_backward).  Lets a data-parallel
 * caller finish the gradient of a GROUP of parameters -- and start its all-reduce -- while the
 * backward pass of the remaining steps is still running. */
int ckb_plan_param_ops(ckb_plan_t* plan, int32_t op_begin, int32_t op_end, int32_t backward,
                       float* const* tensors, float* const* grads, void* stream);

/* ---- SamplingQuery (cirkit/backend/torch/queries.py:187-275): ancestral sampling, root first.
 * One descriptor per layer, in the plan's (topological) order.  The "selection arena" sel is
 * (num_rows, num_samples) int32 with one row per (layer, fold): the unit a sample's path goes
 * through, -1 when the path does not visit the row.  Stateless: no plan handle involved. */
typedef struct {
  int32_t kind;       /* ckb_step_kind: TABLE (Categorical), GAUSSIAN, DENSE, MIXING, HADAMARD,
                         KRONECKER, TUCKER; anything else is refused (the reference raises
                         TypeError for layers without a sample(), layers/input.py:94-108)      */
  int32_t num_folds, arity, k_in, k_out;
  int32_t num_states; /* TABLE: categories V                                                   */
  int32_t flags;      /* CKB_DENSE_CONCAT as in ckb_step_desc_t                                */
  int32_t sel_row;    /* arena row of fold 0 of this layer                                     */
  const int32_t* in_sel_rows; /* device (F*H): arena rows of the inputs (inner layers)         */
  const int32_t* scope_var;   /* device (F): variable of every fold (input layers)             */
  const float* cdf;   /* device: row-wise inclusive CDFs (ckb_sample_cdf_rows) of the mixture
                         weights -- DENSE/TUCKER (F, Ko, Kred), MIXING (F, K, H) -- or of the
                         category probabilities, TABLE (F, K, V); rows need not sum to 1        */
  const float* p0;    /* GAUSSIAN: mean (F, K)                                                 */
  const float* p1;    /* GAUSSIAN: stddev (F, K)                                               */
} ckb_sample_step_t;

/* dst[r, :] = inclusive prefix sums of row r, left to right in fp32.  mode 0: src (rows, cols)
 * weights.  mode 1: src is a log-probability table laid out (F, V, K) with K = units, dst is
 * (F, K, V): rows = F * K, cols = V (Categorical: layers/input.py:423-434). */
int ckb_sample_cdf_rows(const float* src, float* dst, int64_t rows, int32_t cols, int32_t mode,
                        int32_t units, void* stream);

/* Draw num_samples joint samples: x (num_samples, num_vars) int64 (x_is_float == 0) or float32,
 * zero-initialised by the caller (variables outside the scope stay 0, as the reference's padded
 * samples do, queries.py:258-275).  The path of sample n starts at unit root_unit of arena row
 * root_row; its random numbers are the Philox4x32-10 blocks (counter = (sample_base + n, row),
 * key = seed), so a batch may be drawn in several calls.  mix (same shape as sel, may be NULL)
 * receives the mixture component each visited sum row drew (-1 elsewhere). */
int ckb_plan_sample(const ckb_sample_step_t* steps, int32_t n_steps, int64_t num_samples,
                    int64_t sample_base, uint64_t seed, int64_t num_rows, int32_t root_row,
                    int32_t root_unit, int32_t* sel, int32_t* mix, void* x, int32_t num_vars,
                    int32_t x_is_float, void* stream);

/* Data-parallel gradient sum over NVLink / NVSwitch multicast (the reference has no
 * multi-device path, this is what a user would otherwise get from DistributedDataParallel's
 * all-reduce).  multicast_ptr: the multicast (NVLS) address of a buffer of num_floats floats
 * that every rank has mapped at the same offset of one multicast object; in place, two-shot: rank
 * r sums its 1/world share in the switch (multimem.ld_reduce) and broadcasts it (multimem.st).
 * The caller orders it with a cross-GPU barrier on the stream before and after.  num_ctas <= 0:
 * one CTA per SM. */
int ckb_nvls_allreduce(void* multicast_ptr, int64_t num_floats, int32_t rank, int32_t world,
                           int32_t num_ctas, void* stream);

/* The same sum as ONE kernel, cross-GPU ordering included: waits until every rank has reached
 * the call (flags in the ranks' signal pads: signal_pads_dev is a device array of `world`
 * pointers to peer-mapped uint32 pads of at least 2 KB, zero before the first call), reduces and
 * broadcasts this rank's share, waits until every share has been broadcast and copies the
 * complete buffer (local_ptr: this rank's mapping of it) into out.  epoch: 1, 2, 3, ... per call,
 * the same on every rank; counter: a zero-initialised device uint32.  All ranks must call it. */
int ckb_nvls_allreduce_fused(void* multicast_ptr, const float* local_ptr, float* out,
                             int64_t num_floats, int32_t rank, int32_t world, void* signal_pads_dev,
                             uint32_t epoch, void* counter, int32_t num_ctas, void* stream);

/* Number of kernels the last forward/backward call on this plan enqueued (bench bookkeeping). */
int64_t ckb_plan_last_launches(const ckb_plan_t* plan);

/* Process-wide switches (testing / A-B measurements). */
#define CKB_OPT_TENSOR_CORES 0 /* 1 (default): tcgen05 kernels for the shapes that have one; 0: FP32 SIMT only */
#define CKB_OPT_TC_FAST_MATH 1 /* bit 0: MUFU ex2-based exp (error-compensated), bit 1: MUFU lg2-based log in the tcgen05 kernels;
                                  bit 9 (512): tcgen05 kernels for Ki = Ko = 128 (dense128_tc.cu); default 3 | 512 */
int ckb_set_option(int32_t option, int32_t value);

/* Copies the device-side debug timeline (clock64 stamps of the tcgen05 kernels) to host memory. */
int ckb_debug_read(void* dst, size_t bytes);

/* ---- 'complex-lse-sum' per-layer building blocks (validated on the B200 in round 2,
 * tests/test_gpu_zzz_complex_kernels.py; the plan executor runs complex plans through its own step
 * kernels, csrc/complex_plan.cu, and does not call these).  Complex tensors are interleaved
 * (re, im) float pairs in the (fold, batch, unit) layout.  Replaces, per layer:
 * ComplexLSESumSemiring.apply_reduce semiring.py:440-476 under TorchCPTLayer.forward
 * layers/optimized.py:171-178 (x1 == NULL: an arity-1 TorchSumLayer, layers/inner.py:266-273),
 * csafelog utils.py:32-50, TorchEmbeddingLayer.forward layers/input.py:258-266.
 *   cpt:  x0, x1 (F,B,Ki)  w (F,Ko,Ki)  y, gy (F,B,Ko)  gu (F,B,Ki) = gradient of x0 + x1
 *         gw (F,Ko,Ki) ACCUMULATED (caller zeroes; may be NULL)
 *   embedding:  x (B, ld) int64 evidence, var (F,) device int32 column per fold, w (F,K,V),
 *         y, gy (F,B,K), gw (F,K,V) ACCUMULATED */
int ckb_complex_cpt_fwd(const float* x0, const float* x1, const float* w, float* y, int32_t F,
                        int64_t B, int32_t Ki, int32_t Ko, void* stream);
int ckb_complex_cpt_bwd(const float* x0, const float* x1, const float* w, const float* y,
                        const float* gy, float* gu, float* gw, int32_t F, int64_t B, int32_t Ki,
                        int32_t Ko, void* stream);
int ckb_complex_embedding_fwd(const int64_t* x, int64_t ld, const int32_t* var, const float* w,
                              float* y, int32_t F, int64_t B, int32_t K, int32_t V, void* stream);
int ckb_complex_embedding_bwd(const int64_t* x, int64_t ld, const int32_t* var, const float* w,
                              const float* gy, float* gw, int32_t F, int64_t B, int32_t K,
                              int32_t V, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CIRKIT_B200_H */
