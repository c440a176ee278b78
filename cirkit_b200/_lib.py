"""ctypes binding of include/cirkit_b200.h.

The CUDA library is the product: there is no CPU or PyTorch fallback behind these calls.  If
`libcirkit_b200.so` is missing or fails to load, every use raises :class:`LibraryNotBuiltError`.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcirkit_b200.so")

# ckb_step_kind / ckb_param_op_kind / ckb_dtype (keep in sync with include/cirkit_b200.h)
STEP_TABLE, STEP_GAUSSIAN, STEP_CONSTANT, STEP_DENSE, STEP_MIXING, STEP_HADAMARD, STEP_KRONECKER, STEP_TUCKER = range(8)
STEP_TENSORDOT = 9
STEP_EXTERNAL = 10
DENSE_CONCAT = 1
STEP_ROWS64 = 2
STEP_COMPLEX = 4
STEP_REAL_TABLE = 8
STEP_TABLE_INPUT = 16
STEP_NO_GATHER = 32
POP_SOFTMAX, POP_LOG_SOFTMAX_T, POP_LOG_T, POP_COPY_T, POP_SCALED_SIGMOID, POP_LOG, POP_LSE_ROWS, POP_CONJ = range(8)
U8, I32, I64, F32, F64, I16 = range(6)
RUN_PARAM_OPS = 1
USE_GRAPHS = 2

EXPORTS = (
    "ckb_version",
    "ckb_last_error",
    "ckb_plan_create",
    "ckb_plan_destroy",
    "ckb_plan_workspace_bytes",
    "ckb_transpose_input",
    "ckb_transpose_mask",
    "ckb_plan_forward",
    "ckb_plan_backward",
    "ckb_plan_param_ops",
    "ckb_plan_last_launches",
    "ckb_sample_cdf_rows",
    "ckb_plan_sample",
    "ckb_nvls_allreduce",
    "ckb_nvls_allreduce_fused",
    "ckb_set_option",
    "ckb_debug_read",
    # experimental complex-semiring building blocks (not used by the plan executor)
    "ckb_complex_cpt_fwd",
    "ckb_complex_cpt_bwd",
    "ckb_complex_embedding_fwd",
    "ckb_complex_embedding_bwd",
)
OPT_TENSOR_CORES = 0


class LibraryNotBuiltError(RuntimeError):
    pass


class CkbError(RuntimeError):
    pass


class StepDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("num_folds", C.c_int32),
        ("arity", C.c_int32),
        ("k_in", C.c_int32),
        ("k_out", C.c_int32),
        ("flags", C.c_int32),
        ("num_states", C.c_int32),
        ("gin_h", C.c_int32),
        ("out_off", C.c_int64),
        ("gin_off", C.c_int64),
        ("in_rows", C.c_void_p),
        ("scope_var", C.c_void_p),
        ("cons_ptr", C.c_void_p),
        ("cons_rows", C.c_void_p),
        ("slot", C.c_int32 * 4),
        ("int_slot", C.c_int32),
        ("max_consumers", C.c_int32),
        ("aux_off", C.c_int64),
    ]


class SampleStep(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("num_folds", C.c_int32),
        ("arity", C.c_int32),
        ("k_in", C.c_int32),
        ("k_out", C.c_int32),
        ("num_states", C.c_int32),
        ("flags", C.c_int32),
        ("sel_row", C.c_int32),
        ("in_sel_rows", C.c_void_p),
        ("scope_var", C.c_void_p),
        ("cdf", C.c_void_p),
        ("p0", C.c_void_p),
        ("p1", C.c_void_p),
    ]


class ParamOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("src", C.c_int32),
        ("dst", C.c_int32),
        ("cols", C.c_int32),
        ("rows", C.c_int64),
        ("aux", C.c_int32),
        ("a", C.c_float),
        ("b", C.c_float),
    ]


_lib = None


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryNotBuiltError(
            f"{LIB_PATH} not found: build it with `python -m cirkit_b200.build` "
            "(nvcc, sm_100a).  cirkit_b200 has no CPU fallback."
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise LibraryNotBuiltError(f"cannot load {LIB_PATH}: {exc}") from exc
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.ckb_version.restype = C.c_int
    lib.ckb_last_error.restype = C.c_char_p
    lib.ckb_plan_create.argtypes = [C.POINTER(StepDesc), i32, C.POINTER(ParamOp), i32, i32, C.POINTER(vp)]
    lib.ckb_plan_create.restype = C.c_int
    lib.ckb_plan_destroy.argtypes = [vp]
    lib.ckb_plan_destroy.restype = None
    lib.ckb_plan_workspace_bytes.argtypes = [vp, i64]
    lib.ckb_plan_workspace_bytes.restype = C.c_size_t
    lib.ckb_transpose_input.argtypes = [vp, i32, i64, i32, i64, vp, vp]
    lib.ckb_transpose_input.restype = C.c_int
    lib.ckb_transpose_mask.argtypes = [vp, i64, i32, vp, vp]
    lib.ckb_transpose_mask.restype = C.c_int
    lib.ckb_plan_forward.argtypes = [vp, i32, i32, i64, vp, i32, vp, i64, C.POINTER(vp), vp, vp, C.c_size_t, i32, vp]
    lib.ckb_plan_forward.restype = C.c_int
    lib.ckb_plan_backward.argtypes = [vp, i32, i32, i64, vp, i32, vp, i64, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, C.c_size_t, i32, vp]
    lib.ckb_plan_backward.restype = C.c_int
    lib.ckb_plan_param_ops.argtypes = [vp, i32, i32, i32, C.POINTER(vp), C.POINTER(vp), vp]
    lib.ckb_plan_param_ops.restype = C.c_int
    lib.ckb_plan_last_launches.argtypes = [vp]
    lib.ckb_plan_last_launches.restype = i64
    lib.ckb_sample_cdf_rows.argtypes = [vp, vp, i64, i32, i32, i32, vp]
    lib.ckb_sample_cdf_rows.restype = C.c_int
    lib.ckb_plan_sample.argtypes = [C.POINTER(SampleStep), i32, i64, i64, C.c_uint64, i64, i32, i32, vp, vp, vp, i32, i32, vp]
    lib.ckb_plan_sample.restype = C.c_int
    lib.ckb_nvls_allreduce.argtypes = [vp, i64, i32, i32, i32, vp]
    lib.ckb_nvls_allreduce.restype = C.c_int
    lib.ckb_nvls_allreduce_fused.argtypes = [vp, vp, vp, i64, i32, i32, vp, C.c_uint32, vp, i32, vp]
    lib.ckb_nvls_allreduce_fused.restype = C.c_int
    lib.ckb_set_option.argtypes = [i32, i32]
    lib.ckb_set_option.restype = C.c_int
    lib.ckb_debug_read.argtypes = [vp, C.c_size_t]
    lib.ckb_debug_read.restype = C.c_int
    lib.ckb_complex_cpt_fwd.argtypes = [vp, vp, vp, vp, i32, i64, i32, i32, vp]
    lib.ckb_complex_cpt_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, vp]
    lib.ckb_complex_embedding_fwd.argtypes = [vp, i64, vp, vp, vp, i32, i64, i32, i32, vp]
    lib.ckb_complex_embedding_bwd.argtypes = [vp, i64, vp, vp, vp, vp, i32, i64, i32, i32, vp]
    for fn in (lib.ckb_complex_cpt_fwd, lib.ckb_complex_cpt_bwd, lib.ckb_complex_embedding_fwd,
               lib.ckb_complex_embedding_bwd):
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ckb_last_error().decode(errors="replace")
        raise CkbError(f"{what} failed ({rc}): {msg}")
