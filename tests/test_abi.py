"""C-ABI hygiene (no GPU): the library builds, loads, exports exactly what the header declares,
and the host refuses to run without it."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "cirkit_b200.h")


@pytest.fixture(scope="module")
def lib():
    from cirkit_b200 import build, _lib

    build.build()
    return _lib.load()


def _declared() -> list[str]:
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ckb_[a-z_]+)\s*\(", src)))


def test_header_and_exports_agree(lib):
    from cirkit_b200 import _lib

    declared = _declared()
    assert declared == sorted(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.ckb_version() == 1


def test_struct_layouts_match_header(lib):
    from cirkit_b200 import _lib

    # sizes the C compiler gives the two descriptor structs (natural alignment, LP64)
    assert ctypes.sizeof(_lib.StepDesc) == 112
    assert ctypes.sizeof(_lib.ParamOp) == 40
    assert _lib.StepDesc.out_off.offset == 32 and _lib.StepDesc.in_rows.offset == 48
    assert _lib.StepDesc.slot.offset == 80 and _lib.StepDesc.int_slot.offset == 96
    assert _lib.StepDesc.max_consumers.offset == 100 and _lib.StepDesc.aux_off.offset == 104
    assert ctypes.sizeof(_lib.SampleStep) == 72 and _lib.SampleStep.in_sel_rows.offset == 32


def test_argument_errors_are_reported_without_a_gpu(lib):
    handle = ctypes.c_void_p()
    rc = lib.ckb_plan_create(None, 0, None, 0, 0, ctypes.byref(handle))
    assert rc == -1
    assert b"bad arguments" in lib.ckb_last_error()
    assert lib.ckb_plan_workspace_bytes(None, 128) >= 256


def test_experimental_complex_entry_points_validate_arguments(lib):
    """No GPU: the complex building blocks reject null pointers and bad shapes before launching."""
    assert lib.ckb_complex_cpt_fwd(None, None, None, None, 1, 1, 4, 4, None) == -1
    assert b"null pointer" in lib.ckb_last_error()
    buf = (ctypes.c_float * 8)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.ckb_complex_cpt_fwd(p, None, p, p, 0, 1, 4, 4, None) == -1
    assert b"bad shape" in lib.ckb_last_error()
    assert lib.ckb_complex_cpt_bwd(p, None, p, p, p, None, None, 1, 1, 4, 4, None) == -1
    assert lib.ckb_complex_embedding_fwd(None, 0, None, None, None, 1, 1, 4, 4, None) == -1
    assert lib.ckb_complex_embedding_bwd(p, 1, p, p, p, None, 1, 1, 4, 4, None) == -1
    # an empty batch is a no-op, not an error
    assert lib.ckb_complex_cpt_fwd(p, None, p, p, 1, 0, 4, 4, None) == 0


def test_missing_library_fails_loudly(monkeypatch):
    from cirkit_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcirkit_b200.so")
    with pytest.raises(_lib.LibraryNotBuiltError, match="no CPU fallback"):
        _lib.load()


def test_cpu_parameters_are_rejected():
    from cirkit_b200 import B200Circuit
    from helpers import Golden

    g = Golden("gmm1d_k8")
    cc = B200Circuit(g.plan)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cc(torch.zeros(4, 1))


def test_product_does_not_import_the_oracle():
    import subprocess
    import sys

    code = (
        "import sys; sys.path.insert(0, %r); import cirkit_b200, cirkit_b200.runtime, "
        "cirkit_b200.circuit, cirkit_b200.queries, cirkit_b200.adapter, cirkit_b200.plan; "
        "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    ) % REPO
    subprocess.check_call([sys.executable, "-c", code])
    pkg = os.path.join(REPO, "cirkit_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
