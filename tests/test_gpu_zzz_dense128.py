"""EXPERIMENTAL tcgen05 forward for Ki = Ko = 128 (csrc/dense128_tc.cu), opt-in through bit 9 of
CKB_OPT_TC_FAST_MATH.  Written after the last GPU call of round 1, never run; by default K = 128
layers take the FP32 SIMT kernels, which is what the regular tests cover.  Runs only on request:

    CKB_EXPERIMENTAL=1 python -m pytest tests/test_gpu_zzz_dense128.py -m gpu -q
"""
import dataclasses
import os

import pytest
import torch

from helpers import Golden

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(os.environ.get("CKB_EXPERIMENTAL") != "1",
                       reason="experimental kernel: set CKB_EXPERIMENTAL=1"),
]
OPT_TC_FAST_MATH = 1


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name,batch", [("qt8_cp_k4", 1), ("qt8_cp_k4", 128), ("qt8_cp_k4", 300),
                                        ("qg8_cp_k4_densemix", 257), ("pd6_cp_k3_unopt", 45)])
def test_dense128_forward_vs_simt_and_oracle(name, batch, dev):
    from cirkit_b200 import B200Circuit, _lib
    from cirkit_b200.plan import seeded_leaves
    from oracle import OracleCircuit
    from oracle.reference_eval import make_inputs

    g = Golden(name)
    k0 = g.plan.steps[0].num_output_units
    plan = dataclasses.replace(g.plan, meta={"units": k0}).with_units(128)
    assert any(s.kind in ("cpt", "sum") and s.num_input_units == 128 and s.num_output_units == 128
               for s in plan.steps)
    cc = B200Circuit(plan, seed=5).to(dev)
    oc = OracleCircuit(plan, dtype=torch.float64)
    with torch.no_grad():
        for q, v in zip(oc.leaves, seeded_leaves(plan, 5)):
            q.copy_(v)
    x = make_inputs(plan, batch, seed=batch)
    with torch.no_grad():
        yo = oc(x)
    lib = _lib.load()
    res = {}
    try:
        for bits in (3 | 512, 3):
            assert lib.ckb_set_option(OPT_TC_FAST_MATH, bits) == 0
            with torch.no_grad():
                res[bits] = cc(x.to(dev)).double().cpu()
    finally:
        lib.ckb_set_option(OPT_TC_FAST_MATH, 3)
    tol = 5e-7 * yo.abs().max().item() + 1e-5
    assert torch.isfinite(res[3 | 512]).all()
    assert (res[3] - yo).abs().max().item() <= tol  # the SIMT route (sanity)
    err = (res[3 | 512] - yo).abs().max().item()
    assert err <= tol, f"tcgen05 K=128 forward vs oracle: {err:.3e} > {tol:.3e}"
