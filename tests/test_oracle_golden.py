"""The CPU oracle against the fixtures generated from the real reference (tests/golden/)."""
import math

import pytest
import torch

from helpers import Golden, golden_names
from oracle import OracleCircuit

FULL = golden_names("full")
SEEDED = golden_names("seeded")
FP64_RTOL, FP64_ATOL = 1e-8, 1e-12  # the reference's own tolerances, tests/floats.py:5-6


def _oracle(g: Golden, dtype=torch.float64) -> OracleCircuit:
    oc = OracleCircuit(g.plan, dtype=dtype)
    with torch.no_grad():
        for p, v in zip(oc.leaves, g.leaves(dtype)):
            p.copy_(v)
    return oc


@pytest.mark.parametrize("name", FULL)
def test_forward_and_gradients_match_reference(name):
    g = Golden(name)
    oc = _oracle(g)
    y = oc(g.x())
    assert y.shape == g.y().shape
    torch.testing.assert_close(y, g.y(), rtol=FP64_RTOL, atol=FP64_ATOL)
    (-y.mean()).backward()
    for p, gr in zip(oc.leaves, g.grads()):
        got = torch.zeros_like(p) if p.grad is None else p.grad
        torch.testing.assert_close(got, gr, rtol=1e-7, atol=1e-13)


@pytest.mark.parametrize("name", [n for n in FULL if Golden(n).mask()[0] is not None])
def test_integrate_query_matches_reference(name):
    g = Golden(name)
    mask, y_mask = g.mask()
    oc = _oracle(g)
    with torch.no_grad():
        y = oc(g.x(), integrate_mask=mask)
    torch.testing.assert_close(y, y_mask, rtol=FP64_RTOL, atol=FP64_ATOL)


def test_known_answers_categorical():
    """Hand-computed values of the reference test-suite, tests/symbolic/test_utils.py:411-417."""
    g = Golden("ka_categorical_cpt")
    oc = _oracle(g)
    with torch.no_grad():
        y = oc(g.x()).reshape(-1)
        for bits, val in g.meta["evi"].items():
            assert math.isclose(y[int(bits, 2)].exp().item(), val, rel_tol=1e-8)
        # sum over all 2^5 worlds = partition function (test_compile_circuit.py:47-50)
        assert math.isclose(torch.logsumexp(y, 0).exp().item(), g.meta["Z"], rel_tol=1e-8)
        mask, _ = g.mask()
        ym = oc(g.x(), integrate_mask=mask).reshape(-1)
        for bits, val in g.meta["mar"].items():
            assert math.isclose(ym[int(bits, 2)].exp().item(), val, rel_tol=1e-8)
    gz = Golden("ka_categorical_cpt_Z")
    oz = _oracle(gz)
    with torch.no_grad():
        z = oz()
    assert z.shape == gz.y().shape
    assert math.isclose(z.exp().item(), 318.0, rel_tol=1e-8)


def test_known_answers_gaussian():
    """tests/symbolic/test_utils.py:497-503."""
    g = Golden("ka_gaussian")
    oc = _oracle(g)
    with torch.no_grad():
        y = oc(g.x()).reshape(-1)
        assert math.isclose(y[0].exp().item(), 3.744904862456293, rel_tol=1e-8)
        mask, _ = g.mask()
        ym = oc(g.x(), integrate_mask=mask).reshape(-1)
        assert math.isclose(ym[1].exp().item(), 23.528960785605985, rel_tol=1e-8)
        assert math.isclose(ym[3].exp().item(), 44.0, rel_tol=1e-8)  # all variables integrated: Z


@pytest.mark.parametrize("name", [n for n in SEEDED if "tucker" not in n])
def test_seeded_benchmark_circuits(name):
    """Benchmark-size circuits: leaves re-drawn from the seed, outputs vs the reference's."""
    g = Golden(name)
    oc = _oracle(g)
    x = g.x()[:2]
    with torch.no_grad():
        y = oc(x)
    torch.testing.assert_close(y, g.y()[:2], rtol=FP64_RTOL, atol=1e-9)


def test_error_behaviour():
    g = Golden("qt8_cp_k4")
    oc = _oracle(g)
    with pytest.raises(ValueError, match="Expected some input"):
        oc()
    with pytest.raises(ValueError, match="shape"):
        oc(torch.zeros(3, dtype=torch.int64))
