"""The CPU oracle against the fixtures generated from the real reference (tests/golden/)."""
import math

import pytest
import torch

from helpers import Golden, golden_names
from oracle import OracleCircuit

FULL = golden_names("full")
SEEDED = golden_names("seeded")
COMPLEX = golden_names("complex")
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


@pytest.mark.parametrize("name", COMPLEX)
def test_complex_semiring_matches_reference(name):
    """'complex-lse-sum' circuits (semiring.py:440-476, csafelog utils.py:32-50):
    complex128 outputs and the gradients of -mean(Re y) against the reference's."""
    g = Golden(name)
    assert g.plan.semiring == "complex-lse-sum"
    oc = _oracle(g)
    y = oc(g.x())
    assert y.is_complex() and y.shape == g.y().shape
    torch.testing.assert_close(y, g.y(), rtol=FP64_RTOL, atol=FP64_ATOL)
    (-y.real.mean()).backward()
    for p, gr in zip(oc.leaves, g.grads()):
        got = torch.zeros_like(p) if p.grad is None else p.grad
        assert got.dtype == gr.dtype
        torch.testing.assert_close(got, gr, rtol=1e-7, atol=1e-13)
    # fp32 leaves -> complex64 activations, the dtype a CUDA path would compute in
    oc32 = _oracle(g, torch.float32)
    with torch.no_grad():
        y32 = oc32(g.x())
    assert y32.dtype == torch.complex64
    # compare through exp: the imaginary part is a phase (test_compile_circuit_operators.py:236-238)
    torch.testing.assert_close(torch.exp(y32 - g.y().real.to(torch.complex64)),
                               torch.exp(g.y() - g.y().real).to(torch.complex64), rtol=2e-4, atol=2e-4)


def test_complex_safe_log_gradient_is_finite_at_zero():
    """tests/backend/torch/test_semiring.py:29-38: the gradient of the safe complex log at 0."""
    from oracle.reference_eval import _ComplexSafeLog

    z = torch.zeros(5, dtype=torch.complex64)
    z.real[2] = z.imag[2] = 1.0
    z.requires_grad = True
    _ComplexSafeLog.apply(z).mean().real.backward()
    assert torch.all(torch.isfinite(torch.view_as_real(z.grad)))
    assert z.grad[2] == (0.2 / torch.tensor(1 - 1j)).to(torch.complex64)


def test_complex_lse_agrees_with_real_lse_on_tiny_weights():
    """tests/backend/torch/test_semiring.py:41-61: x = (-200, -200, -5), w = (1, 2, 1e-38)."""
    from oracle.reference_eval import complex_lse_apply_reduce, lse_apply_reduce

    x = torch.tensor([[-200.0, -200.0, -5.0]])
    w = torch.tensor([[1.0], [2.0], [1e-38]])
    y1 = lse_apply_reduce(lambda e: torch.einsum("kj,ji->ki", e, w), x)
    y2 = complex_lse_apply_reduce(
        lambda e: torch.einsum("kj,ji->ki", e, w.to(torch.complex64)), x.to(torch.complex64))
    assert torch.isfinite(y1).all() and torch.isfinite(torch.view_as_real(y2)).all()
    torch.testing.assert_close(y1, y2.real)


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


def test_complex_backward_formulas_of_the_kernels():
    """The closed-form backward csrc/complex_kernels.cu implements -- r = gy / conj(S) with
    S = exp(y - m), g_e = r conj(W), g_u = g_e conj(e), g_W = r conj(e), the shift m held constant;
    Embedding: g_W[f,k,x] += gy / conj(W[f,k,x]) -- equals autograd through the oracle's complex
    path (PyTorch's convention for complex gradients), layer by layer, in float64."""
    from oracle.reference_eval import _ComplexSafeLog, complex_lse_apply_reduce

    g = Golden("rbt16_cpt_k4_complex")
    plan = g.plan
    oc = _oracle(g)
    x = g.x()
    y = oc(x)
    outs = oc.last_outputs
    for t in outs:
        t.retain_grad()
    (-y.real.mean()).backward()
    for sid, s in enumerate(plan.steps):
        F, gy = s.num_folds, outs[sid].grad
        w = oc.param(s.params["weight"]).detach()
        w_ = w.clone().requires_grad_()
        if s.kind == "embedding":
            xs = x[:, torch.as_tensor(s.scope_idx, dtype=torch.int64)].t()  # (F, B)
            _ComplexSafeLog.apply(w_[torch.arange(F)[:, None], :, xs]).backward(gy)
            gw = torch.zeros_like(w)
            for f in range(F):
                for b in range(x.shape[0]):
                    gw[f, :, xs[f, b]] += gy[f, b, :] / w[f, :, xs[f, b]].conj()
            torch.testing.assert_close(gw, w_.grad, rtol=1e-12, atol=1e-14)
            continue
        ins = [torch.stack([outs[int(s.in_step[f, h])][int(s.in_fold[f, h])] for f in range(F)])
               for h in range(s.arity)]
        u = sum(ins).detach()
        u_ = u.clone().requires_grad_()
        y_ = complex_lse_apply_reduce(lambda e: torch.einsum("fbi,foi->fbo", e, w_), u_)
        y_.backward(gy)
        m = u.real.amax(-1, keepdim=True)
        e, S = torch.exp(u - m), torch.exp(y_.detach() - m)
        r = gy / S.conj()
        gu = torch.einsum("fbo,foi->fbi", r, w.conj()) * e.conj()
        gw = torch.einsum("fbo,fbi->foi", r, e.conj())
        torch.testing.assert_close(gu, u_.grad, rtol=1e-10, atol=1e-14)
        torch.testing.assert_close(gw, w_.grad, rtol=1e-10, atol=1e-14)
