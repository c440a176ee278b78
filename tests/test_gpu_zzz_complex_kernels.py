"""Layer-by-layer check of the complex-semiring building blocks (csrc/complex_kernels.cu, the
`ckb_complex_*` entry points) against the oracle's complex path on the `*_complex*` fixtures."""
import os

import pytest
import torch

from helpers import Golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _close(got, want, what, tol=2e-4):
    """Complex logs agree when exp(got - want) = 1 (the imaginary part is a phase)."""
    got, want = got.cpu().to(torch.complex128), want.detach().to(torch.complex128)
    finite = torch.isfinite(want.real)
    assert bool((torch.isfinite(got.real) == finite).all()), what
    err = (torch.exp(got[finite] - want[finite]) - 1).abs().max().item() if finite.any() else 0.0
    assert err <= tol, f"{what}: {err:.3e}"


def _close_lin(got, want, what, tol=2e-4):
    got, want = got.cpu().to(torch.complex128), want.detach().to(torch.complex128)
    scale = max(want.abs().max().item(), 1e-30)
    err = (got - want).abs().max().item() / scale
    assert err <= tol, f"{what}: {err:.3e}"


@pytest.mark.parametrize("name", ["rbt16_cpt_k4_complex", "rbt16_cpt_k4_complex_conj"])
def test_complex_layers_vs_oracle(name, dev):
    from cirkit_b200 import _lib
    from oracle import OracleCircuit
    from oracle.reference_eval import _ComplexSafeLog, complex_lse_apply_reduce

    lib = _lib.load()
    g = Golden(name)
    plan = g.plan
    oc = OracleCircuit(plan, dtype=torch.float32)
    with torch.no_grad():
        for p, v in zip(oc.leaves, g.leaves(torch.float32)):
            p.copy_(v)
    x = g.x()
    B = x.shape[0]
    y_ref = oc(x)
    outs = oc.last_outputs
    for t in outs:
        t.retain_grad()
    (-y_ref.real.mean()).backward()
    stream = torch.cuda.current_stream(dev).cuda_stream
    xd = x.to(dev)

    for sid, s in enumerate(plan.steps):
        F, Ki, Ko = s.num_folds, s.num_input_units, s.num_output_units
        gy = outs[sid].grad.to(torch.complex64).contiguous().to(dev)
        if s.kind == "embedding":
            w = oc.param(s.params["weight"]).detach().to(torch.complex64).contiguous()
            V = w.shape[2]
            var = torch.as_tensor(s.scope_idx, dtype=torch.int32, device=dev)
            wd = w.to(dev)
            y = torch.empty(F, B, Ko, dtype=torch.complex64, device=dev)
            _lib.check(lib.ckb_complex_embedding_fwd(xd.data_ptr(), xd.stride(0), var.data_ptr(),
                                                     wd.data_ptr(), y.data_ptr(), F, B, Ko, V, stream),
                       "ckb_complex_embedding_fwd")
            _close(y, outs[sid], f"step {sid} embedding forward")
            gw = torch.zeros_like(wd)
            _lib.check(lib.ckb_complex_embedding_bwd(xd.data_ptr(), xd.stride(0), var.data_ptr(),
                                                     wd.data_ptr(), gy.data_ptr(), gw.data_ptr(),
                                                     F, B, Ko, V, stream), "ckb_complex_embedding_bwd")
            w_ = w.clone().requires_grad_()
            xs = x[:, torch.as_tensor(s.scope_idx, dtype=torch.int64)].t()  # (F, B)
            y_ = _ComplexSafeLog.apply(w_[torch.arange(F)[:, None], :, xs])
            y_.backward(outs[sid].grad.to(torch.complex64))
            _close_lin(gw, w_.grad, f"step {sid} embedding weight gradient")
            continue
        assert s.kind == "cpt" and s.arity == 2
        ins = [torch.stack([outs[int(s.in_step[f, h])][int(s.in_fold[f, h])] for f in range(F)])
               for h in range(2)]
        x0, x1 = (t.detach().to(torch.complex64).contiguous().to(dev) for t in ins)
        w = oc.param(s.params["weight"]).detach().to(torch.complex64).contiguous()
        wd = w.to(dev)
        y = torch.empty(F, B, Ko, dtype=torch.complex64, device=dev)
        _lib.check(lib.ckb_complex_cpt_fwd(x0.data_ptr(), x1.data_ptr(), wd.data_ptr(), y.data_ptr(),
                                           F, B, Ki, Ko, stream), "ckb_complex_cpt_fwd")
        _close(y, outs[sid], f"step {sid} cpt forward")
        gu = torch.empty(F, B, Ki, dtype=torch.complex64, device=dev)
        gw = torch.zeros_like(wd)
        _lib.check(lib.ckb_complex_cpt_bwd(x0.data_ptr(), x1.data_ptr(), wd.data_ptr(), y.data_ptr(),
                                           gy.data_ptr(), gu.data_ptr(), gw.data_ptr(), F, B, Ki, Ko,
                                           stream), "ckb_complex_cpt_bwd")
        u_ = (ins[0] + ins[1]).detach().to(torch.complex64).requires_grad_()
        w_ = w.clone().requires_grad_()
        y_ = complex_lse_apply_reduce(lambda e: torch.einsum("fbi,foi->fbo", e, w_), u_)
        y_.backward(outs[sid].grad.to(torch.complex64))
        _close_lin(gu, u_.grad, f"step {sid} cpt input gradient")
        _close_lin(gw, w_.grad, f"step {sid} cpt weight gradient")
    torch.cuda.synchronize(dev)
