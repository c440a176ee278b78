"""SamplingQuery on the GPU (csrc/sampling_kernels.cu through ckb_plan_sample) against
  (1) the numpy restatement of the same top-down sampler and random stream (oracle/sampling.py):
      sample-by-sample equality, up to the rare draw that lands within an ulp of a CDF boundary
      (the device builds its CDFs from its own fp32 softmax);
  (2) the reference's exact joint distribution (tests/golden/smp_*.npz), with the chi-square test
      of tests/test_sampling_oracle.py and the reference's own criterion (ratios within 3e-2 of
      the probabilities for 10^6 samples, tests/backend/torch/test_queries/test_sampling.py:53)."""
import numpy as np
import pytest
import torch

from helpers import Golden, golden_names
from test_sampling_oracle import chi_square_ok, oracle_of, world_index

pytestmark = pytest.mark.gpu
SAMPLING = golden_names("sampling")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _circuit(g, dev):
    from cirkit_b200 import B200Circuit

    cc = B200Circuit(g.plan)
    with torch.no_grad():
        for p, v in zip(cc.leaves, g.leaves(torch.float32)):
            p.copy_(v)
    return cc.to(dev)


@pytest.mark.parametrize("name", SAMPLING)
def test_kernel_matches_the_numpy_sampler(name, dev):
    from oracle.sampling import ancestral_sample

    g = Golden(name)
    cc = _circuit(g, dev)
    N, seed = 20000, 77
    x, mixes = cc.sample_query(N, seed=seed)
    assert x.shape == (N, g.plan.num_variables) and x.dtype == torch.int64 and x.is_cuda
    xo, mo = ancestral_sample(oracle_of(g, torch.float32), N, seed=seed)
    same = (x.cpu().numpy() == xo).all(axis=1)
    assert same.mean() >= 0.999, f"{(~same).sum()} of {N} samples differ from the numpy sampler"
    sums = [sid for sid, s in enumerate(g.plan.steps) if s.kind in ("sum", "cpt", "mixing", "tucker")]
    assert len(mixes) == len(sums)
    for sid, m in zip(sums, mixes):
        assert m.shape == (g.plan.steps[sid].num_folds, N)
        agree = (m.cpu().numpy() == mo[sid])[:, same]
        assert agree.all(), f"step {sid}: mixture draws differ on samples whose values agree"
    # drawing the batch in chunks continues the same stream
    xc, _ = cc.runtime.sample(N, list(cc.leaves), seed=seed, chunk=3000)
    assert torch.equal(xc, x)
    # and the default seed comes from torch's generator
    torch.manual_seed(3)
    a, _ = cc.sample_query(100)
    torch.manual_seed(3)
    b, _ = cc.sample_query(100)
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", SAMPLING)
def test_samples_follow_the_reference_distribution(name, dev):
    g = Golden(name)
    V = g.meta["num_categories"]
    probs = np.exp(g.y().numpy().reshape(-1))
    cc = _circuit(g, dev)
    N = 1_000_000
    x, _ = cc.runtime.sample(N, list(cc.leaves), seed=11, return_mixtures=False)
    assert int(x.min()) >= 0 and int(x.max()) < V
    counts = np.bincount(world_index(x.cpu().numpy(), V), minlength=len(probs))
    stat, bound = chi_square_ok(counts, probs)
    assert stat < bound, f"chi-square {stat:.1f} >= {bound:.1f}"
    # the reference's criterion (ratios within 3e-2), on the states whose frequency is resolved
    # to better than that at this sample size (p >= 0.02: one standard deviation is 0.7 %)
    big = probs >= 2e-2
    if big.any():
        np.testing.assert_allclose(counts[big] / N, probs[big], rtol=3e-2)


def test_gaussian_mixture_samples(dev):
    """1-D Gaussian mixture (fixture gmm1d_k8): the empirical CDF of 10^6 samples against the
    mixture CDF evaluated in float64 from the fixture's parameters (Kolmogorov distance)."""
    import math

    g = Golden("gmm1d_k8")
    cc = _circuit(g, dev)
    x, mixes = cc.sample_query(1_000_000, seed=5)
    assert x.shape == (1_000_000, 1) and x.dtype == torch.float32
    oc = oracle_of(g)
    steps = g.plan.steps
    with torch.no_grad():
        mean = oc.param(steps[0].params["mean"]).reshape(-1).double()
        std = oc.param(steps[0].params["stddev"]).reshape(-1).double()
        w = oc.param(steps[1].params["weight"]).reshape(-1).double()
    xs = torch.sort(x[:, 0].double().cpu()).values
    cdf = (w * 0.5 * (1 + torch.erf((xs[:, None] - mean) / (std * math.sqrt(2))))).sum(dim=1)
    emp = (torch.arange(len(xs), dtype=torch.float64) + 0.5) / len(xs)
    assert float((cdf - emp).abs().max()) < 2.5e-3  # KS 1e-9 quantile at N = 10^6 is 3.3e-3
    # the component frequencies are the mixture weights
    comp = torch.bincount(mixes[0][0].cpu().long(), minlength=len(w)).double() / len(xs)
    assert float((comp - w).abs().max()) < 3e-3


def test_error_behaviour(dev):
    from cirkit_b200 import B200Circuit, SamplingQuery

    g = Golden("smp_qg9_cpt_k4")
    cc = _circuit(g, dev)
    q = SamplingQuery(cc)
    with pytest.raises(ValueError, match="positive number"):
        q(0)
    # un-normalised sum weights: ValueError from CP-T layers (layers/optimized.py:182-188)
    bad = Golden("qt8_cp_k6_embedding")
    with pytest.raises(TypeError, match="not supported for layers of type embedding"):
        _circuit(bad, dev).sample_query(4)


def test_benchmark_circuit_samples(dev):
    """QuadTree 28x28, K = 64 (the north-star circuit; the reference would need an (F, K, N, D)
    tensor of 784 * 64 * N * 784 values): 4096 images, every pixel in range, per-pixel marginals
    of the first input fold against the circuit's own marginal query."""
    from cirkit_b200 import B200Circuit

    g = Golden("qt28_cp_k64")
    cc = B200Circuit(g.plan, seed=1234).to(dev)
    N = 4096
    x, mixes = cc.sample_query(N, seed=1)
    assert x.shape == (N, 784) and int(x.min()) >= 0 and int(x.max()) <= 255
    assert all((m >= 0).all() for m in mixes)  # a tree: every fold lies on every sample's path
    # marginal of pixel 0 from the circuit: integrate every other variable
    mask = torch.ones(1, 784, dtype=torch.bool, device=dev)
    mask[0, 0] = False
    xs = torch.zeros(256, 784, dtype=torch.int64, device=dev)
    xs[:, 0] = torch.arange(256, device=dev)
    with torch.no_grad():
        p0 = torch.exp(cc.integrate_query(xs, mask).reshape(-1).double()).cpu().numpy()
    assert abs(p0.sum() - 1.0) < 1e-4
    big, _ = cc.runtime.sample(200_000, list(cc.leaves), seed=2, return_mixtures=False, chunk=1 << 14)
    counts = np.bincount(big[:, 0].cpu().numpy(), minlength=256)
    stat, bound = chi_square_ok(counts, p0 / p0.sum())
    assert stat < bound, f"chi-square {stat:.1f} >= {bound:.1f}"
