"""SamplingQuery checkers on the CPU (oracle/sampling.py) against the reference's exact joint
distributions (tests/golden/smp_*.npz: `y` = the reference's log-probability of EVERY joint state,
`ref_counts` = world counts of the reference's own SamplingQuery).  The top-down sampler restated
here in numpy is the algorithm of csrc/sampling_kernels.cu; tests/test_gpu_sampling.py compares the
kernel with it sample by sample."""
import numpy as np
import pytest
import torch

from helpers import Golden, golden_names
from oracle import OracleCircuit
from oracle.sampling import ancestral_sample, philox4x32_10, reference_sample

SAMPLING = golden_names("sampling")


def world_index(x: np.ndarray, V: int) -> np.ndarray:
    D = x.shape[1]
    return (x.astype(np.int64) * np.array([V ** (D - 1 - i) for i in range(D)], dtype=np.int64)).sum(axis=1)


def chi_square_ok(counts: np.ndarray, probs: np.ndarray, min_expected: float = 10.0) -> tuple[float, float]:
    """Pearson statistic of `counts` against `probs` and the acceptance bound dof + 6 sqrt(2 dof)
    (six standard deviations of the chi-square law).  States are sorted by probability and
    merged into cells of expectation >= min_expected, so the test keeps its power when most of
    the joint states are rarer than 1 / N."""
    n = counts.sum()
    order = np.argsort(-probs, kind="stable")
    exp, cnt = n * probs[order], counts[order]
    cells_e, cells_c, e_acc, c_acc = [], [], 0.0, 0
    for e, c in zip(exp, cnt):
        e_acc += e
        c_acc += c
        if e_acc >= min_expected:
            cells_e.append(e_acc)
            cells_c.append(c_acc)
            e_acc, c_acc = 0.0, 0
    if e_acc > 0 and cells_e:
        cells_e[-1] += e_acc
        cells_c[-1] += c_acc
    e, c = np.array(cells_e), np.array(cells_c)
    assert len(e) >= 2, "not enough samples for a chi-square test"
    stat = float(((c - e) ** 2 / e).sum())
    dof = len(e) - 1
    return stat, dof + 6.0 * np.sqrt(2.0 * dof)


def oracle_of(g: Golden, dtype=torch.float64) -> OracleCircuit:
    oc = OracleCircuit(g.plan, dtype=dtype)
    with torch.no_grad():
        for p, v in zip(oc.leaves, g.leaves(dtype)):
            p.copy_(v)
    return oc


def test_philox_known_answers():
    """Random123's published known-answer vectors for philox4x32-10."""
    z = philox4x32_10(np.zeros((1, 4), dtype=np.uint32), 0, 0)[0]
    assert [hex(int(v)) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = philox4x32_10(np.full((1, 4), 0xFFFFFFFF, dtype=np.uint32), 0xFFFFFFFF, 0xFFFFFFFF)[0]
    assert [hex(int(v)) for v in f] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    p = philox4x32_10(np.array([[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], dtype=np.uint32),
                      0xA4093822, 0x299F31D0)[0]
    assert [hex(int(v)) for v in p] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


@pytest.mark.parametrize("name", SAMPLING)
def test_fixture_is_a_distribution(name):
    g = Golden(name)
    V, D = g.meta["num_categories"], g.plan.num_variables
    assert g.x().shape == (V ** D, D)
    assert abs(float(np.exp(g.y().numpy()).sum()) - 1.0) < 1e-9
    if "ref_counts" in g.meta:  # the reference's own sampler follows its own distribution
        stat, bound = chi_square_ok(np.array(g.meta["ref_counts"]), np.exp(g.y().numpy().reshape(-1)))
        assert stat < bound


@pytest.mark.parametrize("name", SAMPLING)
def test_top_down_sampler_follows_the_reference_distribution(name):
    g = Golden(name)
    V = g.meta["num_categories"]
    probs = np.exp(g.y().numpy().reshape(-1))
    N = 60000
    x, mixes = ancestral_sample(oracle_of(g, torch.float32), N, seed=2024)
    assert x.shape == (N, g.plan.num_variables) and x.min() >= 0 and x.max() < V
    counts = np.bincount(world_index(x, V), minlength=len(probs))
    stat, bound = chi_square_ok(counts, probs)
    assert stat < bound, f"chi-square {stat:.1f} >= {bound:.1f}"
    # every sum-type layer reports its draws; a sample visits the root row
    sums = [sid for sid, s in enumerate(g.plan.steps) if s.kind in ("sum", "cpt", "mixing", "tucker")]
    assert sorted(mixes) == sums
    root = int(g.plan.out_step[0])
    assert (mixes[root][int(g.plan.out_fold[0])] >= 0).all()
    # chunked generation continues the same stream
    xa, _ = ancestral_sample(oracle_of(g, torch.float32), 100, seed=2024)
    xb, _ = ancestral_sample(oracle_of(g, torch.float32), 60, seed=2024, sample_base=40)
    assert np.array_equal(xa, x[:100]) and np.array_equal(xb, x[40:100])


@pytest.mark.parametrize("name", [n for n in SAMPLING if "tucker" not in n])
def test_bottom_up_restatement_follows_the_distribution(name):
    """`reference_sample` (pinned bit for bit to the reference in test_oracle_vs_reference.py)
    on the fixtures: the two samplers agree in distribution."""
    g = Golden(name)
    V = g.meta["num_categories"]
    probs = np.exp(g.y().numpy().reshape(-1))
    torch.manual_seed(5)
    smp, mix = reference_sample(oracle_of(g), 4000)
    counts = np.bincount(world_index(smp.numpy(), V), minlength=len(probs))
    stat, bound = chi_square_ok(counts, probs)
    assert stat < bound


def test_unsupported_layers_raise_like_the_reference():
    g = Golden("qt8_cp_k6_embedding")
    with pytest.raises(TypeError):
        ancestral_sample(oracle_of(g, torch.float32), 4, seed=1)
    with pytest.raises(TypeError):
        reference_sample(oracle_of(g), 4)
