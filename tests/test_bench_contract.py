"""bench.py contract (no GPU): the reference arm prints ONE JSON line with the keys the driver
reads, for every workload, and the workload table is consistent with the fixtures."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload", ["qt28_cp_k64", "qt28_cp_k32"])
def test_reference_arm_line(workload):
    out = subprocess.run(
        [sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", workload,
         "--steps", "1", "--warmup", "0", "--cpu-batch", "8"],
        capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["scaling"] == "weak"
    # the unmodified reference from baseline/_ref when that install is present, else the oracle port
    have_ref = os.path.isdir(os.path.join(REPO, "baseline", "_ref", "cirkit"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    if workload == "qt28_cp_k64":
        base = json.load(open(os.path.join(REPO, "BASELINE.json")))
        assert d["metric"].split(" @")[0].replace("x", "×") in base["metric"].replace("x", "×")


def test_workloads_resolve():
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import bench

    sizes = {"qt28_cp_k64": (150401, 19259456), "qt28_cp_k32": (75201, 8026144),
             "qt28_tucker_k64": (100225, 217845760), "pd32_cp_k128": (2864783, 299603982),
             "rbt64_sos_k64": (8065, 1302592)}  # c(x) of the squared circuit (built by the reference's front-end)
    for name in bench.WORKLOADS:
        if name == "rbt64_sos_k64" and not os.path.isdir(os.path.join(REPO, "baseline", "_ref", "cirkit")):
            continue
        plan = bench.load_plan(name).plan
        assert (plan.activation_units(), plan.parameter_elements()) == sizes[name], name
        assert bench.metric_name(name).startswith("samples/sec (fwd+bwd log-lik)")


def test_reference_arm_under_torchrun_prints_once():
    """The driver launches the reference arm like the CUDA arm (torchrun, N ranks): rank 0 alone
    runs and prints the line, the other ranks exit 0 without work."""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(REPO, "bench.py"),
         "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-batch", "8"],
        capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
