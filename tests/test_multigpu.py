"""Real multi-GPU runs under the test runner: every world size the box offers (2 / 4 / 8) is
launched with torchrun and must print `ALL OK` for the whole matrix of
`tests/mp_nvlink_worker.py` — run options x optimizers x {eager, CUDA graph} x {P2P, NVLS} x
early push, sync and async, the in-engine NCCL arm, extended optimizers, variable row counts,
sharded checkpoints and the public collectives' dtype x size sweep against NCCL.

Model: Horovod runs its whole pytest suite under 2 real ranks
(`horovod/.buildkite/gen-pipeline.sh:99-100`, `horovod/test/test_tensorflow.py:70-948`).
Skipped on boxes with a single GPU (`multigpu` marker, tests/conftest.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:  # pragma: no cover
        return 0


def _torchrun(nproc, script, *args, timeout=900):
    port = 29600 + nproc
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), script] + list(args)
    env = dict(os.environ, PYTHONPATH=ROOT)
    # own session: on a timeout the WHOLE tree (torchrun + ranks, possibly with kernels
    # spinning on a dead peer) is killed, so nothing is left on the GPUs for the next test
    p = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        import signal
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        out += "\n[test_multigpu] TIMEOUT after %d s — process group killed\n" % timeout
    return subprocess.CompletedProcess(cmd, p.returncode, out, err)


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_correctness_matrix(nproc):
    if _ngpu() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    quick = ["--quick"] if nproc != 2 else []        # full matrix at 2 ranks, the quick one above
    r = _torchrun(nproc, os.path.join("tests", "mp_nvlink_worker.py"), *quick)
    tail = "\n".join((r.stdout + "\n" + r.stderr).splitlines()[-60:])
    assert r.returncode == 0 and "ALL OK" in r.stdout, tail
    assert "FAIL" not in r.stdout, tail


@pytest.mark.timeout(900)
@pytest.mark.parametrize("nproc", [2, 8])
def test_example_families(nproc):
    """NMT / skip-thoughts / CNN harness on N real GPUs (the reference ships them as
    multi-GPU examples, `parallax/parallax/examples/{nmt,skip_thoughts,tf_cnn_benchmarks}`)."""
    if _ngpu() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    r = _torchrun(nproc, os.path.join("tests", "mp_examples_worker.py"), timeout=800)
    tail = "\n".join((r.stdout + "\n" + r.stderr).splitlines()[-40:])
    assert r.returncode == 0 and "ALL OK" in r.stdout, tail


@pytest.mark.timeout(300)
def test_stalled_rank_is_reported():
    """One rank stops issuing steps: the watchdog names the missing rank and shuts the job
    down instead of hanging (`horovod/test/test_stall.py:13-26`)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, os.path.join("tests", "mp_stall_worker.py"), timeout=150)
    out = r.stdout + r.stderr
    assert "missing rank" in out.lower() or "stalled" in out.lower(), out[-3000:]
    assert r.returncode != 0
