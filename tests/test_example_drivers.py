"""The example drivers as a user runs them: the script is the launcher AND the
worker; two CPU workers (host fabric) are spawned from a two-line resource file."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp_path, script, args, timeout=420):
    res = tmp_path / "resource_info"
    res.write_text("localhost\nlocalhost\n")
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("PARALLAX_") and k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.update(PARALLAX_FABRIC="host", CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    cmd = [sys.executable, os.path.join(ROOT, "examples", script), "--resource_info_file", str(res),
           "--redirect_path", str(tmp_path / "logs")] + args
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout,
                       cwd=str(tmp_path))
    logs = {}
    for i in (0, 1):
        p = tmp_path / "logs" / ("log_worker%d_stderr" % i)
        logs[i] = p.read_text() if p.exists() else ""
    assert r.returncode == 0, (r.stderr[-1500:], logs[0][-1500:], logs[1][-1500:])
    return logs


def test_nmt_driver_two_workers(tmp_path):
    logs = _run(tmp_path, "nmt/nmt_distributed_driver.py", [
        "--synthetic", "--out_dir", str(tmp_path / "nmt"), "--num_train_steps", "24",
        "--steps_per_eval", "12", "--save_ckpt_steps", "12", "--hparams",
        "num_units=16,num_layers=2,batch_size=16,steps_per_stats=6,infer_batch_size=100,"
        "num_embeddings_partitions=2,beam_width=2"])
    assert "step 24" in logs[0] and "final: dev/test ppl" in logs[0]
    assert "final: dev/test ppl" not in logs[1]             # only worker 0 reports
    assert (tmp_path / "nmt" / "model.ckpt-24.pt").exists()
    assert (tmp_path / "nmt" / "output_dev").read_text().count("\n") == 200


def test_skip_thoughts_driver_two_workers(tmp_path):
    logs = _run(tmp_path, "skip_thoughts/skip_distributed_driver.py", [
        "--synthetic", "--train_dir", str(tmp_path / "st"), "--vocab_size", "100",
        "--word_embedding_dim", "8", "--encoder_dim", "16", "--batch_size", "16",
        "--num_embedding_partitions", "2", "--max_steps", "20", "--log_frequency", "10",
        "--learning_rate", "0.01"])
    assert "global step 20" in logs[0] and "global step" not in logs[1]


@pytest.mark.parametrize("extra", [["--model", "lenet", "--optimizer", "momentum"],
                                   ["--model", "trivial", "--forward_only"]])
def test_cnn_driver_two_workers(tmp_path, extra):
    logs = _run(tmp_path, "cnn_benchmarks/CNNBenchmark_distributed_driver.py", extra + [
        "--batch_size", "4", "--num_batches", "6", "--num_warmup_batches", "2",
        "--display_every", "3"])
    assert "total images/sec" in logs[0] and "Batch size:  8 global / 4 per device" in logs[0]
    assert "total images/sec" not in logs[1]


def test_lm1b_and_simple_drivers_two_workers(tmp_path):
    logs = _run(tmp_path, "lm1b/lm1b_distributed_driver.py", [
        "--use_synthetic", "--tiny", "--max_steps", "11", "--log_frequency", "5", "--hpconfig",
        "batch_size=8,num_steps=4,keep_prob=1.0", "--logdir", str(tmp_path / "lm1b")])
    assert "Iteration 11" in logs[0] and (tmp_path / "lm1b" / "analysis_worker_1.json").exists()
    logs = _run(tmp_path, "simple/simple_driver.py", [])
    assert "learned" in logs[0] or "step" in logs[0]


def test_horovod_style_examples_under_run_cli(tmp_path):
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("PARALLAX_") and k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.update(CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    base = [sys.executable, "-m", "parallax_b200.run", "-np", "2"]
    r = subprocess.run(base + [os.path.join(ROOT, "examples/horovod/pytorch_mnist.py"),
                               "--epochs", "4", "--no-cuda", "--lr", "0.05", "--num-synthetic",
                               "2048"], env=env, cwd=ROOT, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "finished gradual learning rate warmup to 0.1" in r.stdout
    last = [l for l in r.stdout.splitlines() if "Epoch 4:" in l][-1]
    assert float(last.split("accuracy")[1].split()[0]) > 0.5
    r = subprocess.run(base + [os.path.join(ROOT, "examples/horovod/pytorch_synthetic_benchmark.py"),
                               "--model", "lenet", "--batch-size", "4", "--num-iters", "2",
                               "--num-batches-per-iter", "2", "--num-warmup-batches", "1",
                               "--fp16-allreduce", "--no-cuda"], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "Total img/sec on 2 worker(s)" in r.stdout


@pytest.mark.parametrize("extra", [[], ["--sparse-as-dense"]])
def test_horovod_style_word2vec_sparse_gradients(tmp_path, extra):
    """Row-sparse embedding gradients through `DistributedOptimizer`: all-gather of (indices,
    values) — Horovod's IndexedSlices path — or a dense all-reduce; replicas stay identical."""
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("PARALLAX_") and k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.update(CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "parallax_b200.run", "-np", "2",
                        os.path.join(ROOT, "examples/horovod/pytorch_word2vec.py"),
                        "--steps", "120", "--no-cuda"] + extra, env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    last = [l for l in r.stdout.splitlines() if "neighbour score" in l][-1]
    first_loss, last_loss = (float(x) for x in
                             last.split("loss ")[1].split(";")[0].split(" -> "))
    assert last_loss < 0.3 * first_loss
    nb, rnd = float(last.split("neighbour score ")[1].split()[0]), \
        float(last.split("vs random ")[1].split(";")[0])
    assert nb > rnd + 1.0
    assert float(last.split("replica spread ")[1].split(";")[0]) < 1e-6
    assert ("dense all-reduce" if extra else "all-gather of (indices, values)") in last


def test_lm1b_train_checkpoint_then_eval_with_and_without_ema(tmp_path):
    """`lm1b_distributed_driver.py` writes checkpoints, `lm1b_eval.py` restores the latest one and
    reports the perplexity of the full softmax — with the raw weights and with the EMA shadows of
    the LSTM variables (`examples/lm1b/lm1b_eval.py:96-104` of the reference)."""
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("PARALLAX_") and k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.update(PARALLAX_FABRIC="host", CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    ck = str(tmp_path / "ck")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples/lm1b/lm1b_distributed_driver.py"),
                        "--use_synthetic", "--tiny", "--max_steps", "12", "--ckpt_dir", ck,
                        "--save_ckpt_steps", "6", "--logdir", str(tmp_path / "log")],
                       env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stderr[-1500:]
    assert os.path.exists(os.path.join(ck, "model.ckpt-12.pt"))
    ppl = {}
    for extra in ([], ["--use_ema"]):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "examples/lm1b/lm1b_eval.py"),
                            "--use_synthetic", "--tiny", "--ckpt_dir", ck, "--max_batches", "3"]
                           + extra, env=env, cwd=str(tmp_path), capture_output=True, text=True,
                           timeout=400)
        assert r.returncode == 0, r.stderr[-1500:]
        assert "global_step 12" in r.stderr + r.stdout
        ppl[bool(extra)] = float(r.stdout.strip().splitlines()[-1].split()[1])
    assert all(1000.0 < v < 1e6 for v in ppl.values())       # ~ vocabulary size on random ids
    assert ppl[True] != ppl[False]                            # the EMA shadows were used
