"""Master/launcher path: process spawning over a resource file, env protocol,
redirect logs, analysis export, data sharding, and the partition search loop
(reference `common/runner.py:62-137`, `common/partitions.py:53-170`)."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "launch_script.py")


def _run(tmp_path, run_option, steps, search, timeout=300):
    res = tmp_path / "resource_info"
    res.write_text("localhost\nlocalhost\n")          # two CPU workers on this host
    env = dict(os.environ)
    for k in list(env):
        if k.startswith("PARALLAX_") or k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k)
    env["PARALLAX_FABRIC"] = "host"
    env["CUDA_VISIBLE_DEVICES"] = ""
    env["OMP_NUM_THREADS"] = "1"
    r = subprocess.run([sys.executable, SCRIPT, str(tmp_path), str(res), run_option,
                        str(steps), "1" if search else "0"], env=env, capture_output=True,
                       text=True, timeout=timeout)
    return r


def test_master_spawns_workers_and_exits(tmp_path):
    r = _run(tmp_path, "HYBRID", 5, False)
    assert r.returncode == 0, r.stderr[-2000:]
    outs = sorted(glob.glob(str(tmp_path / "worker_*.json")))
    assert len(outs) == 2
    infos = [json.load(open(o)) for o in outs]
    assert sorted(i["worker_id"] for i in infos) == [0, 1]
    for i in infos:
        assert i["num_workers"] == 2 and i["global_step"] == 5
        assert i["env_role"] == "PARALLAX_RUN_HYBRID"
        # shard(ds): element k kept iff k % 2 == worker_id
        assert i["first"] == [i["worker_id"], i["worker_id"] + 2, i["worker_id"] + 4]
    # redirect files + per-worker analysis dump
    assert os.path.exists(tmp_path / "logs" / "log_worker0_stderr")
    assert os.path.exists(tmp_path / "graph" / "analysis_worker_1.json")
    rep = json.load(open(tmp_path / "graph" / "analysis_worker_0.json"))
    assert rep["num_sparse"] == 1 and rep["tables"]["emb.weight"]["P"] == 2


def test_partition_search_relaunches_and_converges(tmp_path):
    """The launcher re-runs the job with P = 2, 4, ... measuring steps 50-100,
    stops when a candidate is slower / out of range, then runs the final job
    with the fitted optimum."""
    r = _run(tmp_path, "HYBRID", 110, True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    outs = glob.glob(str(tmp_path / "worker_*.json"))
    # the final (non-search) run completed on both workers with one agreed P
    final = [json.load(open(o)) for o in outs]
    ps = {}
    for i in final:
        ps.setdefault(i["partitions"], []).append(i["worker_id"])
    assert any(sorted(v) == [0, 1] for v in ps.values())
    assert "partition search: trying P=2" in r.stderr
    assert "optimal partitions" in r.stderr
    for i in final:
        assert i["tableP"] == i["partitions"]


def test_run_cli_like_horovodrun(tmp_path):
    """`python -m parallax_b200.run -np 2 script.py`: rank-prefixed output, rendezvous env,
    failure of one rank fails the job (horovodrun, `horovod/run/run.py`)."""
    script = tmp_path / "job.py"
    script.write_text(
        "import os, sys\n"
        "sys.path.insert(0, %r)\n"
        "import torch\n"
        "from parallax_b200 import collectives as hvd\n"
        "hvd.init()\n"
        "s = hvd.allreduce(torch.tensor([float(hvd.rank() + 1)]), average=False)\n"
        "r = hvd.rank()\n"
        "print('rank', r, 'sum', s.item(), os.environ.get('PARALLAX_START_TIMEOUT'))\n"
        "hvd.shutdown()\n"
        "sys.exit(3 if '--fail' in sys.argv and r == 1 else 0)\n" % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    base = [sys.executable, "-m", "parallax_b200.run"]
    r = subprocess.run(base + ["-np", "2", "--start-timeout", "90", "--verbose", str(script)],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "[0] rank 0 sum 3.0 90" in r.stdout and "[1] rank 1 sum 3.0 90" in r.stdout
    assert "[run] rank 1 on localhost" in r.stderr
    r = subprocess.run(base + ["-np", "2", str(script), "--fail"], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 3
    r = subprocess.run(base + ["--version"], env=env, cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "0.1.0"
    r = subprocess.run(base + ["-np", "3", "-H", "localhost:2", str(script)], env=env, cwd=ROOT,
                       capture_output=True, text=True)
    assert r.returncode != 0 and "not enough slots" in r.stderr


def test_multi_host_launch_uses_routable_master_and_keeps_secrets_off_argv(monkeypatch):
    """A worker started over ssh on another host must not be told MASTER_ADDR=127.0.0.1,
    and the search authkey must not appear on any command line."""
    import parallax_b200 as parallax
    from parallax_b200 import consts, launcher
    from parallax_b200.resource import parse_resource_info
    ri = parse_resource_info("localhost:0\nfarawayhost:0,1\n")
    seen = {"local": [], "remote": []}

    class _P(object):
        pid = 1

        def poll(self):
            return 0

    def fake_popen(cmd, env=None, **kw):
        seen["local"].append((cmd, env))
        return _P()

    def fake_remote(script, host, stdout=None, stderr=None, env=None, python_venv=None,
                    port=22, secret_names=()):
        seen["remote"].append((launcher.remote_command(script, host, env, python_venv, port,
                                                       [k for k in secret_names if k in env]),
                               env))
        return _P()
    monkeypatch.setattr(launcher.subprocess, "Popen", fake_popen)
    monkeypatch.setattr(launcher, "remote_exec", fake_remote)
    launcher.launch_workers("HYBRID", ri, parallax.Config(),
                            extra_env={consts.PARALLAX_SEARCH_AUTHKEY: "s3cr3t",
                                       consts.PARALLAX_SEARCH_ADDR: "x:1"})
    assert len(seen["local"]) == 1 and len(seen["remote"]) == 2
    addrs = {e["MASTER_ADDR"] for _, e in seen["local"]} | \
        {e["MASTER_ADDR"] for _, e in seen["remote"]}
    assert len(addrs) == 1 and not addrs.pop().startswith("127.")
    for cmd, env in seen["remote"]:
        assert "s3cr3t" not in " ".join(cmd) and "read -r PARALLAX_SEARCH_AUTHKEY" in cmd[-1]
    # a purely local job keeps loopback
    seen["local"].clear()
    launcher.launch_workers("HYBRID", parse_resource_info("localhost:0,1\n"), parallax.Config())
    assert {e["MASTER_ADDR"] for _, e in seen["local"]} == {"127.0.0.1"}
