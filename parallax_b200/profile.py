"""Step profiling driven by `ProfileConfig`.

Parity: reference `common/session_context.py:74-88,149-167` (on a profile
step force FULL_TRACE and dump RunMetadata to
``<profile_dir>/<host>/worker:<id>/run_meta/run_meta_<global_step>``),
`common/lib.py:333-358` (directory layout, ``task_info`` file) and
`mpi/runner.py:169-172` (only one worker per machine may attach the CUPTI
tracer → `profile_worker`).

Here a profile step runs under `torch.profiler` (CUPTI kernel records on
CUDA, CPU ops otherwise) and writes a Chrome trace
``run_meta_<global_step>.json`` plus a ``.txt`` kernel table in the same
directory layout.
"""
import os
import socket

from .log import parallax_log


def create_profile_directory(profile_dir, hostname, worker_id):
    d = os.path.join(profile_dir, hostname, "worker:%d" % worker_id, "run_meta")
    os.makedirs(d, exist_ok=True)
    return d


def append_task_info(profile_dir, hostname, tasks):
    d = os.path.join(profile_dir, hostname)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "task_info"), "a") as f:
        for t in tasks:
            f.write(t + "\n")


class StepProfiler(object):
    def __init__(self, profile_config, worker_id, local_rank=0):
        pc = profile_config
        self.dir = pc.profile_dir
        self.steps = set(pc.profile_steps) if pc.profile_steps else None
        self.range = tuple(pc.profile_range) if pc.profile_range else None
        target = pc.profile_worker
        # default: the first worker of each machine (one CUPTI client/machine)
        self.enabled = bool(self.dir) and (
            (target is None and local_rank == 0) or target == worker_id)
        self.worker_id = worker_id
        self.host = socket.gethostname()
        self._prof = None
        if self.enabled:
            self.out = create_profile_directory(self.dir, self.host, worker_id)
            append_task_info(self.dir, self.host, ["worker:%d" % worker_id])

    def is_profile_step(self, step):
        if not self.enabled:
            return False
        if self.steps is not None:
            return step in self.steps
        if self.range is not None:
            return self.range[0] <= step < self.range[1]
        return False

    def start(self):
        import torch
        from torch.profiler import profile, ProfilerActivity
        acts = [ProfilerActivity.CPU]
        if torch.cuda.is_available():
            acts.append(ProfilerActivity.CUDA)
        self._prof = profile(activities=acts, record_shapes=False)
        self._prof.__enter__()

    def stop(self, step):
        prof, self._prof = self._prof, None
        prof.__exit__(None, None, None)
        base = os.path.join(self.out, "run_meta_%d" % step)
        try:
            prof.export_chrome_trace(base + ".json")
            with open(base + ".txt", "w") as f:
                f.write(prof.key_averages().table(row_limit=50))
        except Exception as e:  # pragma: no cover
            parallax_log.warning("profile dump failed: %s", e)
        return base + ".json"
