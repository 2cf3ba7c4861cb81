"""Parameter-server launcher — kept for command-line parity with the reference's
`parallax/core/python/tools/launch_ps.py:22-53`, which starts a
`tf.train.Server(job_name='ps')` per host and joins it.

On an NVSwitch box there are no separate server processes: the "server" of a
variable is the GPU that owns it (dense: the rank owning the bucket slice;
sparse: the rank owning the partition), and it is served by the worker process
of that GPU.  This entry point therefore only validates its arguments, prints
the owner map that the engine would use, and exits 0 so that scripts written
for the reference keep working.
"""
import argparse
import sys

from ..analyzer import greedy_load_balance


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--ps_hosts", default="")
    ap.add_argument("--worker_hosts", default="")
    ap.add_argument("--job_name", default="ps")
    ap.add_argument("--task_index", type=int, default=0)
    ap.add_argument("--protocol", default="grpc")
    a = ap.parse_args(argv)
    workers = [h for h in a.worker_hosts.split(",") if h]
    print("parallax_b200: no stand-alone parameter servers; %d worker GPU(s) own the "
          "variables (protocol=%s is served over NVLink peer memory)."
          % (max(len(workers), 1), a.protocol))
    print("example byte-greedy owner map for sizes [8,4,4,2,1] over %d owners: %s"
          % (max(len(workers), 1),
             greedy_load_balance([8, 4, 4, 2, 1], max(len(workers), 1))))
    return 0


if __name__ == "__main__":
    sys.exit(main())
