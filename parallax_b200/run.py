"""`python -m parallax_b200.run -np 8 [-H host:slots,...] script.py args…`

Launch a script on N ranks with the rendezvous environment the collectives API
(`parallax_b200.collectives`) expects — the analogue of `horovodrun`
(`horovod/run/run.py`, `horovod/bin/horovodrun`): local ranks are plain
subprocesses, ranks on other hosts are started over ssh; stdout/stderr of every
rank is prefixed with its rank; the job fails as soon as one rank fails.
"""
import argparse
import os
import subprocess
import sys
import threading

from .launcher import kill_all, remote_exec
from .resource import get_empty_port, is_local_host


def parse_hosts(spec, np_):
    if not spec:
        return [("localhost", np_)]
    out = []
    for item in spec.split(","):
        h, _, s = item.partition(":")
        out.append((h, int(s) if s else 1))
    return out


def _pump(stream, prefix):
    for line in iter(stream.readline, b""):
        sys.stdout.write("[%s] %s" % (prefix, line.decode(errors="replace")))
        sys.stdout.flush()


def main(argv=None):
    ap = argparse.ArgumentParser(prog="parallax_b200.run")
    ap.add_argument("-v", "--version", action="store_true", help="print the version and exit")
    ap.add_argument("-np", "--num-proc", type=int, default=None)
    ap.add_argument("-H", "--hosts", "--host", default=None, help="host:slots[,host:slots...]")
    ap.add_argument("-p", "--ssh-port", type=int, default=22, help="ssh port of remote hosts")
    ap.add_argument("--master-port", type=int, default=None)
    ap.add_argument("--start-timeout", type=int, default=None,
                    help="seconds the ranks wait for each other at start-up (default 600)")
    ap.add_argument("--disable-cache", action="store_true",
                    help="accepted for horovodrun compatibility (there are no cached host checks)")
    ap.add_argument("--verbose", action="store_true", help="log the launch commands")
    ap.add_argument("command", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    if a.version:
        from . import __version__
        print(__version__)
        return 0
    if a.num_proc is None:
        ap.error("-np is required")
    if not a.command:
        ap.error("no command given")
    hosts = parse_hosts(a.hosts, a.num_proc)
    if sum(s for _, s in hosts) < a.num_proc:
        ap.error("not enough slots for -np %d" % a.num_proc)
    # loopback only when every rank runs here; ranks on other hosts need a reachable address
    from .resource import routable_address
    everything_local = all(is_local_host(h) for h, _ in hosts)
    master = "127.0.0.1" if everything_local else routable_address(hosts[0][0])
    port = a.master_port or get_empty_port(1)[0]
    cmd = a.command if a.command[0] != "--" else a.command[1:]
    if cmd[0].endswith(".py"):
        cmd = [sys.executable] + cmd
    procs, threads, rank = [], [], 0
    for host, slots in hosts:
        for local in range(slots):
            if rank >= a.num_proc:
                break
            env = {"RANK": rank, "WORLD_SIZE": a.num_proc, "LOCAL_RANK": local,
                   "LOCAL_WORLD_SIZE": slots, "MASTER_ADDR": master, "MASTER_PORT": port}
            if a.start_timeout:
                env["PARALLAX_START_TIMEOUT"] = a.start_timeout
            if a.verbose:
                sys.stderr.write("[run] rank %d on %s: %s\n" % (rank, host, " ".join(cmd)))
            if is_local_host(host):
                penv = dict(os.environ)
                penv.update({k: str(v) for k, v in env.items()})
                p = subprocess.Popen(cmd, env=penv, stdout=subprocess.PIPE,
                                     stderr=subprocess.STDOUT, preexec_fn=os.setsid)
            else:
                p = remote_exec("cd %s; %s" % (os.getcwd(), " ".join(cmd)), host,
                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env,
                                port=a.ssh_port)
            t = threading.Thread(target=_pump, args=(p.stdout, str(rank)), daemon=True)
            t.start()
            procs.append(p)
            threads.append(t)
            rank += 1
    rc = 0
    try:
        import time
        while any(p.poll() is None for p in procs):
            for p in procs:
                if p.poll() not in (None, 0):
                    rc = p.returncode
                    kill_all(procs)
                    break
            time.sleep(0.1)
        rc = rc or max((p.returncode or 0) for p in procs)
    except KeyboardInterrupt:
        kill_all(procs)
        rc = 130
    for t in threads:
        t.join(timeout=2)
    return rc


if __name__ == "__main__":
    sys.exit(main())
