"""Resource-file parsing and (de)serialisation.

Parity: reference `common/lib.py:101-191` (`parse_machine_info`,
`parse_resource_info`, `serialize_resource_info`,
`deserialize_resource_info`, `get_cluster_str_for_hosts`) and
`doc/quick_start.md:5-11` for the file format: one host per line,
``host[:gpu,gpu,...]``; the first host is the master/chief.

Differences by design: GPU discovery for a host that lists no GPUs is local
(`torch.cuda.device_count()` / ``/proc/driver/nvidia/gpus``) for localhost
and ssh for remote hosts; ports come from the local kernel
(bind to port 0) instead of ``python -m ephemeral_port_reserve`` over ssh.
"""
import os
import socket
import subprocess

LOCAL_NAMES = ("localhost", "127.0.0.1", "::1")


def is_local_host(hostname):
    if hostname in LOCAL_NAMES:
        return True
    try:
        return hostname in (socket.gethostname(), socket.getfqdn())
    except Exception:  # pragma: no cover
        return False


def routable_address(hostname):
    """An address of `hostname` that OTHER machines can connect to: the name as written
    in the resource file, unless it is a loopback alias of this machine — then this
    machine's own name / primary address."""
    if hostname not in LOCAL_NAMES:
        return hostname
    try:
        name = socket.getfqdn()
        addr = socket.gethostbyname(name)
        if not addr.startswith("127."):
            return addr
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        try:
            s.connect(("10.255.255.255", 1))        # no packet is sent
            return s.getsockname()[0]
        finally:
            s.close()
    except Exception:  # pragma: no cover
        return socket.gethostname()


def all_local(resource_info):
    return all(is_local_host(w["hostname"]) for w in resource_info["worker"])


def _get_available_gpus(hostname):
    """GPU ordinals on `hostname` (reference `common/lib.py:101-103`)."""
    if is_local_host(hostname):
        path = "/proc/driver/nvidia/gpus"
        if os.path.isdir(path):
            return list(range(len(os.listdir(path))))
        try:
            import torch
            return list(range(torch.cuda.device_count()))
        except Exception:  # pragma: no cover
            return []
    out = subprocess.check_output(
        "ssh %s ls /proc/driver/nvidia/gpus" % hostname, shell=True).decode()
    return list(range(len(out.strip().split("\n"))))


def get_empty_port(num_ports=1):
    """Reserve `num_ports` free TCP ports on this host."""
    socks, ports = [], []
    for _ in range(num_ports):
        s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(("127.0.0.1", 0))
        socks.append(s)
        ports.append(s.getsockname()[1])
    for s in socks:
        s.close()
    return ports


def parse_machine_info(machine_str):
    machine_str = machine_str.strip()
    if not machine_str or machine_str.startswith("#"):
        return []
    hostname_gpus = machine_str.split(":")
    hostname = hostname_gpus[0].strip()
    if len(hostname_gpus) == 1 or hostname_gpus[1].strip() == "":
        gpus = _get_available_gpus(hostname)
    elif len(hostname_gpus) == 2:
        gpus = [int(g) for g in hostname_gpus[1].split(",") if g.strip() != ""]
    else:
        raise ValueError("bad resource line %r (expected host[:gpu,gpu,...])"
                         % machine_str)
    return [(hostname, gpus)]


def parse_resource_info(path_or_text, run_option="HYBRID"):
    """Returns ``{'master': [...], 'ps': [...], 'worker': [...]}`` exactly like
    the reference (`common/lib.py:136-150`): master = first host, one PS entry
    per host, HYBRID workers reserve one port per GPU.

    `path_or_text` may be a path, or the text of a resource file (handy for
    tests and single-box runs: ``"localhost:0,1,2,3"``).
    """
    if os.path.exists(path_or_text):
        with open(path_or_text) as f:
            lines = f.readlines()
    else:
        lines = path_or_text.splitlines()
    machines = []
    for line in lines:
        machines.extend(parse_machine_info(line))
    if not machines:
        raise ValueError("resource info names no machines")
    master_host = machines[0][0]
    master = [{"hostname": master_host, "port": get_empty_port(1), "gpus": []}]
    ps = [{"hostname": h, "port": get_empty_port(1), "gpus": []}
          for h, _ in machines]
    worker = [{"hostname": h,
               "port": get_empty_port(1 if run_option != "HYBRID" or len(g) == 0
                                      else len(g)),
               "gpus": list(g)} for h, g in machines]
    return {"master": master, "ps": ps, "worker": worker}


def serialize_resource_info(resource_info):
    """``type_host:ports:gpus+…^…`` (reference `common/lib.py:153-158`)."""
    def ser_machine(m):
        return "%s:%s:%s" % (m["hostname"],
                             ",".join(str(p) for p in m["port"]),
                             ",".join(str(g) for g in m["gpus"]))
    return "^".join("%s_%s" % (t, "+".join(ser_machine(m) for m in ms))
                    for t, ms in resource_info.items())


def deserialize_resource_info(serialized):
    def de_list(s):
        return [int(x) for x in s.split(",")] if s else []

    def de_machine(m):
        hostname, ports, gpus = m.split(":")
        return {"hostname": hostname, "port": de_list(ports),
                "gpus": de_list(gpus)}
    info = {}
    for tm in serialized.strip().split("^"):
        t, machines = tm.split("_", 1)
        info[t] = [de_machine(m) for m in machines.split("+")]
    return info


def get_cluster_str_for_hosts(hosts, with_slots):
    if with_slots:
        return ",".join("%s:%d" % (h["hostname"], len(h["gpus"])) for h in hosts)
    out = []
    for h in hosts:
        for p in h["port"]:
            out.append("%s:%d" % (h["hostname"], p))
    return ",".join(out)


def worker_layout(resource_info):
    """Flat list of ``(hostname, machine_id, local_rank, gpu_ordinal)`` — one
    entry per worker process (one per GPU; hosts without GPUs contribute one
    CPU worker).  worker_id = Σ gpus(previous machines) + local_rank
    (reference `hybrid/runner.py:197-200`)."""
    layout = []
    for mid, w in enumerate(resource_info["worker"]):
        gpus = w["gpus"] if w["gpus"] else [None]
        for lr, g in enumerate(gpus):
            layout.append((w["hostname"], mid, lr, g))
    return layout


def num_machines(resource_info):
    return len(resource_info["worker"])
