"""Configuration objects — same names, fields and defaults as the reference.

Parity: reference `parallax/parallax/core/python/common/config.py:21-179`
(PSConfig, MPIConfig, CommunicationConfig, CheckPointConfig, ProfileConfig,
ParallaxConfig) and `doc/parallax_api.md:40-100`.

What each knob means on an NVSwitch box (one process per GPU, no CPU
parameter servers):

* ``run_option`` — ``MPI``/``AR``: dense grads all-reduced, sparse grads
  all-gathered and applied to a full local replica of the table;
  ``PS``: every variable has an owner GPU (greedy byte balancing), dense =
  reduce-to-owner → owner applies → mirror refresh, sparse = push to owner;
  ``HYBRID`` (default): dense like AR, sparse like PS.
* ``PSConfig.protocol`` — the reference picks the PS transport
  (grpc / grpc+verbs / grpc+gdr / grpc+mpi).  Here the transport is NVLink
  peer memory; accepted values additionally include ``"nvlink"`` (default
  behaviour for every legacy value) and ``"nccl"`` — the in-engine library arm: same
  engine, buckets and CUDA graph, but every cross-GPU byte goes through NCCL
  (dense: ncclAllReduce + local fused optimizer; sparse: all-gather of ids and rows +
  owner kernel; lookups: all-gather(ids) + local gather + reduce-scatter).  `bench.py`
  reports it as ``same_engine_nccl``.
* ``PSConfig.replicate_variables`` — True: owners *push* updated dense
  values into every GPU's mirror right after the update; False: workers
  *pull* owner values at the start of the next step.
* ``PSConfig.local_aggregation`` — dedup/sum duplicate indices on the sender
  before shipping (shared-memory hash tables inside the push kernel); False ships
  every (index,row) pair and the owner merges them.
* ``boundary_among_servers`` — the reference assigns each post-aggregation op to the
  server that owns the variable it feeds (`graph_transform_lib.py:174-327`).  Here the
  owner of a slice / partition always runs them; the flag selects how partitions are
  placed on owners: True = byte-greedy over all sparse variables
  (`ps/between_graph_parallel.py:49-70`), False = round-robin.
* ``boundary_between_workers_and_servers`` — the reference moves size-reducing ops to
  the producer and size-increasing casts to the consumer of every worker↔server edge
  (`graph_transform_lib.py:1315-1370`).  True: bf16 gradients cross NVLink as bf16 and
  are widened / accumulated in fp32 by the owner, ScaleGradients runs on the sender;
  False: the sender widens to fp32 (2x the wire bytes) and the owner scales.
"""
from .consts import RUN_OPTIONS, RUN_OPTION_ALIASES

_PROTOCOLS = ("grpc", "grpc+verbs", "grpc+gdr", "grpc+mpi", "nvlink", "nccl")


class PSConfig(object):
    def __init__(self,
                 protocol='grpc',
                 replicate_variables=True,
                 local_aggregation=True,
                 boundary_among_servers=True,
                 boundary_between_workers_and_servers=True):
        assert protocol in _PROTOCOLS, \
            'protocol must be one of %s' % (_PROTOCOLS,)
        self.protocol = protocol
        self.replicate_variables = replicate_variables
        self.local_aggregation = local_aggregation
        self.boundary_among_servers = boundary_among_servers
        self.boundary_between_workers_and_servers = \
            boundary_between_workers_and_servers

    def __repr__(self):
        return "PSConfig(%s)" % ", ".join(
            "%s=%r" % kv for kv in sorted(vars(self).items()))


class MPIConfig(object):
    """Launcher options for the AR path.

    The reference forwards ``mpirun_options`` to ``mpirun``
    (`mpi/runner.py:87-110`).  There is no mpirun here; the options string is
    kept, parsed for ``-x NAME=VALUE`` pairs, and those are exported into the
    environment of every spawned worker — the part of mpirun's behaviour
    user scripts relied on.
    """

    def __init__(self, mpirun_options='', use_allgatherv=None):
        self.mpirun_options = self.parse_mpirun_options(mpirun_options)
        # The reference's docs still show `MPIConfig(use_allgatherv=…)` although the
        # option was removed from its code (`doc/parallax_api.md:62,81` vs
        # `common/config.py:51-53`).  Accepted so scripts written from the docs run;
        # the AR route's sparse aggregation is always a variable-length all-gather.
        self.use_allgatherv = use_allgatherv

    def parse_mpirun_options(self, mpirun_options):
        if isinstance(mpirun_options, str):
            return mpirun_options
        elif isinstance(mpirun_options, (list, tuple)):
            return ' '.join([str(option) for option in mpirun_options])
        else:
            assert False, \
                'mpirun_options should be a string or a list of strings'

    def exported_env(self):
        """``-x A=B`` pairs found in the options string."""
        env = {}
        toks = self.mpirun_options.split()
        i = 0
        while i < len(toks):
            if toks[i] == '-x' and i + 1 < len(toks):
                if '=' in toks[i + 1]:
                    k, v = toks[i + 1].split('=', 1)
                    env[k] = v
                i += 2
            else:
                i += 1
        return env

    def __repr__(self):
        return "MPIConfig(mpirun_options=%r)" % (self.mpirun_options,)


class CommunicationConfig(object):
    def __init__(self, ps_config=None, mpi_config=None):
        ps_config = PSConfig() if ps_config is None else ps_config
        mpi_config = MPIConfig() if mpi_config is None else mpi_config
        assert isinstance(ps_config, PSConfig)
        assert isinstance(mpi_config, MPIConfig)
        self.ps_config = ps_config
        self.mpi_config = mpi_config


class CheckPointConfig(object):
    def __init__(self, ckpt_dir=None, save_ckpt_steps=None,
                 save_ckpt_secs=None):
        self.ckpt_dir = ckpt_dir
        self.save_ckpt_steps = save_ckpt_steps
        self.save_ckpt_secs = save_ckpt_secs


class ProfileConfig(object):
    def __init__(self, profile_dir=None, profile_steps=None,
                 profile_range=None, profile_worker=None):
        # steps and range are mutually exclusive
        # (reference `session_context.py:126`).
        assert profile_steps is None or profile_range is None, \
            'profile_steps and profile_range are mutually exclusive'
        if profile_range is not None:
            assert len(profile_range) == 2 and \
                profile_range[0] <= profile_range[1]
        self.profile_dir = profile_dir
        self.profile_steps = profile_steps
        self.profile_range = profile_range
        self.profile_worker = profile_worker


class ParallaxConfig(object):
    def __init__(self,
                 run_option='HYBRID',
                 average_sparse=False,
                 sess_config=None,
                 redirect_path=None,
                 search_partitions=True,
                 export_graph_path=None,
                 communication_config=None,
                 ckpt_config=None,
                 profile_config=None):
        """See module docstring.  ``sess_config`` is a dict of engine options
        (the analogue of ``tf.ConfigProto``): ``compute_dtype``,
        ``wire_dtype``, ``bucket_bytes``, ``cuda_graph``, ``fabric``,
        ``allreduce_algo`` …  ``export_graph_path`` dumps the analysis
        report (variable → dense/sparse, owner, bucket, partitions) per
        worker, the analogue of the transformed-MetaGraph dump
        (`common/lib.py:258-264`)."""
        self.run_option = run_option
        self.average_sparse = average_sparse
        self.sess_config = sess_config
        self.redirect_path = redirect_path
        self.search_partitions = search_partitions
        self.export_graph_path = export_graph_path
        self.communication_config = CommunicationConfig() \
            if communication_config is None else communication_config
        self.ckpt_config = CheckPointConfig() \
            if ckpt_config is None else ckpt_config
        self.profile_config = ProfileConfig() \
            if profile_config is None else profile_config
        self._sync = None
        self._resource_info = None

    # -- helpers ------------------------------------------------------------
    def normalized_run_option(self):
        opt = self.run_option
        if isinstance(opt, (tuple, list)) and len(opt) == 1:
            # the reference docs contain `cfg.run_option = run_option,`
            opt = opt[0]
        opt = str(opt).upper()
        opt = RUN_OPTION_ALIASES.get(opt, opt)
        if opt not in RUN_OPTIONS:
            raise ValueError('run_option must be one of %s (got %r)'
                             % (RUN_OPTIONS, self.run_option))
        return opt

    def get_ckpt_config(self):
        return self.ckpt_config

    def set_sync(self, sync):
        self._sync = sync

    @property
    def sync(self):
        return self._sync

    def set_resource_info(self, resource_info):
        self._resource_info = resource_info

    @property
    def resource_info(self):
        return self._resource_info

    def sess_option(self, key, default=None):
        if isinstance(self.sess_config, dict):
            return self.sess_config.get(key, default)
        return default
