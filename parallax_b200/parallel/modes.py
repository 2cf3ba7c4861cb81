"""Run-option → data-path routing table.

Parity: reference mode selection `common/runner.py:88-128`, validation
`:159-164` (MPI and HYBRID require ``sync=True``) and
`ps/graph_transform.py:27-28` (``replicate_variables`` only with sync).

================  =======================  ================================
run option        dense variables           sparse variables
================  =======================  ================================
MPI (AR)          all-reduce (mean), every  all-gather of (indices, rows)
                  rank updates its replica  from every rank; table is
                                            replicated; every rank applies
PS                owner-reduce → owner      rows pushed to the owning rank,
                  applies → mirrors         owner accumulates + applies;
                  refreshed                 lookups read the owner's shard
HYBRID (default)  as MPI                    as PS
================  =======================  ================================

``sync=False`` is only legal for PS: no step barrier, every worker applies
its own gradient to the owner's copy as soon as it is computed (Hogwild —
reference `ps/between_graph_parallel.py:137-146`).
"""

DENSE_ALLREDUCE = "allreduce"
DENSE_OWNER = "owner"
SPARSE_ALLGATHER = "allgather"
SPARSE_OWNER = "owner"


class Route(object):
    def __init__(self, run_option, sync, dense, sparse):
        self.run_option = run_option
        self.sync = sync
        self.dense = dense
        self.sparse = sparse

    def __repr__(self):
        return "Route(%s, sync=%s, dense=%s, sparse=%s)" % (
            self.run_option, self.sync, self.dense, self.sparse)


def validate(run_option, sync, ps_config=None):
    if run_option in ("MPI", "HYBRID") and not sync:
        raise ValueError("%s requires sync=True (asynchronous training is "
                         "only supported by run_option='PS')" % run_option)
    if ps_config is not None and run_option == "PS" and not sync and \
            ps_config.replicate_variables:
        # the reference asserts; we downgrade with the documented meaning:
        # async workers pull owner values at step start.
        ps_config.replicate_variables = False


def route_for(run_option, sync):
    if run_option == "MPI":
        return Route("MPI", True, DENSE_ALLREDUCE, SPARSE_ALLGATHER)
    if run_option == "PS":
        return Route("PS", bool(sync), DENSE_OWNER, SPARSE_OWNER)
    if run_option == "HYBRID":
        return Route("HYBRID", True, DENSE_ALLREDUCE, SPARSE_OWNER)
    raise ValueError(run_option)
