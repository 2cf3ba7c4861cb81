"""User script used by tests/test_launcher.py: the SAME file is executed by the
launcher (master) and re-executed by every worker it spawns — the reference
contract (`common/runner.py:166,185-193`)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import parallax_b200 as parallax
from parallax_b200.models.simple import MLPWithEmbedding

out_dir, resource, run_option, steps, search = sys.argv[1:6]
torch.manual_seed(0)
part = parallax.get_partitioner(2)
model = MLPWithEmbedding(64, partitioner=part)
graph = parallax.Graph(model, optimizer=parallax.optim.Adagrad(0.1, 1.0))
cfg = parallax.Config(run_option=run_option, search_partitions=(search == "1"),
                      redirect_path=os.path.join(out_dir, "logs"),
                      export_graph_path=os.path.join(out_dir, "graph"))
sess, num_workers, worker_id, nrep = parallax.parallel_run(
    graph, resource, sync=True, parallax_config=cfg)
# ---- only workers get here ---------------------------------------------------------
ds = parallax.shard.shard(list(range(1000)))
first = list(ds)[:3]
g = torch.Generator().manual_seed(worker_id)
losses = []
for s in range(int(steps)):
    ids = torch.randint(0, 64, (4, 3), generator=g)
    labels = torch.randint(0, 4, (4,), generator=g)
    loss, gs, _ = sess.run(["loss", "global_step", "train_op"],
                           {"ids": [ids], "labels": [labels]})
    losses.append(loss[0])
with open(os.path.join(out_dir, "worker_%d_P%d.json" % (worker_id, part.num_partitions)), "w") as f:
    json.dump({"worker_id": worker_id, "num_workers": num_workers, "first": first,
               "global_step": gs[0], "partitions": part.num_partitions,
               "tableP": sess.engine.tables["emb.weight"].layout.P,
               "env_role": os.environ.get("PARALLAX_RUN_OPTION")}, f)
sess.close()
