"""Kernel-time breakdown of one training step with torch.profiler (CUPTI).
Usage: python tools/profile_step.py [--model lm1b|resnet50] [--graph]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity

import parallax_b200 as parallax
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="lm1b")
ap.add_argument("--graph", action="store_true")
ap.add_argument("--out", default="gpurun_out/profile_step.txt")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--trace", default=None, help="also write a per-stream kernel timeline (text)")
a = ap.parse_args()

class A: pass
args = A(); args.small = False; args.batch = None
builder = bench.build_lm1b if a.model == "lm1b" else bench.build_resnet
graph, make_batch, desc, metric, unit, baseline = builder(args, parallax, torch)
cfg = parallax.Config(run_option="HYBRID", search_partitions=False,
                      sess_config={"compute_dtype": "bf16", "cuda_graph": a.graph})
sess, *_ = parallax.parallel_run(graph, "localhost:0", sync=True, parallax_config=cfg)
eng = sess.engine
gen = torch.Generator().manual_seed(0)
batches = [{k: v.cuda() for k, v in make_batch(gen).items()} for _ in range(2)]
for i in range(5):
    eng.train_step(batches[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(a.steps):
        eng.train_step(batches[i % 2])
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45,
                                max_name_column_width=70)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
total = sum(e.device_time for e in prof.key_averages() if e.device_time > 0)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as f:
    f.write("model=%s graph=%s steps=%d\n" % (a.model, a.graph, a.steps))
    f.write(tab)
print(tab[-6000:])
if a.trace:
    # one line per kernel of the LAST profiled step: start (us, relative), duration, stream, name
    ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks.sort(key=lambda e: e.time_range.start)
    t_end = ks[-1].time_range.end
    per_step = (t_end - ks[0].time_range.start) / float(a.steps)
    t0 = t_end - per_step * 1.02
    with open(a.trace, "w") as f:
        f.write("# start_us dur_us stream name   (last step of %d; step ~%.0f us)\n"
                % (a.steps, per_step))
        for e in ks:
            if e.time_range.start < t0:
                continue
            stream = getattr(e, "stream", None)
            if stream is None:
                stream = getattr(e, "device_resource_id", -1)
            f.write("%9.1f %8.1f %4s %s\n" % (e.time_range.start - t0, e.time_range.elapsed_us(),
                                              stream, e.name[:90]))
sess.close()
