"""The example model families on N real GPUs (torchrun): the GNMT NMT model and the CNN
harness reproduce the host-fabric (gloo) run of the same N-rank job, skip-thoughts trains in
bf16 with three lookups per step into one partitioned table.  Launched by
tests/test_multigpu.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import parallax_b200 as parallax


def nmt_losses(fabric, dtype, steps, rank):
    import parallax_b200.models.nmt as nmt
    torch.manual_seed(0)
    hp = nmt.create_hparams(num_units=32, num_layers=2, encoder_type="gnmt",
                            attention="normed_bahdanau", attention_architecture="gnmt_v2",
                            residual=True, dropout=0.0, num_embeddings_partitions=3,
                            learning_rate=0.5)
    nmt.extend_hparams(hp, 40, 40)
    m = nmt.create_model(hp)
    sc = {"fabric": fabric}
    if dtype:
        sc["compute_dtype"] = dtype
    sess, *_ = parallax.parallel_run(nmt.nmt_graph(m, hp), "localhost",
                                     parallax_config=parallax.Config(search_partitions=False,
                                                                     sess_config=sc))
    g = torch.Generator().manual_seed(1 + rank)
    losses = []
    for s in range(steps):
        B, S, T = 8, 5 + (s + rank) % 3, 4 + (s + 2 * rank) % 3      # variable lengths
        feed = {"source": [torch.randint(3, 40, (B, S), generator=g)],
                "target_input": [torch.randint(3, 40, (B, T), generator=g)],
                "target_output": [torch.randint(3, 40, (B, T), generator=g)],
                "source_sequence_length": [torch.randint(1, S + 1, (B,), generator=g)],
                "target_sequence_length": [torch.randint(1, T + 1, (B,), generator=g)]}
        losses.append(sess.run(["loss", "train_op"], feed)[0][0])
    sess.close()
    return losses


def skip_thoughts_losses(rank):
    from parallax_b200.models import skip_thoughts as st
    from parallax_b200.models.skip_thoughts.input_ops import parse_example_batch
    torch.manual_seed(0)
    mc = st.model_config(vocab_size=48, word_embedding_dim=16, encoder_dim=32, batch_size=4,
                         num_embedding_partitions=3, bidirectional_encoder=True)
    tc = st.training_config(learning_rate=0.01)
    model = st.SkipThoughtsModel(mc)
    sess, *_ = parallax.parallel_run(
        st.skip_thoughts_graph(model, tc), "localhost",
        parallax_config=parallax.Config(search_partitions=False,
                                        sess_config={"fabric": "nvlink", "compute_dtype": "bf16"}))
    r = rank % 3
    batch = parse_example_batch([([3 + r, 4, 5, 0], [6, 7, 0], [8, 0]), ([9, 0], [3, 0], [4, 5, 6, 0]),
                                 ([10, 11, 0], [12 + r, 0], [13, 14, 0]), ([5, 0], [6, 0], [7, 0])])
    losses = [sess.run(["loss", "train_op"], st.feed_from_batch(batch))[0][0] for _ in range(15)]
    sess.close()
    return losses


def cnn_harness():
    from parallax_b200.models import cnn_benchmarks as cb
    bench = cb.BenchmarkCNN(cb.make_params(model="lenet", batch_size=16, num_batches=8,
                                           num_warmup_batches=4, display_every=4, use_fp16=True,
                                           optimizer="momentum", learning_rate=0.01))
    cfg = parallax.Config(run_option="MPI", search_partitions=False,
                          sess_config=dict(bench.sess_config(), fabric="nvlink"))
    sess, nw, wid, _ = parallax.parallel_run(bench.build_graph(), "localhost", parallax_config=cfg)
    res = bench.run(sess, nw, wid)
    captured = bool(getattr(sess.engine, "graph_captured", False))
    sess.close()
    return res, captured


def main():
    from parallax_b200.parallel.fabric import Comm
    comm = Comm.from_env()
    world, rank = comm.world, comm.rank
    ok = True

    def check(name, cond):
        nonlocal ok
        flags = comm.all_gather_object(bool(cond))
        if rank == 0:
            print("%-70s %s" % (name, "OK" if all(flags) else "FAIL %s" % flags), flush=True)
        ok = ok and all(flags)
    ref = nmt_losses("host", None, 6, rank)
    nv = nmt_losses("nvlink", None, 6, rank)
    check("NMT (GNMT v2, variable lengths) nvlink == host fabric, %d ranks" % world,
          np.allclose(nv, ref, rtol=3e-3, atol=3e-3))
    bf = nmt_losses("nvlink", "bf16", 12, rank)
    check("NMT bf16 trains", np.isfinite(bf).all())
    st_l = skip_thoughts_losses(rank)
    check("skip-thoughts bf16 trains (3 lookups / step / table)",
          np.isfinite(st_l).all() and st_l[-1] < st_l[0])
    res, captured = cnn_harness()
    check("CNN benchmark harness (lenet, AR, CUDA graph)", res["num_steps"] == 8 and
          np.isfinite(res["average_loss"]) and captured)
    if rank == 0:
        print("ALL OK" if ok else "SOME FAILED", flush=True)
    comm.shutdown()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
