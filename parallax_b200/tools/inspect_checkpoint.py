"""Inspect / convert Parallax checkpoints.

    python -m parallax_b200.tools.inspect_checkpoint <ckpt_dir | model.ckpt-N.pt | model.ckpt-N/>
    python -m parallax_b200.tools.inspect_checkpoint <ckpt> --tensor W_P
    python -m parallax_b200.tools.inspect_checkpoint <ckpt> --to_state_dict out.pt [--ema]

The counterpart of TensorFlow's `inspect_checkpoint` for the reference's
`tf.train.Saver` files (`common/lib.py:38-56`): lists every logical variable
(dense master weights, optimizer slots, EMA shadows, sparse tables and their
slots) with shape / dtype / bytes, prints one tensor, or writes a plain
``name → tensor`` state_dict loadable into the single-device model
(`--ema` substitutes the EMA shadows, like `lm1b_eval.py:96-104`).  Sharded
checkpoints of the NVLink fabric (a directory with a manifest and one file per
owner) are assembled offline, without an engine or a GPU.
"""
import argparse
import os
import sys

import torch

from .. import checkpoint as _ckpt


_is_sharded = _ckpt.is_sharded


def load(path, max_table_bytes=None):
    """(resolved path, logical state dict).  `path`: a checkpoint directory (its latest
    checkpoint is taken), a single-file checkpoint, or a sharded checkpoint directory
    (`model.ckpt-N/` with a manifest and one shard file per owner), which is assembled
    offline — tables larger than `max_table_bytes` are only listed."""
    if os.path.isdir(path) and not _is_sharded(path):
        found = _ckpt.latest_checkpoint(path)
        if found is None:
            raise FileNotFoundError("no checkpoint in %s" % path)
        path = found
    return path, _ckpt.load_logical(path, max_table_bytes)


def entries(sd):
    """→ list of (kind, name, tensor)"""
    out = []
    dense = sd.get("dense") or {}
    for n, t in sorted((dense.get("master") or {}).items()):
        out.append(("dense", n, t))
    for n, slots in sorted((dense.get("slots") or {}).items()):
        for i, t in enumerate(slots):
            out.append(("dense-slot%d" % i, n, t))
    for n, t in sorted((dense.get("ema") or {}).items()):
        out.append(("dense-ema", n, t))
    for n, d in sorted((sd.get("sparse") or {}).items()):
        out.append(("sparse", n, d["weight"]))
        for i, t in enumerate(d.get("slots") or []):
            out.append(("sparse-slot%d" % i, n, t))
    for n, t in sorted((sd.get("buffers") or {}).items()):
        out.append(("buffer", n, t))
    return out


def to_state_dict(sd, use_ema=False):
    """plain state_dict of the single-device model (weights only)"""
    out = {}
    dense = sd.get("dense") or {}
    for n, t in (dense.get("master") or {}).items():
        out[n] = t
    if use_ema:
        for n, t in (dense.get("ema") or {}).items():
            out[n] = t
    for n, d in (sd.get("sparse") or {}).items():
        out[n] = d["weight"]
    for n, t in (sd.get("buffers") or {}).items():
        out[n] = t
    return out


def summarize(sd, out=sys.stdout):
    total = 0
    out.write("global_step: %d\n" % int(sd.get("global_step", 0)))
    out.write("%-14s %-44s %-22s %-10s %12s\n" % ("kind", "name", "shape", "dtype", "bytes"))
    for kind, name, t in entries(sd):
        nbytes = t.numel() * t.element_size()
        total += nbytes
        out.write("%-14s %-44s %-22s %-10s %12d\n" % (
            kind, name, "x".join(map(str, t.shape)) or "scalar",
            str(t.dtype).replace("torch.", ""), nbytes))
    out.write("total: %.2f MiB in %d tensors\n" % (total / 2 ** 20, len(entries(sd))))
    for name in sd.get("skipped", []):
        out.write("not assembled (larger than --max_table_gib): %s\n" % name)
    return total


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("checkpoint", help="checkpoint file or directory")
    ap.add_argument("--tensor", default=None, help="print this variable's values")
    ap.add_argument("--to_state_dict", default=None, help="write a plain state_dict here")
    ap.add_argument("--ema", action="store_true", help="use EMA shadows where they exist")
    ap.add_argument("--max_table_gib", type=float, default=8.0,
                    help="sharded checkpoints: do not assemble tables larger than this")
    a = ap.parse_args(argv)
    path, sd = load(a.checkpoint, int(a.max_table_gib * 2 ** 30))
    print("checkpoint:", path + (" (sharded: %d sparse variable(s) assembled from their owners' "
                                 "files)" % len(sd["sparse"]) if _is_sharded(path) else ""))
    if a.tensor:
        hits = [(k, n, t) for k, n, t in entries(sd) if n == a.tensor]
        if not hits:
            raise SystemExit("no variable named %r" % a.tensor)
        for kind, name, t in hits:
            print("%s %s %s" % (kind, name, tuple(t.shape)))
            print(t)
    elif a.to_state_dict:
        torch.save(to_state_dict(sd, a.ema), a.to_state_dict)
        print("wrote", a.to_state_dict)
    else:
        summarize(sd)
    return 0


if __name__ == "__main__":
    sys.exit(main())
