"""tf_cnn_benchmarks parity: parameter table, learning-rate rules, datasets
(ImageNet TFRecords, CIFAR-10 pickles, synthetic), image preprocessing, per-model
defaults / registry, and the `BenchmarkCNN` train + eval loops on the host fabric."""
import io
import os
import pickle
import random

import numpy as np
import pytest
import torch

import parallax_b200 as parallax
from parallax_b200.models import cnn, cnn_benchmarks as cb
from parallax_b200.models.cnn_benchmarks import datasets, model_config, preprocessing
from parallax_b200.utils import dataloader as dl


def _jpeg(w, h, color):
    from PIL import Image
    buf = io.BytesIO()
    Image.new("RGB", (w, h), color).save(buf, "JPEG")
    return buf.getvalue()


def _imagenet_fixture(d, n=24, shards=2):
    os.makedirs(d, exist_ok=True)
    for subset in ("train", "validation"):
        for s in range(shards):
            with dl.TFRecordWriter(os.path.join(d, "%s-%05d-of-%05d" % (subset, s, shards))) as w:
                for i in range(n // shards):
                    label = 1 + (i + s) % 3                     # colour encodes the class
                    color = [(250, 10, 10), (10, 250, 10), (10, 10, 250)][label - 1]
                    w.write(dl.encode_example({
                        "image/encoded": _jpeg(48 + i, 40, color), "image/class/label": label,
                        "image/class/text": "c%d" % label,
                        "image/object/bbox/xmin": [0.2], "image/object/bbox/ymin": [0.2],
                        "image/object/bbox/xmax": [0.8], "image/object/bbox/ymax": [0.9]}))
    return d


def _cifar_fixture(d, per_batch=20):
    os.makedirs(d, exist_ok=True)
    rng = np.random.RandomState(0)
    for name in ["data_batch_%d" % i for i in range(1, 6)] + ["test_batch"]:
        labels = rng.randint(0, 10, per_batch)
        data = np.zeros((per_batch, 3, 32, 32), np.uint8)
        for i, l in enumerate(labels):                          # brightness encodes the class
            data[i] = 20 * l + rng.randint(0, 10, (3, 32, 32))
        with open(os.path.join(d, name), "wb") as f:
            pickle.dump({b"data": data.reshape(per_batch, -1), b"labels": labels.tolist()}, f)
    return d


# ------------------------------------------------------------------- params
def test_params_table_and_validation():
    p = cb.make_params(model="resnet50", optimizer="momentum")
    assert p.weight_decay == 0.00004 and p.momentum == 0.9 and p.rmsprop_epsilon == 1.0
    assert p.batch_size == 0 and p.resize_method == "bilinear" and not p.eval
    with pytest.raises(ValueError, match="Invalid parameter"):
        cb.make_params(no_such_flag=1)
    with pytest.raises(ValueError, match="forward_only and eval"):
        cb.BenchmarkCNN(cb.make_params(eval=True, forward_only=True))
    with pytest.raises(ValueError, match="not recognized"):
        cb.BenchmarkCNN(cb.make_params(optimizer="adam"))
    with pytest.raises(ValueError, match="Invalid model name"):
        cb.BenchmarkCNN(cb.make_params(model="resnet20"))               # CIFAR-only model
    with pytest.raises(ValueError, match="requires use_fp16"):
        cb.BenchmarkCNN(cb.make_params(fp16_loss_scale=128.0))
    import argparse
    flags = cb.benchmark_cnn.add_arguments(argparse.ArgumentParser()).parse_args(
        ["--model", "vgg16", "--use_fp16", "--distortions", "false", "--batch_size", "7"])
    q = cb.benchmark_cnn.make_params_from_flags(flags)
    assert (q.model, q.use_fp16, q.distortions, q.batch_size) == ("vgg16", True, False, 7)
    b = cb.BenchmarkCNN(q)
    assert b.sess_config() == {"cuda_graph": True, "compute_dtype": "bf16"} and b.batch_size == 7


def test_learning_rate_rules():
    ds = datasets.ImagenetData()
    conf = model_config.get_model_config("resnet50", ds)
    # no --learning_rate: the model's schedule (Goyal: linear warm-up to 0.1·B/256, ÷10 steps)
    lr = cb.get_learning_rate(cb.make_params(model="resnet50"), 1281167, conf, 1024)
    spe = 1281167 / 1024.0
    assert abs(lr(int(5 * spe) + 1) - 0.4) < 1e-6 and lr(1) < 0.1 and abs(lr(int(31 * spe)) - 0.04) < 1e-6
    assert cb.get_learning_rate(cb.make_params(model="vgg16"), 1000,
                                model_config.get_model_config("vgg16", ds), 10) == 0.005
    # explicit rate with staircase decay per `num_epochs_per_decay` epochs and a floor
    p = cb.make_params(learning_rate=0.1, num_epochs_per_decay=2, learning_rate_decay_factor=0.5,
                       minimum_learning_rate=0.03)
    f = cb.get_learning_rate(p, 1000, conf, 100)          # 10 steps/epoch ⇒ decay every 20
    assert [f(1), f(20), f(21), f(41), f(61)] == [0.1, 0.1, 0.05, 0.03, 0.03]
    assert cb.get_learning_rate(cb.make_params(learning_rate=0.2), 1000, conf, 100) == 0.2
    with pytest.raises(ValueError):
        cb.get_learning_rate(cb.make_params(num_epochs_per_decay=2), 1000, conf, 100)


def test_model_config_registry_and_defaults():
    inet, cifar = datasets.ImagenetData(), datasets.Cifar10Data()
    a = model_config.get_model_config("alexnet", inet)
    assert (a.get_image_size(), a.get_default_batch_size(), a.learning_rate) == (224, 512, 0.005)
    c = model_config.get_model_config("alexnet", cifar)
    assert (c.get_image_size(), c.get_default_batch_size(), c.learning_rate) == (32, 128, 0.1)
    assert isinstance(c.build(11), cnn.AlexNetCifar)
    assert model_config.get_model_config("resnet56_v2", cifar).get_default_batch_size() == 128
    for name in ("vgg11", "inception4", "resnet152_v2", "overfeat", "googlenet", "lenet"):
        assert model_config.get_model_config(name, inet).name
    with pytest.raises(ValueError):
        model_config.get_model_config("densenet40_k12", inet)
    with pytest.raises(ValueError):
        model_config._get_model_map("mnist")
    model_config.register_model("tiny_test_model", "cifar10",
                                lambda: model_config.ModelConfig("tiny", "trivial_cifar", 32, 4, 0.01))
    assert model_config.get_model_config("tiny_test_model", cifar).get_default_batch_size() == 4
    with pytest.raises(ValueError, match="already registered"):
        model_config.register_model("tiny_test_model", "cifar10", None)


# ------------------------------------------------------------------ datasets
def test_datasets(tmp_path):
    assert datasets.create_dataset(None, None).name == "imagenet"
    assert datasets.create_dataset(None, None).use_synthetic_gpu_images()
    assert datasets.create_dataset("/x/cifar10_data", None).name == "cifar10"
    with pytest.raises(ValueError, match="Could not identify"):
        datasets.create_dataset("/x/unknown", None)
    with pytest.raises(ValueError, match="Unknown dataset"):
        datasets.create_dataset("/x", "mnist")
    inet = datasets.ImagenetData(str(tmp_path))
    assert inet.num_examples_per_epoch("train") == 1281167 and inet.num_classes == 1000
    with pytest.raises(ValueError):
        inet.num_examples_per_epoch("test")
    with pytest.raises(ValueError, match="no train records"):
        inet.files("train")
    c = datasets.Cifar10Data(_cifar_fixture(str(tmp_path / "cifar")))
    x, y = c.read_data_files("train")
    assert x.shape == (100, 3, 32, 32) and y.shape == (100,) and x.dtype == np.uint8
    assert c.read_data_files("validation")[0].shape[0] == 20 and c.num_classes == 10


# ------------------------------------------------------------- preprocessing
def test_distorted_bounding_box_constraints():
    rng = random.Random(0)
    bbox = np.array([[0.2, 0.3, 0.8, 0.9]], np.float32)
    for _ in range(200):
        l, t, r, b = preprocessing.sample_distorted_bounding_box(200, 100, bbox, rng)
        w, h = r - l, b - t
        assert 0 <= l < r <= 200 and 0 <= t < b <= 100
        if (w, h) != (200, 100):
            assert 0.05 * 20000 * 0.9 <= w * h <= 20000 and 0.70 <= w / h <= 1.40
            iw = min(r, 180) - max(l, 60)
            ih = min(b, 80) - max(t, 20)
            assert iw > 0 and ih > 0 and iw * ih >= 0.1 * 120 * 60 - 1
    # impossible constraints fall back to the whole image
    assert preprocessing.sample_distorted_bounding_box(
        10, 10, bbox, rng, area_range=(4.0, 5.0)) == (0, 0, 10, 10)


def test_image_functions():
    buf = _jpeg(64, 48, (255, 0, 0))
    tr = preprocessing.train_image(buf, 24, 32, None, rng=random.Random(1))
    assert tr.shape == (3, 24, 32) and tr.dtype == torch.float32
    assert -1.0 <= float(tr.min()) and float(tr.max()) <= 1.0
    ev = preprocessing.eval_image(buf, 24, 24)
    assert ev.shape == (3, 24, 24) and float(ev[0].mean()) > 0.9 and float(ev[1].mean()) < -0.9
    plain = preprocessing.train_image(buf, 24, 24, None, distortions=False, rng=random.Random(1))
    assert float(plain[0].mean()) > 0.9                     # no colour distortion: still red
    for i in range(4):
        preprocessing.train_image(buf, 16, 16, None, batch_position=i, resize_method="round_robin",
                                  rng=random.Random(i))
    with pytest.raises(ValueError):
        preprocessing.train_image(buf, 16, 16, None, resize_method="lanczos")
    raw = dl.encode_example({"image/encoded": buf, "image/class/label": 7,
                             "image/class/text": "cat", "image/object/bbox/xmin": [0.1, 0.2],
                             "image/object/bbox/ymin": [0.0, 0.1], "image/object/bbox/xmax": [0.9, 1.0],
                             "image/object/bbox/ymax": [1.0, 0.8]})
    b, label, bbox, text = preprocessing.parse_example_proto(raw)
    assert (b, label, text) == (buf, 7, "cat") and bbox.shape == (2, 4)
    assert np.allclose(bbox[1], [0.1, 0.2, 0.8, 1.0])       # ymin, xmin, ymax, xmax
    with pytest.raises(ValueError, match="multiple of num_splits"):
        preprocessing.SyntheticImagePreprocessor(8, 8, 6, num_splits=4)


def test_record_preprocessor_shards_records(tmp_path):
    inet = datasets.ImagenetData(_imagenet_fixture(str(tmp_path)))
    pre = preprocessing.RecordInputImagePreprocessor(16, 16, 4, train=False, pin_memory=False)
    parallax.shard.update_shard_values_for_worker(2, 1, 1)       # worker 1 of 2
    batches = list(pre.minibatch(inet, "validation"))
    assert len(batches) == 3 and batches[0][0].shape == (4, 3, 16, 16)
    labels = torch.cat([b[1] for b in batches]).tolist()
    all_labels = [dl.parse_example(r)["image/class/label"][0]
                  for r in dl.RecordLoader(inet.files("validation"), dl.TFRECORD, shard=None)]
    assert labels == all_labels[1::2]
    for img, lab in zip(batches[0][0], batches[0][1]):           # colour channel = class
        assert int(img.mean((1, 2)).argmax()) == int(lab) - 1


def test_cifar_preprocessor(tmp_path):
    c = datasets.Cifar10Data(_cifar_fixture(str(tmp_path)))
    pre = preprocessing.Cifar10ImagePreprocessor(32, 32, 10, train=True, seed=1, pin_memory=False)
    x, y = next(pre.minibatch(c, "train"))
    assert x.shape == (10, 3, 32, 32) and -1.0 <= float(x.min()) <= float(x.max()) <= 1.0
    ev = preprocessing.Cifar10ImagePreprocessor(32, 32, 10, train=False, pin_memory=False)
    got = list(ev.minibatch(c, "validation"))
    assert len(got) == 2
    raw, labels = c.read_data_files("validation")
    assert torch.allclose(got[0][0][3], torch.as_tensor(raw[3]).float() / 127.5 - 1.0)
    assert got[0][1].tolist() == labels[:10].tolist()


# ------------------------------------------------------------------ harness
def _session(bench, run_option="MPI", ckpt=None):
    cfg = parallax.Config(run_option=run_option, search_partitions=False,
                          sess_config=dict(bench.sess_config(), fabric="host"))
    if ckpt:
        cfg.ckpt_config = parallax.CheckPointConfig(ckpt_dir=ckpt, save_ckpt_steps=1000)
    return parallax.parallel_run(bench.build_graph(), "localhost", parallax_config=cfg)


@pytest.mark.parametrize("optimizer,extra", [
    ("momentum", {"gradient_clip": 0.5}), ("sgd", {"use_fp16": True, "fp16_loss_scale": 64.0}),
    ("rmsprop", {"weight_decay": 0.0})])
def test_benchmark_trains_on_synthetic(optimizer, extra):
    bench = cb.BenchmarkCNN(cb.make_params(
        model="lenet", batch_size=8, num_batches=12, num_warmup_batches=2, display_every=6,
        optimizer=optimizer, learning_rate=0.01, print_training_accuracy=True, params_stat=True,
        **extra))
    sess, nw, wid, _ = _session(bench)
    try:
        assert sess.engine.run_option == "MPI"          # dense only ⇒ AR, as in the reference
        first = sess.run("loss", {k: [v] for k, v in zip(("images", "labels"),
                                                         next(bench.input_iterator()))})[0]
        res = bench.run(sess, nw, wid)
        assert res["num_steps"] == 12 and res["images_per_sec"] > 0
        assert res["average_loss"] < first              # one fixed batch: the loss must fall
        if extra.get("fp16_loss_scale"):
            assert bench.build_graph().loss_scale == 64.0
    finally:
        sess.close()


def test_benchmark_cifar_train_then_eval_from_checkpoint(tmp_path):
    data = _cifar_fixture(str(tmp_path / "cifar10"), per_batch=40)
    ck = str(tmp_path / "ck")
    common = dict(model="alexnet", data_dir=data, batch_size=20, display_every=50,
                  tf_random_seed=3, deterministic=True)
    train = cb.BenchmarkCNN(cb.make_params(num_batches=60, num_warmup_batches=0, optimizer="sgd",
                                           learning_rate=0.01, distortions=False, **common))
    assert train.dataset.name == "cifar10" and train.model_conf.get_image_size() == 32
    sess, nw, wid, _ = _session(train, ckpt=ck)
    try:
        res = train.run(sess, nw, wid)
        sess.save_checkpoint()
    finally:
        sess.close()
    assert np.isfinite(res["average_loss"]) and res["average_loss"] < 2.6      # ≈ ln(11) at start
    ev = cb.BenchmarkCNN(cb.make_params(eval=True, checkpoint_dir=ck, num_batches_for_eval=2, **common))
    sess, nw, wid, _ = _session(ev, ckpt=ck)               # restore-on-start
    try:
        assert sess.engine.global_step == 60
        out = ev.run(sess, nw, wid)
    finally:
        sess.close()
    assert out["num_examples"] == 40 and 0.0 <= out["top_1_accuracy"] <= out["top_5_accuracy"] <= 1.0


def test_benchmark_imagenet_records_forward_only(tmp_path):
    data = _imagenet_fixture(str(tmp_path / "imagenet"), n=16)
    bench = cb.BenchmarkCNN(cb.make_params(model="lenet", data_dir=data, data_name="imagenet",
                                           batch_size=4, forward_only=True, num_batches=3,
                                           num_warmup_batches=1, display_every=3, deterministic=True))
    assert not bench.train and bench.build_graph().optimizer is None
    sess, nw, wid, _ = _session(bench)
    try:
        res = bench.run(sess, nw, wid)
        assert res["num_steps"] == 3 and sess.engine.global_step == 0      # nothing was updated
        x, y = next(bench.input_iterator())
        assert x.shape == (4, 3, 28, 28) and set(y.tolist()) <= {1, 2, 3}
    finally:
        sess.close()
