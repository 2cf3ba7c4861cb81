"""Image preprocessing for the CNN benchmarks.

Parity: `examples/tf_cnn_benchmarks/preprocessing.py:33-821` —
`parse_example_proto` (ImageNet `tf.Example`: encoded JPEG, label, boxes),
`train_image` (random crop from a distorted bounding box: area ∈ [0.05, 1],
aspect ∈ [3/4, 4/3], ≥ 10 % of the object covered, ≤ 100 attempts; resize with
the chosen method — ``round_robin`` cycles through them by batch position;
random flip; optional colour distortion in two orders), `eval_image` (87.5 %
central crop + bilinear resize), pixel range [-1, 1];
`RecordInputImagePreprocessor` (records are read through `parallax.shard.shard`,
`:517`), `Cifar10ImagePreprocessor` (pad-4 random crop + flip) and
`SyntheticImagePreprocessor` (`:704-750`, random images on the device).

Records come from the native loader (`utils/dataloader.py`: reader threads,
shuffle pool, record sharding by the worker's `(num_shards, shard_id)`); JPEG
decode + augmentation run in a Python thread pool (PIL releases the GIL) and
batches are assembled in pinned memory for the engine's asynchronous H2D copy.
"""
import io
import math
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from ...utils import dataloader as dl

RESIZE_METHODS = ("round_robin", "nearest", "bilinear", "bicubic", "area")


def parse_example_proto(example_serialized):
    """→ (image_buffer bytes, label int, bbox float [N,4] as ymin,xmin,ymax,xmax, text)"""
    ex = dl.parse_example(example_serialized)
    box = lambda k: ex.get("image/object/bbox/" + k, [])
    bbox = np.array(list(zip(box("ymin"), box("xmin"), box("ymax"), box("xmax"))),
                    dtype=np.float32).reshape(-1, 4)
    text = ex.get("image/class/text", [b""])
    return ex["image/encoded"][0], int(ex["image/class/label"][0]), bbox, \
        (text[0] if text else b"").decode("utf-8", "replace")


def _pil():
    from PIL import Image
    return Image


def _resample(method, batch_position=0):
    Image = _pil()
    table = {"nearest": Image.NEAREST, "bilinear": Image.BILINEAR, "bicubic": Image.BICUBIC,
             "area": Image.BOX}
    if method == "round_robin":
        order = ("bilinear", "nearest", "bicubic", "area")
        return table[order[batch_position % len(order)]]
    if method not in table:
        raise ValueError("invalid resize method %r (one of %s)" % (method, RESIZE_METHODS))
    return table[method]


def decode_jpeg(image_buffer):
    return _pil().open(io.BytesIO(image_buffer)).convert("RGB")


def sample_distorted_bounding_box(width, height, bbox, rng, min_object_covered=0.1,
                                  aspect_ratio_range=(0.75, 1.33), area_range=(0.05, 1.0),
                                  max_attempts=100):
    """→ (left, top, right, bottom) in pixels; the whole image when no box satisfies the
    constraints within `max_attempts`"""
    if bbox is None or len(bbox) == 0:
        bbox = np.array([[0.0, 0.0, 1.0, 1.0]], dtype=np.float32)
    ymin, xmin, ymax, xmax = bbox[rng.randrange(len(bbox))]
    obj = (xmin * width, ymin * height, xmax * width, ymax * height)
    obj_area = max((obj[2] - obj[0]) * (obj[3] - obj[1]), 1e-6)
    for _ in range(max_attempts):
        area = rng.uniform(*area_range) * width * height
        ar = rng.uniform(*aspect_ratio_range)
        w, h = int(round(math.sqrt(area * ar))), int(round(math.sqrt(area / ar)))
        if w < 1 or h < 1 or w > width or h > height:
            continue
        left, top = rng.randint(0, width - w), rng.randint(0, height - h)
        iw = min(left + w, obj[2]) - max(left, obj[0])
        ih = min(top + h, obj[3]) - max(top, obj[1])
        if iw > 0 and ih > 0 and iw * ih >= min_object_covered * obj_area:
            return left, top, left + w, top + h
    return 0, 0, width, height


def distort_color(img, rng, batch_position=0):
    """brightness ±32/255, saturation ×[0.5,1.5], hue ±0.2, contrast ×[0.5,1.5]; odd
    batch positions apply them in the second order of `distort_color` (`:386-437`)"""
    from PIL import ImageEnhance
    b = lambda im: ImageEnhance.Brightness(im).enhance(1.0 + rng.uniform(-32.0, 32.0) / 255.0)
    s = lambda im: ImageEnhance.Color(im).enhance(rng.uniform(0.5, 1.5))
    c = lambda im: ImageEnhance.Contrast(im).enhance(rng.uniform(0.5, 1.5))

    def h(im):
        hsv = np.array(im.convert("HSV"), dtype=np.int16)
        hsv[..., 0] = (hsv[..., 0] + int(rng.uniform(-0.2, 0.2) * 255)) % 256
        return _pil().fromarray(hsv.astype(np.uint8), "HSV").convert("RGB")
    order = (b, s, h, c) if batch_position % 2 == 0 else (b, c, s, h)
    for fn in order:
        img = fn(img)
    return img


def _to_tensor(img):
    """PIL RGB → float32 CHW in [-1, 1] (`images / 127.5 - 1`)"""
    a = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1)
    return a.to(torch.float32).div_(127.5).sub_(1.0)


def train_image(image_buffer, height, width, bbox, batch_position=0, resize_method="bilinear",
                distortions=True, rng=random):
    img = decode_jpeg(image_buffer)
    img = img.crop(sample_distorted_bounding_box(img.width, img.height, bbox, rng))
    img = img.resize((width, height), _resample(resize_method, batch_position))
    if rng.random() < 0.5:
        img = img.transpose(_pil().FLIP_LEFT_RIGHT)
    if distortions:
        img = distort_color(img, rng, batch_position)
    return _to_tensor(img)


def eval_image(image_buffer, height, width, batch_position=0, resize_method="bilinear"):
    img = decode_jpeg(image_buffer)
    w, h = img.size
    cw, ch = int(round(w * 0.875)), int(round(h * 0.875))
    left, top = (w - cw) // 2, (h - ch) // 2
    img = img.crop((left, top, left + cw, top + ch))
    return _to_tensor(img.resize((width, height), _resample(
        "bilinear" if resize_method == "round_robin" else resize_method, batch_position)))


class _Base(object):
    def __init__(self, height, width, batch_size, num_splits=1, dtype=torch.float32, train=True,
                 distortions=True, resize_method="bilinear", shift_ratio=0, seed=None,
                 num_threads=8, pin_memory=True):
        self.height, self.width, self.batch_size = height, width, batch_size
        self.num_splits, self.dtype, self.train = num_splits, dtype, train
        self.distortions, self.resize_method = distortions, resize_method
        self.shift_ratio, self.seed, self.num_threads = shift_ratio, seed, num_threads
        self.pin = pin_memory and torch.cuda.is_available()
        if batch_size % num_splits:
            raise ValueError("batch_size must be a multiple of num_splits: batch_size %d, "
                             "num_splits: %d" % (batch_size, num_splits))

    def _finish(self, images, labels):
        images = torch.stack(images).to(self.dtype)
        labels = torch.tensor(labels, dtype=torch.int64)
        if self.pin:
            images, labels = images.pin_memory(), labels.pin_memory()
        return images, labels


class RecordInputImagePreprocessor(_Base):
    """ImageNet TFRecords → (images [B,3,H,W] in [-1,1], labels [B] in 1..1000)"""

    def preprocess(self, image_buffer, bbox, batch_position, rng):
        if self.train:
            return train_image(image_buffer, self.height, self.width, bbox, batch_position,
                               self.resize_method, self.distortions, rng)
        return eval_image(image_buffer, self.height, self.width, batch_position,
                          self.resize_method)

    def parse_and_preprocess(self, value, batch_position, rng=random):
        buf, label, bbox, _ = parse_example_proto(value)
        return self.preprocess(buf, bbox, batch_position, rng), label

    def minibatch(self, dataset, subset, epochs=None, capacity=10000):
        """generator of (images, labels); training reads forever unless `epochs`"""
        loader = dl.RecordLoader(dataset.files(subset), dl.TFRECORD, shuffle=self.train,
                                 capacity=capacity, seed=self.seed or 301,
                                 epochs=(epochs or 0) if self.train else (epochs or 1),
                                 shard="record", num_threads=4)
        rngs = [random.Random((self.seed or 0) * 7919 + i) for i in range(self.batch_size)]
        with ThreadPoolExecutor(self.num_threads) as pool:
            while True:
                recs = loader.next_batch(self.batch_size)
                if len(recs) < self.batch_size and (self.train or not recs):
                    loader.close()
                    return
                out = list(pool.map(lambda a: self.parse_and_preprocess(a[1], a[0], rngs[a[0]]),
                                    enumerate(recs)))
                yield self._finish([o[0] for o in out], [o[1] for o in out])


class Cifar10ImagePreprocessor(_Base):
    """python-pickle CIFAR-10 → (images [B,3,32,32] in [-1,1], labels [B])"""

    def _distort_image(self, img, rng):
        padded = torch.nn.functional.pad(img, (4, 4, 4, 4))
        top, left = rng.randint(0, 8), rng.randint(0, 8)
        img = padded[:, top:top + 32, left:left + 32]
        return img.flip(-1) if rng.random() < 0.5 else img

    def preprocess(self, raw, rng=random):
        img = torch.as_tensor(raw).to(torch.float32)
        if self.train and self.distortions:
            img = self._distort_image(img, rng)
        return img.div(127.5).sub(1.0)

    def minibatch(self, dataset, subset, epochs=None):
        from ... import shard as _shard
        images, labels = dataset.read_data_files(subset)
        index = _shard.shard(list(range(len(labels))))      # this worker's examples
        rng = random.Random(self.seed or 0)
        epoch = 0
        while epochs is None or epoch < epochs:
            order = list(index)
            if self.train:
                rng.shuffle(order)
            for s in range(0, len(order) - self.batch_size + 1, self.batch_size):
                sel = order[s:s + self.batch_size]
                yield self._finish([self.preprocess(images[i], rng) for i in sel],
                                   [int(labels[i]) for i in sel])
            epoch += 1
            if not self.train:
                return


class SyntheticImagePreprocessor(_Base):
    """ONE random batch reused every step, created on `device` (`:704-750`: a
    truncated-normal image variable and uniform random labels)"""

    def minibatch(self, dataset, subset, epochs=None, device=None):
        g = torch.Generator().manual_seed(self.seed or 0)
        images = torch.nn.init.trunc_normal_(
            torch.empty(self.batch_size, 3, self.height, self.width), mean=0.0, std=60.0 / 127.5,
            a=-1.0, b=1.0, generator=g).to(self.dtype)
        labels = torch.randint(1, dataset.num_classes + 1, (self.batch_size,), generator=g)
        if device is not None:
            images, labels = images.to(device), labels.to(device)
        elif self.pin:
            images, labels = images.pin_memory(), labels.pin_memory()
        n = 0
        while epochs is None or n < epochs:
            yield images, labels
            n += 1
