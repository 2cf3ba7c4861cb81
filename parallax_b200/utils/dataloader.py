"""Python face of the native input pipeline (`ops/csrc/runtime/dataloader.cpp`).

* `RecordLoader` — N native reader threads over text-line or TFRecord files,
  bounded pool with optional `RandomShuffleQueue`-style shuffling
  (`capacity`, `min_after_dequeue`), sharding by file or by record with the
  `(num_shards, shard_id)` pair `parallax.shard` assigns to this worker.
* `TFRecordWriter` / `tfrecord_iterator` — the TFRecord framing
  (length, masked crc32c, payload, masked crc32c).
* `encode_example` / `parse_example` — the `tf.train.Example` wire format
  (bytes / float / int64 feature lists), enough to read and write the ImageNet
  records of `tf_cnn_benchmarks` (`preprocessing.py:33-120` `parse_example_proto`)
  without TensorFlow.
* `NativeVocab` — whitespace tokenisation + word→id lookup in native code
  (`lookup_ops.index_table_from_file`).
"""
import ctypes
import glob as _glob
import struct

import numpy as np

from .. import ops
from .. import shard as _shard

TEXT, TFRECORD = 0, 1

_c = ctypes
ops.register_signatures({
    "px_crc32c": (_c.c_uint32, [_c.c_void_p, _c.c_size_t]),
    "px_masked_crc32c": (_c.c_uint32, [_c.c_void_p, _c.c_size_t]),
    "px_loader_open": (_c.c_int, [_c.POINTER(_c.c_char_p), _c.c_int, _c.c_int, _c.c_int,
                                  _c.c_long, _c.c_long, _c.c_int, _c.c_ulonglong, _c.c_int,
                                  _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    "px_loader_next": (_c.c_int, [_c.c_int, _c.c_void_p, _c.c_size_t,
                                  _c.POINTER(_c.c_size_t)]),
    "px_loader_next_batch": (_c.c_int, [_c.c_int, _c.c_int, _c.c_void_p, _c.c_size_t,
                                        _c.POINTER(_c.c_size_t)]),
    "px_loader_stats": (_c.c_int, [_c.c_int, _c.POINTER(_c.c_long), _c.POINTER(_c.c_long),
                                   _c.POINTER(_c.c_long)]),
    "px_loader_error": (_c.c_int, [_c.c_int, _c.c_char_p, _c.c_size_t]),
    "px_loader_close": (_c.c_int, [_c.c_int]),
    "px_vocab_create": (_c.c_int, [_c.c_char_p, _c.c_size_t, _c.c_int]),
    "px_vocab_encode": (_c.c_int, [_c.c_int, _c.c_char_p, _c.c_size_t,
                                   _c.POINTER(_c.c_longlong), _c.c_int]),
    "px_vocab_free": (_c.c_int, [_c.c_int]),
})


def masked_crc32c(data):
    return int(ops.lib().px_masked_crc32c(data, len(data)))


def crc32c(data):
    return int(ops.lib().px_crc32c(data, len(data)))


def expand_files(file_pattern):
    """comma separated glob patterns → sorted file list"""
    if isinstance(file_pattern, (list, tuple)):
        return sorted(file_pattern)
    files = []
    for p in file_pattern.split(","):
        files.extend(_glob.glob(p))
    return sorted(files)


class RecordLoader(object):
    """Iterate the records (bytes) of `files`.

    shard="file"  : this worker reads files ``i % num_shards == shard_id``;
    shard="record": every worker scans all files and keeps records
                    ``j % num_shards == shard_id`` (`Dataset.shard` semantics);
    shard=None    : no sharding.  `num_shards`/`shard_id` default to the
    late-bound values of `parallax.shard`, read when iteration starts.
    """

    def __init__(self, files, kind=TEXT, num_threads=4, shuffle=False, capacity=4096,
                 min_after_dequeue=None, seed=0, epochs=1, shard="record", num_shards=None,
                 shard_id=None, verify_crc=True, max_record_bytes=1 << 20):
        self.files = expand_files(files)
        if not self.files:
            raise ValueError("Found no input files matching %s" % (files,))
        self.kind, self.num_threads, self.shuffle = kind, int(num_threads), bool(shuffle)
        self.capacity = int(capacity)
        self.min_after = int(0.6 * capacity if min_after_dequeue is None
                             else min_after_dequeue)
        self.seed, self.epochs, self.shard = int(seed), int(epochs or 0), shard
        if shard is not None and num_shards is None:
            num_shards, shard_id = _shard._get_or_create_num_shards_and_shard_id()
        self.num_shards, self.shard_id = num_shards or 1, shard_id or 0
        self.verify = verify_crc
        self._buf = ctypes.create_string_buffer(int(max_record_bytes))
        self._h = None

    # -- lifecycle -------------------------------------------------------------------
    def open(self):
        if self._h is not None:
            return self
        files, ns, sid = self.files, int(self.num_shards), int(self.shard_id)
        by_record = 0
        if self.shard == "file" and ns > 1:
            files = files[sid::ns]
            if not files:
                raise ValueError("shard %d of %d has no input files" % (sid, ns))
            ns, sid = 1, 0
        elif self.shard == "record" and ns > 1:
            by_record = 1
        arr = (ctypes.c_char_p * len(files))(*[f.encode() for f in files])
        h = ops.lib().px_loader_open(arr, len(files), self.kind, self.num_threads, self.capacity,
                                     self.min_after, int(self.shuffle), self.seed, self.epochs,
                                     ns, sid, by_record, int(self.verify))
        if h < 0:
            raise RuntimeError("px_loader_open failed")
        self._h = h
        return self

    def close(self):
        if self._h is not None:
            ops.lib().px_loader_close(self._h)
            self._h = None

    def __enter__(self):
        return self.open()

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _error(self):
        buf = ctypes.create_string_buffer(1024)
        ops.lib().px_loader_error(self._h, buf, 1024)
        return buf.value.decode()

    # -- reading ---------------------------------------------------------------------
    def next(self):
        """→ bytes, or None at the end of the data"""
        self.open()
        n = ctypes.c_size_t(0)
        L = ops.lib()
        rc = L.px_loader_next(self._h, self._buf, len(self._buf), ctypes.byref(n))
        if rc == 2:                                     # grow and retry
            self._buf = ctypes.create_string_buffer(int(n.value) * 2)
            rc = L.px_loader_next(self._h, self._buf, len(self._buf), ctypes.byref(n))
        if rc == 1:
            return None
        if rc != 0:
            raise RuntimeError("record loader failed: %s" % self._error())
        return self._buf.raw[:n.value]

    def next_batch(self, n):
        """exactly `n` records (fewer only at the end of the data) → list of bytes;
        one native call moves as many records as fit the transfer buffer"""
        self.open()
        L, out = ops.lib(), []
        offs = (ctypes.c_size_t * (n + 1))()
        while len(out) < n:
            want = n - len(out)
            got = L.px_loader_next_batch(self._h, want, self._buf, len(self._buf), offs)
            if got == -2:                               # first record alone does not fit
                self._buf = ctypes.create_string_buffer(int(offs[1]) * 2)
                continue
            if got < 0:
                raise RuntimeError("record loader failed: %s" % self._error())
            if got == 0:
                break
            raw = self._buf.raw[:offs[got]]
            out.extend(raw[offs[i]:offs[i + 1]] for i in range(got))
            if got < want and len(self._buf) < (64 << 20):
                self._buf = ctypes.create_string_buffer(len(self._buf) * 2)
        return out

    def __iter__(self):
        self.open()
        while True:
            recs = self.next_batch(256)
            if not recs:
                self.close()
                return
            for r in recs:
                yield r

    def stats(self):
        a, b, c = ctypes.c_long(), ctypes.c_long(), ctypes.c_long()
        ops.lib().px_loader_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {"records": a.value, "bytes": b.value, "crc_errors": c.value}


# ------------------------------------------------------------------ TFRecord
class TFRecordWriter(object):
    def __init__(self, path):
        self._f = open(path, "wb")

    def write(self, record):
        hdr = struct.pack("<Q", len(record))
        self._f.write(hdr)
        self._f.write(struct.pack("<I", masked_crc32c(hdr)))
        self._f.write(record)
        self._f.write(struct.pack("<I", masked_crc32c(record)))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def tfrecord_iterator(path, verify=True):
    """pure-python reader (tests / small files); `RecordLoader` is the fast path"""
    with open(path, "rb") as f:
        while True:
            hdr = f.read(12)
            if not hdr:
                return
            (n,), (lcrc,) = struct.unpack("<Q", hdr[:8]), struct.unpack("<I", hdr[8:])
            if verify and masked_crc32c(hdr[:8]) != lcrc:
                raise IOError("corrupted record length in %s" % path)
            data = f.read(n)
            (dcrc,) = struct.unpack("<I", f.read(4))
            if verify and masked_crc32c(data) != dcrc:
                raise IOError("corrupted record data in %s" % path)
            yield data


# ------------------------------------------------------- tf.train.Example wire
def _varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, i):
    shift = val = 0
    while True:
        b = buf[i]
        i += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, i
        shift += 7


def _ld(field, payload):                       # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_example(features):
    """dict name → bytes | str | list of (bytes|str) | ints | floats → serialized
    `tf.train.Example`"""
    entries = b""
    for name in sorted(features):
        v = features[name]
        if isinstance(v, (bytes, str, int, float, np.integer, np.floating)):
            v = [v]
        v = list(v)
        if v and isinstance(v[0], (bytes, str)):
            body = b"".join(_ld(1, x.encode() if isinstance(x, str) else x) for x in v)
            feat = _ld(1, body)                                     # BytesList
        elif v and isinstance(v[0], (float, np.floating)):
            feat = _ld(2, _ld(1, struct.pack("<%df" % len(v), *v)))  # FloatList (packed)
        else:
            feat = _ld(3, _ld(1, b"".join(_varint(int(x)) for x in v)))   # Int64List (packed)
        entries += _ld(1, _ld(1, name.encode()) + _ld(2, feat))     # map entry
    return _ld(1, entries)                                          # Example.features


def _fields(buf):
    i, n = 0, len(buf)
    while i < n:
        key, i = _read_varint(buf, i)
        field, wt = key >> 3, key & 7
        if wt == 2:
            ln, i = _read_varint(buf, i)
            yield field, wt, buf[i:i + ln]
            i += ln
        elif wt == 0:
            v, i = _read_varint(buf, i)
            yield field, wt, v
        elif wt == 5:
            yield field, wt, buf[i:i + 4]
            i += 4
        elif wt == 1:
            yield field, wt, buf[i:i + 8]
            i += 8
        else:
            raise ValueError("unsupported wire type %d" % wt)


def parse_example(serialized):
    """serialized `tf.train.Example` → dict name → list (bytes / float / int)"""
    out = {}
    for f, _, features in _fields(serialized):
        if f != 1:
            continue
        for f2, _, entry in _fields(features):
            if f2 != 1:
                continue
            name, feat = None, b""
            for f3, _, v in _fields(entry):
                if f3 == 1:
                    name = bytes(v).decode()
                elif f3 == 2:
                    feat = v
            vals = []
            for kind, _, lst in _fields(feat):
                for f5, wt, v in _fields(lst):
                    if kind == 1:
                        vals.append(bytes(v))
                    elif kind == 2:
                        if wt == 2:
                            vals.extend(struct.unpack("<%df" % (len(v) // 4), bytes(v)))
                        else:
                            vals.append(struct.unpack("<f", bytes(v))[0])
                    elif kind == 3:
                        if wt == 2:
                            j, b = 0, bytes(v)
                            while j < len(b):
                                x, j = _read_varint(b, j)
                                vals.append(x - (1 << 64) if x >= (1 << 63) else x)
                        else:
                            vals.append(v - (1 << 64) if v >= (1 << 63) else v)
            out[name] = vals
    return out


# ---------------------------------------------------------------- vocabulary
class NativeVocab(object):
    """word → id table with whitespace tokenisation done in native code"""

    def __init__(self, words, unk_id=0):
        blob = "\n".join(words).encode("utf-8")
        self._h = ops.lib().px_vocab_create(blob, len(blob), int(unk_id))
        self.size = len(words)
        self._out = (ctypes.c_longlong * 4096)()

    @classmethod
    def from_file(cls, path, unk_id=0):
        with open(path, encoding="utf-8") as f:
            words = [line.rstrip("\n") for line in f]
        while words and words[-1] == "":
            words.pop()
        return cls(words, unk_id)

    def encode(self, line):
        """str | bytes line → list of ids"""
        if isinstance(line, str):
            line = line.encode("utf-8")
        n = ops.lib().px_vocab_encode(self._h, line, len(line), self._out, len(self._out))
        if n > len(self._out):
            self._out = (ctypes.c_longlong * (2 * n))()
            n = ops.lib().px_vocab_encode(self._h, line, len(line), self._out, len(self._out))
        return self._out[:n]

    def __del__(self):
        try:
            ops.lib().px_vocab_free(self._h)
        except Exception:
            pass
