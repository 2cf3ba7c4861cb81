"""Native input pipeline (`ops/csrc/runtime/dataloader.cpp`): TFRecord framing and
crc32c, tf.Example wire format, multi-threaded shuffle pool, sharding by file and
by record, text lines, native vocabulary lookup."""
import collections
import os

import pytest

import parallax_b200 as parallax
from parallax_b200.utils import dataloader as dl


def _write(d, nfiles=3, per=10):
    for i in range(nfiles):
        with dl.TFRecordWriter(os.path.join(d, "part-%05d" % i)) as w:
            for j in range(per):
                w.write(dl.encode_example({"image/class/label": i * per + j,
                                           "image/encoded": bytes([j]) * (37 * j),
                                           "image/object/bbox/xmin": [0.1, 0.25],
                                           "image/class/text": "cls%d" % j}))
    return os.path.join(d, "part-*")


def _labels(recs):
    return [dl.parse_example(r)["image/class/label"][0] for r in recs]


def test_crc32c_known_values():
    assert dl.crc32c(b"123456789") == 0xE3069283          # the standard check value
    assert dl.crc32c(b"") == 0
    assert dl.crc32c(bytes(32)) == 0x8A9136AA             # RFC 3720 B.4: 32 zero bytes
    assert dl.masked_crc32c(b"abc") == (((dl.crc32c(b"abc") >> 15) |
                                         (dl.crc32c(b"abc") << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_example_roundtrip():
    ex = {"a": [1, -2, 1 << 40], "b": [0.5, -1.25], "c": [b"\x00\xff", b"xyz"], "d": "text", "e": 7}
    got = dl.parse_example(dl.encode_example(ex))
    assert got == {"a": [1, -2, 1 << 40], "b": [0.5, -1.25], "c": [b"\x00\xff", b"xyz"],
                   "d": [b"text"], "e": [7]}


def test_tfrecord_read_in_order_and_python_reader_agrees(tmp_path):
    pat = _write(str(tmp_path))
    loader = dl.RecordLoader(pat, dl.TFRECORD, shard=None)
    recs = list(loader)
    assert _labels(recs) == list(range(30))
    py = [r for f in sorted(os.listdir(tmp_path)) for r in dl.tfrecord_iterator(str(tmp_path / f))]
    assert py == recs
    ex = dl.parse_example(recs[13])
    assert ex["image/encoded"] == [bytes([3]) * 111] and ex["image/class/text"] == [b"cls3"]
    assert [round(v, 3) for v in ex["image/object/bbox/xmin"]] == [0.1, 0.25]
    with dl.RecordLoader(pat, dl.TFRECORD, shard=None, max_record_bytes=16) as small:
        assert small.next() == recs[0] and len(small.next_batch(40)) == 29     # buffer grows
        assert small.next() is None
        st = small.stats()
        assert st["records"] == 30 and st["crc_errors"] == 0 and st["bytes"] == sum(map(len, recs))


def test_sharding_by_record_and_by_file(tmp_path):
    pat = _write(str(tmp_path))
    seen = []
    for k in range(4):
        got = _labels(dl.RecordLoader(pat, dl.TFRECORD, shard="record", num_shards=4, shard_id=k))
        assert got == list(range(k, 30, 4))
        seen += got
    assert sorted(seen) == list(range(30))
    by_file = _labels(dl.RecordLoader(pat, dl.TFRECORD, shard="file", num_shards=3, shard_id=1))
    assert by_file == list(range(10, 20))
    with pytest.raises(ValueError):
        dl.RecordLoader(pat, dl.TFRECORD, shard="file", num_shards=5, shard_id=4).open()
    # default: the values parallax assigns to this worker, resolved when reading starts
    loader = dl.RecordLoader(pat, dl.TFRECORD)
    parallax.shard.update_shard_values_for_worker(2, 1, 1)
    assert _labels(loader) == list(range(1, 30, 2))


def test_shuffle_pool_and_epochs(tmp_path):
    pat = _write(str(tmp_path), nfiles=4, per=25)
    a = _labels(dl.RecordLoader(pat, dl.TFRECORD, shard=None, shuffle=True, capacity=32, seed=5,
                                epochs=3, num_threads=3))
    assert len(a) == 300 and collections.Counter(a) == collections.Counter(list(range(100)) * 3)
    assert a[:100] != sorted(a[:100])
    b = _labels(dl.RecordLoader(pat, dl.TFRECORD, shard=None, shuffle=True, capacity=32, seed=6,
                                epochs=1, num_threads=1))
    c = _labels(dl.RecordLoader(pat, dl.TFRECORD, shard=None, shuffle=True, capacity=32, seed=6,
                                epochs=1, num_threads=1))
    assert b == c and sorted(b) == list(range(100))      # one reader thread ⇒ reproducible
    # infinite epochs: keeps producing, closing stops the reader threads
    inf = dl.RecordLoader(pat, dl.TFRECORD, shard=None, shuffle=True, capacity=8, epochs=0)
    assert len([inf.next() for _ in range(250)]) == 250
    inf.close()


def test_corruption_is_detected(tmp_path):
    pat = _write(str(tmp_path), nfiles=1)
    fn = str(tmp_path / "part-00000")
    raw = bytearray(open(fn, "rb").read())
    raw[len(raw) // 2] ^= 0xFF
    open(fn, "wb").write(bytes(raw))
    with pytest.raises(RuntimeError, match="corrupted"):
        list(dl.RecordLoader(pat, dl.TFRECORD, shard=None))
    with pytest.raises(IOError):
        list(dl.tfrecord_iterator(fn))
    assert len(list(dl.RecordLoader(pat, dl.TFRECORD, shard=None, verify_crc=False))) == 10
    with pytest.raises(ValueError):
        dl.RecordLoader(str(tmp_path / "nothing-*"), dl.TFRECORD)


def test_text_lines_and_native_vocab(tmp_path):
    f = tmp_path / "a.txt"
    f.write_bytes(b"hello world\r\nfoo  bar baz\n\n" + b"x " * 40000 + b"\nlast line no newline")
    lines = list(dl.RecordLoader(str(f), dl.TEXT, shard=None))
    assert lines[:3] == [b"hello world", b"foo  bar baz", b""] and len(lines[3]) == 80000
    assert lines[4] == b"last line no newline"
    vf = tmp_path / "vocab.txt"
    vf.write_text("<unk>\nhello\nworld\nfoo\nhello\n")
    v = dl.NativeVocab.from_file(str(vf))
    assert v.size == 5 and v.encode("hello  foo\tzzz world\n") == [1, 3, 0, 2]
    assert v.encode(b"") == [] and len(v.encode(lines[3])) == 40000
    ids = [v.encode(l) for l in dl.RecordLoader(str(f), dl.TEXT, shard="record", num_shards=2,
                                                shard_id=0)]
    assert ids[0] == [1, 2] and ids[1] == []
