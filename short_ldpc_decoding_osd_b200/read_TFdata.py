"""TensorFlow-free reader/writer of the reference's on-disk interchange format.

Every stage of the reference talks to the next through TFRecord files of
``tf.train.Example{feature: FloatList[128], label: Int64List[128], shape: Int64List[1]}``
(writer ``LDPC_128/Ldpc_128_testing/data_generating.py:8-26`` / ``Testing_data_gen_128/Main_test.py:66-83``,
reader ``*/read_TFdata.py:10-29``), including the 13-rows-per-failure retest files of
``ms_test.save_decoded_data`` (``ms_test.py:251-272``).  This module reads and writes that format with
NumPy only, so files produced by the reference feed this decoder and vice versa:

* ``data_handler(code_length, file_name, batch_size)`` -> dataset with ``as_numpy_iterator()`` yielding
  ``(float32[b,128], int64[b,128], int32[b])`` batches (``drop_remainder=False``), like the reference's.
* ``make_tfrecord((feats, labels), out_filename)`` -- same signature as the reference's writer.

TFRecord framing: uint64 length, masked CRC32C(length), payload, masked CRC32C(payload) (little-endian).
"""
from __future__ import annotations

import struct
from typing import Iterator, Tuple

import numpy as np

# ---- CRC32C (Castagnoli), table driven, vectorised over a record with NumPy -------------------------
_POLY = 0x82F63B78
_TABLE = np.zeros(256, dtype=np.uint32)
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _TABLE[_i] = _c
_TABLE_LIST = [int(v) for v in _TABLE]


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    t = _TABLE_LIST
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- protobuf wire helpers -------------------------------------------------------------------------------
def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _ld(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def serialize_example(feature: np.ndarray, label: np.ndarray) -> bytes:
    """Example with keys feature (float_list), label (int64_list), shape (int64_list = [len(feature)])."""
    f = np.asarray(feature, dtype="<f4").reshape(-1)
    lab = np.asarray(label, dtype=np.int64).reshape(-1)
    float_list = _ld(1, f.tobytes())                                   # FloatList.value, packed
    feat_f = _ld(2, float_list)                                        # Feature.float_list
    int_list = _ld(1, b"".join(_varint(int(v)) for v in lab))          # Int64List.value, packed
    feat_l = _ld(3, int_list)                                          # Feature.int64_list
    feat_s = _ld(3, _ld(1, _varint(f.shape[0])))
    entries = b"".join(_ld(1, _ld(1, k) + _ld(2, v)) for k, v in ((b"feature", feat_f), (b"label", feat_l), (b"shape", feat_s)))
    return _ld(1, entries)                                             # Example.features


def _parse_list(buf: bytes, kind: int) -> np.ndarray:
    """FloatList (kind 2) / Int64List (kind 3): packed or repeated scalar encodings."""
    pos, vals = 0, []
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        wt = tag & 7
        if wt == 2:
            n, pos = _read_varint(buf, pos)
            chunk = buf[pos:pos + n]
            pos += n
            if kind == 2:
                vals.append(np.frombuffer(chunk, dtype="<f4"))
            else:
                p, out = 0, []
                while p < len(chunk):
                    v, p = _read_varint(chunk, p)
                    out.append(v - (1 << 64) if v >> 63 else v)
                vals.append(np.array(out, dtype=np.int64))
        elif wt == 5:
            vals.append(np.frombuffer(buf[pos:pos + 4], dtype="<f4"))
            pos += 4
        elif wt == 0:
            v, pos = _read_varint(buf, pos)
            vals.append(np.array([v - (1 << 64) if v >> 63 else v], dtype=np.int64))
        else:
            raise ValueError(f"unexpected wire type {wt}")
    if not vals:
        return np.zeros(0, dtype=np.float32 if kind == 2 else np.int64)
    return np.concatenate(vals)


def parse_example(buf: bytes) -> dict:
    out = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        n, pos = _read_varint(buf, pos)
        features, pos = buf[pos:pos + n], pos + n
        if tag >> 3 != 1:
            continue
        p = 0
        while p < len(features):
            _, p = _read_varint(features, p)
            m, p = _read_varint(features, p)
            entry, p = features[p:p + m], p + m
            q, key, val = 0, None, None
            while q < len(entry):
                t, q = _read_varint(entry, q)
                ln, q = _read_varint(entry, q)
                chunk, q = entry[q:q + ln], q + ln
                if t >> 3 == 1:
                    key = chunk.decode()
                else:
                    val = chunk
            if key is None or val is None:
                continue
            t, r = _read_varint(val, 0)
            ln, r = _read_varint(val, r)
            kind = t >> 3
            out[key] = val[r:r + ln] if kind == 1 else _parse_list(val[r:r + ln], kind)
    return out


def tfrecord_iterator(path: str, verify: bool = True) -> Iterator[bytes]:
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError(f"{path}: truncated record header")
            (length,), (lcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify and masked_crc(head[:8]) != lcrc:
                raise IOError(f"{path}: corrupted record length")
            data = f.read(length)
            tail = f.read(4)
            if len(data) < length or len(tail) < 4:
                raise IOError(f"{path}: truncated record")
            if verify and masked_crc(data) != struct.unpack("<I", tail)[0]:
                raise IOError(f"{path}: corrupted record payload")
            yield data


class TFRecordWriter:
    def __init__(self, path: str):
        self._f = open(path, "wb")

    def write(self, record: bytes) -> None:
        head = struct.pack("<Q", len(record))
        self._f.write(head + struct.pack("<I", masked_crc(head)) + record + struct.pack("<I", masked_crc(record)))

    def close(self) -> None:
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def make_tfrecord(data, out_filename: str) -> None:
    """data = (feats [n,128], labels [n,128]) -> TFRecord file (data_generating.py:16-26)."""
    feats, labels = data
    with TFRecordWriter(out_filename) as w:
        for inx in range(len(labels)):
            w.write(serialize_example(feats[inx], labels[inx]))


class _Dataset:
    def __init__(self, fname: str, code_length: int, batch_size: int, verify: bool):
        self.fname, self.code_length, self.batch_size, self.verify = fname, code_length, batch_size, verify

    def as_numpy_iterator(self):
        feats, labs, shapes = [], [], []
        for rec in tfrecord_iterator(self.fname, self.verify):
            ex = parse_example(rec)
            f, lab = ex["feature"], ex["label"]
            if f.shape[0] != self.code_length or lab.shape[0] != self.code_length:
                raise ValueError(f"record with {f.shape[0]} features / {lab.shape[0]} labels, expected {self.code_length}")
            feats.append(f.astype(np.float32))
            labs.append(lab)
            shapes.append(np.int32(ex["shape"][0]) if "shape" in ex and len(ex["shape"]) else np.int32(self.code_length))
            if len(feats) == self.batch_size:
                yield np.stack(feats), np.stack(labs), np.array(shapes, dtype=np.int32)
                feats, labs, shapes = [], [], []
        if feats:  # drop_remainder=False (read_TFdata.py:28)
            yield np.stack(feats), np.stack(labs), np.array(shapes, dtype=np.int32)

    def __iter__(self):
        return self.as_numpy_iterator()


def get_dataset(fname, code_length, verify: bool = True):
    return _Dataset(fname, code_length, 1, verify)


def data_handler(code_length, file_name, batch_size=1, verify: bool = True):
    return _Dataset(file_name, code_length, batch_size, verify)
