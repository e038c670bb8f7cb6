"""Minimal reader / writer for the safetensors container (no dependency): the persisted weight artefact of
SURVEY.md section 8f rank 2.  Layout: u64 little-endian header length, a JSON header
``{name: {"dtype", "shape", "data_offsets": [begin, end]}, "__metadata__": {str: str}}``, then the raw
little-endian tensor bytes.  Only the dtypes this package stores are supported."""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Tuple

import numpy as np

_DTYPES = {"F32": np.dtype("<f4"), "F16": np.dtype("<f2"), "U8": np.dtype("u1"), "U16": np.dtype("<u2"), "I32": np.dtype("<i4"), "I64": np.dtype("<i8")}
_NAMES = {v: k for k, v in _DTYPES.items()}


def save(path, tensors: Dict[str, np.ndarray], metadata: Dict[str, str] | None = None) -> None:
    header, blobs, off = {}, [], 0
    for name in sorted(tensors):
        a = np.ascontiguousarray(tensors[name])
        dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
        if np.dtype(dt) not in _NAMES:
            raise ValueError("tensor %r: dtype %s is not supported" % (name, a.dtype))
        raw = a.astype(dt, copy=False).tobytes()
        header[name] = {"dtype": _NAMES[np.dtype(dt)], "shape": list(a.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * (-len(hj) % 8)                      # keep the data 8-byte aligned
    tmp = os.fspath(path) + ".tmp"
    with open(tmp, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for raw in blobs:
            f.write(raw)
    os.replace(tmp, path)


def load(path) -> Tuple[Dict[str, np.ndarray], Dict[str, str]]:
    """Returns ({name: array}, metadata).  Arrays are read-only views of one buffer.  ValueError on a malformed file."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 8:
        raise ValueError("%s: too short for a safetensors header" % path)
    (hlen,) = struct.unpack("<Q", buf[:8])
    if hlen > len(buf) - 8 or hlen > (100 << 20):
        raise ValueError("%s: header length %d exceeds the file" % (path, hlen))
    try:
        header = json.loads(buf[8:8 + hlen].decode("utf-8"))
    except (UnicodeDecodeError, json.JSONDecodeError) as e:
        raise ValueError("%s: header is not JSON (%s)" % (path, e)) from None
    if not isinstance(header, dict):
        raise ValueError("%s: header is not a JSON object" % path)
    meta = header.pop("__metadata__", {}) or {}
    data = memoryview(buf)[8 + hlen:]
    out = {}
    for name, d in header.items():
        try:
            dt, shape, (b, e) = _DTYPES[d["dtype"]], tuple(int(s) for s in d["shape"]), d["data_offsets"]
        except (KeyError, TypeError, ValueError):
            raise ValueError("%s: bad entry for tensor %r" % (path, name)) from None
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if not (0 <= b <= e <= len(data)) or e - b != n * dt.itemsize:
            raise ValueError("%s: tensor %r: offsets [%s,%s) do not match shape %s" % (path, name, b, e, shape))
        out[name] = np.frombuffer(data[b:e], dtype=dt).reshape(shape)
    return out, {str(k): str(v) for k, v in meta.items()}
