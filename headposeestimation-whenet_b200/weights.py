"""Weight sources for ``WHENet(snapshot)``: the reference's Keras ``.h5``
(reference whenet.py:15-16), the converted ``.npz`` that travels with the repo, a
``.safetensors`` artefact (same raw tensors, no HDF5 walk: SURVEY.md 8f rank 2), or a
seeded random initialisation (``snapshot=None``, reference whenet.py:15)."""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

from . import arch, h5lite, stlite

# the reference's WHENet.h5 tensors (bit for bit, converted by tools/convert_weights.py): ships inside the package
DEFAULT_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "whenet_weights.npz")


def load_snapshot(snapshot) -> Tuple[List[str], Dict[str, np.ndarray]]:
    """Returns (layer_names, {name: float32 array}).  Raises OSError / ValueError like Keras."""
    if snapshot is None:
        return [], arch.random_weights(0)
    snapshot = os.fspath(snapshot)
    if not os.path.exists(snapshot):
        raise OSError("Unable to open file (name = %r, no such file)" % snapshot)
    if snapshot.endswith(".npz"):
        z = np.load(snapshot, allow_pickle=False)
        names = [str(s) for s in z["__layer_names__"]] if "__layer_names__" in z.files else []
        w = {k: np.ascontiguousarray(z[k], dtype=np.float32) for k in z.files if not k.startswith("__")}
    elif snapshot.endswith(".safetensors"):
        z, meta = stlite.load(snapshot)
        names = [s for s in meta.get("layer_names", "").split(",") if s]
        w = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in z.items()}
    else:
        names, w, _meta = h5lite.read_keras_weights(snapshot)
        w = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in w.items()}
    arch.assign_weights(w)   # shape-check every tensor, as load_weights does
    return names, w


def save_safetensors(path, layer_names: List[str], w: Dict[str, np.ndarray], extra: Dict[str, str] | None = None) -> None:
    """Persist the raw float32 tensors (bit for bit, original names) with the graph-order layer list as metadata."""
    meta = {"layer_names": ",".join(layer_names), "format": "whenet-keras-raw-f32"}
    meta.update(extra or {})
    stlite.save(path, {k: np.asarray(v, dtype=np.float32) for k, v in w.items()}, meta)
