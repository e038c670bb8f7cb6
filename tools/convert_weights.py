"""Convert the reference's Keras ``WHENet.h5`` into a plain ``.npz`` (or, by extension, ``.safetensors``).

``/root/reference`` does not exist on the GPU box, so the weights the parity
tests need must travel inside the repo.  The ``.npz`` holds the 315 float32
tensors bit-for-bit under their original names plus ``__layer_names__`` (the
file's graph-order attribute).  No folding, no re-layout: the CUDA library and
the oracle both start from the same raw tensors.

    python tools/convert_weights.py /root/reference/WHENet.h5 headposeestimation-whenet_b200/data/whenet_weights.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                "headposeestimation-whenet_b200"))
import h5lite  # noqa: E402


def main(src, dst):
    if dst.endswith(".safetensors"):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
        from whenet_b200 import weights as wmod
        names, w = wmod.load_snapshot(src)                 # .h5, .npz or .safetensors in
        wmod.save_safetensors(dst, names, w)
        print("wrote %s: %d tensors" % (dst, len(w)))
        return
    layer_names, weights, meta = h5lite.read_keras_weights(src)
    out = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
    out["__layer_names__"] = np.array(layer_names)
    out["__backend__"] = np.array(meta.get("backend", ""))
    out["__keras_version__"] = np.array(meta.get("keras_version", ""))
    np.savez(dst, **out)
    n = sum(v.size for k, v in out.items() if not k.startswith("__"))
    print("wrote %s: %d tensors, %d values" % (dst, len(weights), n))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
