"""Generate the committed parity fixtures under tests/golden/.

Run HERE (the container that has /root/reference); the GPU box only reads the
outputs.  Two kinds of fixture:

* ``sample_crops.npy``  - the reference's own two inputs: ``Sample/*.jpeg`` cropped
  with ``Sample/bbox.txt`` through the exact recipe of reference ``demo.py:7-12``
  (imread -> BGR2RGB -> slice [y0:y1, x0:x1] -> cv2.resize 224x224 default
  INTER_LINEAR -> uint8).
* ``golden.json``       - float64 oracle angles + logits for those crops and for a
  seeded synthetic batch.  NOT Keras outputs (Keras cannot run here): see the
  "parity unpinned" note in oracle/whenet_oracle.py.

    python tools/make_golden.py
"""
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from whenet_oracle import load_oracle  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def demo_crop(path, bbox):
    img = cv2.imread(path)                                  # demo.py:7
    img_rgb = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)          # demo.py:8
    x_min, y_min, x_max, y_max = bbox                       # demo.py:9
    img_rgb = img_rgb[y_min:y_max, x_min:x_max]             # demo.py:10
    return cv2.resize(img_rgb, (224, 224))                  # demo.py:11


def main():
    crops, names, boxes = [], [], []
    with open(os.path.join(REF, "Sample", "bbox.txt")) as f:
        for line in f.read().splitlines():
            if not line.strip():
                continue
            fn, bb = line.split(",")                         # demo.py:28-30
            bb = [int(b) for b in bb.split(" ")]
            crops.append(demo_crop(os.path.join(REF, "Sample", fn), bb))
            names.append(fn); boxes.append(bb)
    crops = np.stack(crops).astype(np.uint8)
    np.save(os.path.join(GOLD, "sample_crops.npy"), crops)

    o64 = load_oracle(os.path.join(REF, "WHENet.h5"), np.float64)
    (yaw, pitch, roll), logits = o64.get_angle(crops, return_logits=True)
    # float64 angles (before the float32 cast of the public surface)
    from whenet_oracle import decode
    y64, p64, r64 = decode(*logits)
    gold = {
        "note": "float64 CPU oracle (tools/make_golden.py), NOT Keras; parity unpinned by the reference",
        "samples": [
            {"file": names[i], "bbox": boxes[i],
             "yaw": float(y64[i]), "pitch": float(p64[i]), "roll": float(r64[i]),
             "argmax": [int(np.argmax(l[i])) for l in logits]}
            for i in range(len(names))],
    }
    # seeded synthetic batch: natural-image statistics via jittered re-crops of the samples
    rng = np.random.default_rng(1)
    syn = []
    for i in range(6):
        src = cv2.cvtColor(cv2.imread(os.path.join(REF, "Sample", names[i % 2])), cv2.COLOR_BGR2RGB)
        x0, y0, x1, y1 = boxes[i % 2]
        w, h = x1 - x0, y1 - y0
        dx, dy = rng.integers(-w // 10, w // 10 + 1), rng.integers(0, h // 10 + 1)
        s = 1.0 + rng.uniform(-0.1, 0.1)
        nx0 = int(np.clip(x0 + dx, 0, src.shape[1] - 8)); ny0 = int(np.clip(y0 + dy, 0, src.shape[0] - 8))
        nx1 = int(np.clip(nx0 + w * s, nx0 + 8, src.shape[1])); ny1 = int(np.clip(ny0 + h * s, ny0 + 8, src.shape[0]))
        c = cv2.resize(src[ny0:ny1, nx0:nx1], (224, 224))
        if rng.integers(0, 2):
            c = c[:, ::-1]
        syn.append(np.ascontiguousarray(c))
    syn = np.stack(syn).astype(np.uint8)
    np.save(os.path.join(GOLD, "jitter_crops.npy"), syn)
    _a, lg = o64.get_angle(syn, return_logits=True)
    ys, ps, rs = decode(*lg)
    gold["jitter"] = {"yaw": [float(v) for v in ys], "pitch": [float(v) for v in ps],
                      "roll": [float(v) for v in rs]}
    np.save(os.path.join(GOLD, "sample_logits_f64.npy"), np.concatenate(logits, axis=1))
    with open(os.path.join(GOLD, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print(json.dumps(gold, indent=1))


if __name__ == "__main__":
    main()
