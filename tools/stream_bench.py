"""BASELINE configs[4]: 1080p stream, WHENet stage only (the YOLO head detector's weights are not in the reference
tree, so boxes are synthetic: 1..32 heads per frame, side 48..400 px).  Per frame: H2D of the frame,
GPU crop front-end (demo_video.py:13-23 for all heads at once), WHENet forward, D2H of the angles."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
frames = int(os.environ.get("FRAMES", "200"))
prec = os.environ.get("PREC", "bf16")
rng = np.random.default_rng(2)
m = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision=prec, max_batch=32)
frame = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
work = []
for _ in range(frames):
    n = int(rng.integers(1, 33))
    side = rng.integers(48, 401, n)
    y0 = rng.integers(0, 1080 - side); x0 = rng.integers(0, 1920 - side)
    work.append(np.stack([y0, x0, y0 + side, x0 + side], axis=1).astype(np.float32))
for b in work[:10]:
    m.get_angle_from_frame(frame, b)
t0 = time.perf_counter()
heads = 0
for b in work:
    m.get_angle_from_frame(frame, b)
    heads += len(b)
dt = time.perf_counter() - t0
print(json.dumps({"config": "1080p stream, synthetic boxes (1..32 heads/frame), WHENet stage only, %s" % prec,
                  "frames": frames, "heads": heads, "frames_per_s": frames / dt, "crops_per_s": heads / dt,
                  "ms_per_frame": dt / frames * 1e3}))
