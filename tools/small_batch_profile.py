"""Per-kernel device times at a small batch (default N=1): where the single-crop latency goes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
n = int(os.environ.get("N", "1"))
m = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision="bf16", max_batch=max(n, 8))
x = np.random.default_rng(0).integers(0, 256, (n, 224, 224, 3), dtype=np.uint8)
for _ in range(5):
    m.get_angle(x)
m.enable_profile(True)
for _ in range(20):
    m.get_angle(x)
st = m.read_profile()
tot = sum(s["ms"] for s in st) / 20
print("N=%d sum of kernel times %.3f ms" % (n, tot))
for s in st:
    print("  %-16s %.4f ms" % (s["name"], s["ms"] / s["launches"]))
