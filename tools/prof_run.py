"""Tiny driver for ncu: a few bf16 forwards at N crops (default 128), nothing else."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
n = int(os.environ.get("N", "128"))
reps = int(os.environ.get("REPS", "2"))
m = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision=os.environ.get("PREC", "bf16"), max_batch=n)
m.set_option("chunk", n)
for k, v in (("tensor_cores", "TC"), ("dw_variant", "DWV"), ("fused", "FUSED"), ("streams", "STREAMS")):
    if v in os.environ:
        m.set_option(k, int(os.environ[v]))
for kv in os.environ.get("OPTS", "").split(","):
    if "=" in kv:
        k, v = kv.split("=")
        m.set_option(k, int(v))
x = np.random.default_rng(0).integers(0, 256, (n, 224, 224, 3), dtype=np.uint8)
for _ in range(reps):
    m.get_angle(x)
print("done", m.launch_count())
