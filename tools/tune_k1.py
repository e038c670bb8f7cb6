"""Time every feasible K1 tile plan of every block on the GPU and print the fastest (tuning aid)."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
from whenet_b200 import arch
N = int(os.environ.get("N", "256"))
x = np.random.default_rng(0).integers(0, 256, (N, 224, 224, 3), dtype=np.uint8)
m = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision="bf16", max_batch=N)
m.set_option("chunk", N)
def run():
    m.get_angle(x)
    m.enable_profile(True); m.get_angle(x); m.get_angle(x); st = m.read_profile(); m.enable_profile(False)
    return {s["name"]: s["ms"] / s["launches"] for s in st}
base = run()
print("default plans:", {k: round(v, 4) for k, v in base.items() if k.endswith(".k1")}, flush=True)
best = {}
tiles = [(14, 14, 7), (7, 14, 7), (7, 7, 7), (7, 7, 4), (8, 8, 4), (14, 7, 7)]
for (th, tw, r), cc in itertools.product(tiles, (128, 112, 96, 80, 64, 48, 32)):
    ok = []
    for b in arch.blocks():
        if b.has_expand and m.set_k1_plan(b.idx, th, tw, r, cc):
            ok.append(b.idx)
    if not ok:
        continue
    try:
        t = run()
    except Exception as e:
        print("plan", th, tw, r, cc, "failed:", e, flush=True)
        continue
    for i in ok:
        v = t.get("b%02d.k1" % i)
        if v is not None and (i not in best or v < best[i][0]):
            best[i] = (v, th, tw, r, cc)
    print("plan %dx%d r%d cc%d:" % (th, tw, r, cc), {i: round(t.get("b%02d.k1" % i, -1), 4) for i in ok}, flush=True)
print("BEST per block (ms per launch at N=%d):" % N)
for i in sorted(best):
    print("  block %2d: %.4f ms  th=%d tw=%d r=%d cc=%d   (default %.4f)" % ((i,) + best[i] + (base.get("b%02d.k1" % i, -1),)))
