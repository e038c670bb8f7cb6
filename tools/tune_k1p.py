"""Time K1P (k1_variant=3) tile plans per block against the K1 default (tuning aid, round-2 starting point).

Per-kernel CUDA-event times at N crops on one stream; every plan is checked against K1's angles first (8 crops)."""
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
m.set_option("streams", 1)
def run():
    m.get_angle(x)
    m.enable_profile(True)
    for _ in range(int(os.environ.get("REPS", "3"))):
        m.get_angle(x)
    st = m.read_profile(); m.enable_profile(False)
    return {s["name"]: s["ms"] / s["launches"] for s in st}
base = run()
ref = np.stack(m.get_angle(x[:8]), axis=1)
print("K1:", {k: round(v, 4) for k, v in base.items() if k.endswith(".k1")}, flush=True)
m.set_option("k1_variant", 3)
best = {}
tiles = [(14, 14, 7), (14, 14, 4), (7, 14, 7), (14, 7, 7), (7, 7, 7), (7, 7, 4), (8, 8, 4), (8, 7, 4)]
for (th, tw, r), cc, ew in itertools.product(tiles, (96, 80, 64, 48, 32), (8, 4)):
    ok = [b.idx for b in arch.blocks() if b.has_expand and m.set_k1p_plan(b.idx, th, tw, r, cc, ew)]
    if not ok:
        continue
    try:
        got = np.stack(m.get_angle(x[:8]), axis=1)
        if np.abs(got - ref).max() > 0.25:
            print("plan", th, tw, r, cc, ew, "WRONG by", float(np.abs(got - ref).max()), flush=True)
            continue
        t = run()
    except Exception as e:
        print("plan", th, tw, r, cc, ew, "failed:", e, flush=True)
        continue
    for i in ok:
        v = t.get("b%02d.k1" % i)
        if v is not None and (i not in best or v < best[i][0]):
            best[i] = (v, th, tw, r, cc, ew)
    print("plan %dx%d r%d cc%d epi%d:" % (th, tw, r, cc, ew), {i: round(t.get("b%02d.k1" % i, -1), 4) for i in ok}, flush=True)
print("BEST K1P per block (ms per launch at N=%d) vs K1:" % N)
for i in sorted(best):
    print("  block %2d: %.4f ms  th=%d tw=%d r=%d cc=%d epi_warps=%d   (K1 %.4f)" % ((i,) + best[i] + (base.get("b%02d.k1" % i, -1),)))
