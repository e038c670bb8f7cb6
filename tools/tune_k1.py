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
m.set_option("streams", 1)      # per-kernel event times need the kernels of one forward back to back on one stream
def run():
    m.get_angle(x)
    m.enable_profile(True)
    for _ in range(int(os.environ.get("REPS", "3"))):
        m.get_angle(x)
    st = m.read_profile(); m.enable_profile(False)
    return {s["name"]: s["ms"] / s["launches"] for s in st}
base = run()
print("default plans:", {k: round(v, 4) for k, v in base.items() if k.endswith(".k1")}, flush=True)
best = {}
allres = {}
tiles = [(14, 14, 7), (7, 14, 7), (14, 7, 7), (7, 7, 7), (7, 7, 4), (8, 8, 4), (8, 7, 4)]
modes = [(256, 1), (512, 1), (512, 2), (256, 2)]
only = [int(v) for v in os.environ.get("BLOCKS", "").split(",") if v]
for (th, tw, r), cc, (nt, nb) in itertools.product(tiles, (128, 112, 96, 80, 64, 48, 32), modes):
    ok = []
    for b in arch.blocks():
        if b.has_expand and (not only or b.idx in only) and m.set_k1_plan(b.idx, th, tw, r, cc, nt, nb):
            ok.append(b.idx)
    if not ok:
        continue
    try:
        t = run()
    except Exception as e:
        print("plan", th, tw, r, cc, nt, nb, "failed:", e, flush=True)
        continue
    for i in ok:
        v = t.get("b%02d.k1" % i)
        if v is not None:
            allres.setdefault(i, []).append((v, th, tw, r, cc, nt, nb))
            if i not in best or v < best[i][0]:
                best[i] = (v, th, tw, r, cc, nt, nb)
    print("plan %dx%d r%d cc%d nt%d nb%d:" % (th, tw, r, cc, nt, nb), {i: round(t.get("b%02d.k1" % i, -1), 4) for i in ok}, flush=True)
print("BEST per block (ms per launch at N=%d):" % N)
for i in sorted(best):
    print("  block %2d: %.4f ms  th=%d tw=%d r=%d cc=%d nt=%d nb=%d   (default %.4f)" % ((i,) + best[i] + (base.get("b%02d.k1" % i, -1),)))
    for v in sorted(allres[i])[1:4]:
        print("            %.4f ms  th=%d tw=%d r=%d cc=%d nt=%d nb=%d" % v)
print("sum default %.4f ms, sum best %.4f ms" % (sum(base.get("b%02d.k1" % i, 0) for i in best), sum(best[i][0] for i in best)))
