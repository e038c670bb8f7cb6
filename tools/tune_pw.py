"""Time the 1x1-conv (project / head) kernels under different ring-depth / N-split settings (tuning aid)."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
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
names = None
rows = []
for min_ctas, kb in itertools.product((296, 148, 100, 0), (54, 75, 110, 180)):
    m.set_option("pw_min_ctas", min_ctas)
    m.set_option("pw_smem_kb", kb)
    t = run()
    if names is None:
        names = [k for k in t if k.endswith(".project") or k == "head.conv"]
        print("%-18s" % "min_ctas/smem_kb", " ".join("%7s" % n.replace(".project", ".p").replace("head.conv", "head") for n in names), "    sum", flush=True)
    rows.append(((min_ctas, kb), [t[n] for n in names]))
    print("%-18s" % ("%d/%d" % (min_ctas, kb)), " ".join("%7.4f" % t[n] for n in names), "%7.4f" % sum(t[n] for n in names), flush=True)
best = [min(r[1][i] for r in rows) for i in range(len(names))]
print("%-18s" % "best per layer", " ".join("%7.4f" % b for b in best), "%7.4f" % sum(best))
for i, n in enumerate(names):
    cfg = min(rows, key=lambda r: r[1][i])[0]
    print("  %-14s best %.4f ms at min_ctas=%d smem_kb=%d" % (n, best[i], cfg[0], cfg[1]))
