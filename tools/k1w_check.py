"""K1W (k1_variant=4) against K1: taps of every block, then per-kernel CUDA-event times at N crops (streams=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
GOLD = os.path.join(ROOT, "tests", "golden")
crops = np.concatenate([np.load(os.path.join(GOLD, "sample_crops.npy")), np.load(os.path.join(GOLD, "jitter_crops.npy"))])[:int(os.environ.get("NT", "3"))]
N = int(os.environ.get("N", "256"))
m = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision=os.environ.get("PREC", "bf16"), max_batch=N)
m.set_option("streams", 1)
def taps(variant):
    m.set_option("k1_variant", variant)
    m.enable_taps(True)
    ang = np.stack(m.get_angle(crops), axis=1)
    out = {"angles": ang}
    for i in range(2, 17):
        for kind in ("dw", "gate", "block"):
            out["%s%d" % (kind, i)] = m.tap("%s%d" % (kind, i)).astype(np.float64)
    m.enable_taps(False)
    return out
ref = taps(1)
try:
    got = taps(4)
    for k in ref:
        d = np.abs(got[k] - ref[k])
        rel = np.sqrt((d ** 2).mean()) / (np.sqrt((ref[k] ** 2).mean()) + 1e-30)
        print("%-8s max|d| %.3e  rms-rel %.3e (ref max %.3e, nan %d)" % (k, d.max(), rel, np.abs(ref[k]).max(), int(np.isnan(got[k]).sum())), flush=True)
except Exception as e:
    print("K1W taps FAILED:", e, flush=True)
    sys.exit(0)
x = np.random.default_rng(0).integers(0, 256, (N, 224, 224, 3), dtype=np.uint8)
def prof(variant):
    m.set_option("k1_variant", variant)
    m.set_option("chunk", N)
    m.get_angle(x)
    m.enable_profile(True)
    for _ in range(3):
        m.get_angle(x)
    st = m.read_profile(); m.enable_profile(False)
    return {s["name"]: s["ms"] / s["launches"] for s in st}
try:
    a = prof(1)
    b = prof(4)
    ta = tb = 0.0
    for i in range(2, 17):
        k = "b%02d.k1" % i
        se = "b%02d.se" % i
        print("%s  K1 %.4f ms (+se %.4f)   K1W %.4f ms (+se %.4f)   x%.2f" % (k, a[k], a.get(se, 0.0), b[k], b.get(se, 0.0), (a[k] + a.get(se, 0.0)) / (b[k] + b.get(se, 0.0))), flush=True)
        ta += a[k]; tb += b[k]
    print("K1 family: K1 %.3f ms  K1W %.3f ms;   total ms/forward: K1 %.3f  K1W %.3f" % (ta, tb, sum(a.values()), sum(b.values())), flush=True)
    full = np.stack(m.get_angle(x[:64]), axis=1)
    one = np.stack(m.get_angle(x[5:6]), axis=1)
    print("batch invariance (64 vs 1):", bool(np.array_equal(full[5], one[0])))
except Exception as e:
    print("K1W timing FAILED:", e, flush=True)
