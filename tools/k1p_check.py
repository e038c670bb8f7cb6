"""K1P (k1_variant=3, persistent warp-specialised K1 for blocks 2-6) against K1: per-block taps, then per-kernel times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
GOLD = os.path.join(ROOT, "tests", "golden")
crops = np.concatenate([np.load(os.path.join(GOLD, "sample_crops.npy")), np.load(os.path.join(GOLD, "jitter_crops.npy"))])
crops = np.concatenate([crops] * 4)[:int(os.environ.get("NT", "8"))]
N = int(os.environ.get("N", "256"))
m = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision=os.environ.get("PREC", "bf16"), max_batch=N)
m.set_option("streams", 1)
def taps(variant):
    m.set_option("k1_variant", variant)
    m.enable_taps(True)
    ang = np.stack(m.get_angle(crops), axis=1)
    out = {"angles": ang}
    for i in range(2, 8):
        for kind in ("dw", "gate", "block"):
            out["%s%d" % (kind, i)] = m.tap("%s%d" % (kind, i)).astype(np.float64)
    m.enable_taps(False)
    return out
ref = taps(1)
try:
    got = taps(3)
    for k in ref:
        d = np.abs(got[k] - ref[k])
        print("%-8s max|d| %.3e  (ref max %.3e, mismatching %d / %d, nan %d)" % (k, d.max(), np.abs(ref[k]).max(), int((d > 0).sum()), d.size, int(np.isnan(got[k]).sum())), flush=True)
except Exception as e:
    print("K1P taps FAILED:", e, flush=True)
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
    for ew in (8, 4):
        m.set_option("k1p_epi_warps", ew)
        b = prof(3)
        for i in range(2, 7):
            k = "b%02d.k1" % i
            print("epi_warps %d  %s  K1 %.4f ms   K1P %.4f ms   x%.2f" % (ew, k, a[k], b[k], a[k] / b[k]), flush=True)
        print("epi_warps %d  total ms/forward: K1 %.3f  K1P %.3f" % (ew, sum(a.values()), sum(b.values())), flush=True)
        chk = np.stack(m.get_angle(crops), axis=1)
        print("epi_warps %d  angles equal to K1's: %s" % (ew, bool(np.array_equal(chk, ref["angles"]))), flush=True)
    m.set_option("k1p_epi_warps", 8)
    full = np.stack(m.get_angle(x[:64]), axis=1)
    one = np.stack(m.get_angle(x[5:6]), axis=1)
    print("batch invariance (64 vs 1):", bool(np.array_equal(full[5], one[0])))
except Exception as e:
    print("K1P timing FAILED:", e, flush=True)
