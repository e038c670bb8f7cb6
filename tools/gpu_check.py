"""First-contact GPU script: parity summary + rough timings (not a bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import whenet_b200
from whenet_oracle import load_oracle
GOLD = os.path.join(ROOT, "tests", "golden"); SNAP = os.path.join(ROOT, "headposeestimation-whenet_b200", "data", "whenet_weights.npz")
crops = np.load(os.path.join(GOLD, "sample_crops.npy"))
o = load_oracle(SNAP, np.float32); taps = {}
ref = np.stack(o.get_angle(crops, taps), axis=1)
for prec in os.environ.get("PRECS", "fp32,bf16,fp16").split(","):
    for tc in ((0,) if prec == "fp32" else tuple(int(t) for t in os.environ.get("TCS", "0,1").split(","))):
        m = whenet_b200.WHENet(SNAP, device=0, precision=prec, max_batch=512)
        m.set_option("tensor_cores", tc)
        for kv in filter(None, os.environ.get("OPTS", "").split(",")):
            k, v = kv.split("=")
            m.set_option(k, int(v))
        if os.environ.get("CHUNK"):
            m.set_option("chunk", int(os.environ["CHUNK"]))
        m.enable_taps(True)
        got = np.stack(m.get_angle(crops), axis=1)
        m.enable_taps(False)
        print(prec, "tc=%d" % tc, "angles", got.round(4).tolist(), "max|d|", float(np.abs(got - ref).max()))
        for nm in ["stem", "dw1", "gate1", "block1", "dw2", "block2", "block5", "block11", "block16", "head", "pooled"]:
            r = taps[nm].reshape(-1).astype(np.float64); g = m.tap(nm).astype(np.float64)
            print("   %-8s rel-max %.3e rel-rms %.3e" % (nm, np.abs(g - r).max() / np.abs(r).max(), np.sqrt(((g - r) ** 2).mean() / (r ** 2).mean())))
        rng = np.random.default_rng(0)
        NP = int(os.environ.get("NPROF", "512"))
        for n in (32, NP):
            x = rng.integers(0, 256, (n, 224, 224, 3), dtype=np.uint8)
            m.get_angle(x)
            t = time.time(); m.get_angle(x); dt = time.time() - t
            print("   N=%d host-e2e %.2f ms -> %.0f crops/s" % (n, dt * 1e3, n / dt))
        m.enable_profile(True); m.get_angle(x); st = m.read_profile(); m.enable_profile(False)
        tot = sum(s["ms"] for s in st)
        print("   profile N=%d total kernel ms %.3f -> %.0f crops/s device" % (NP, tot, NP / tot * 1e3))
        for s in (st if os.environ.get("FULL") else sorted(st, key=lambda s: -s["ms"])[:12]):
            print("     %-16s %.3f ms  %.1f GB/s  %.2f TFLOP/s" % (s["name"], s["ms"], s["bytes"] / s["ms"] / 1e6, s["flops"] / s["ms"] / 1e9))
        m.close()
