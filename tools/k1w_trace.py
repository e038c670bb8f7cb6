"""Per-role wait breakdown of the K1W launches (option k1w_trace): which stage of the pipeline each role waits for."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
N = int(os.environ.get("N", "256"))
m = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision="bf16", max_batch=N)
m.set_option("streams", 1); m.set_option("k1_variant", 4)
for kv in os.environ.get("OPTS", "").split(","):
    if "=" in kv:
        k, v = kv.split("="); m.set_option(k, int(v))
x = np.random.default_rng(0).integers(0, 256, (N, 224, 224, 3), dtype=np.uint8)
m.get_angle(x)
print("block   CTA kcyc | producer: wait a_empty | MMA: wait a_full  wait t_empty | epilogue: wait t_full  wait e_empty | depthwise: wait e_full  group barrier   (medians over CTAs, % of CTA cycles)")
for blk in [int(b) for b in os.environ.get("BLOCKS", "2,3,4,5,6,7,9,10,12,13,16").split(",")]:
    m.set_option("k1w_trace", blk)
    m.get_angle(x)
    t = m.read_trace(148).astype(np.float64)
    t = t[t[:, 0] > 0]
    tot = np.median(t[:, 0])
    pct = lambda c: 100.0 * np.median(t[:, c] / np.maximum(t[:, 0], 1))
    units = np.maximum(np.median(t[:, 15]), 1)
    print("b%02d   %8.1f |  %5.1f%%               |  %5.1f%%   %5.1f%%          |  %5.1f%%   %5.1f%%              |  %5.1f%%   %5.1f%%   | epilogue warp: %d units, %.0f cyc/unit in wait::ld, %.0f cyc/unit in process" %
          (blk, tot / 1e3, pct(1), pct(4), pct(5), pct(7), pct(8), pct(10), pct(11), units, np.median(t[:, 13]) / units, np.median(t[:, 14]) / units), flush=True)
m.set_option("k1w_trace", 0)
