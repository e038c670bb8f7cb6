"""Time K1W (k1_variant=4) tile plans per block (tuning aid): CUDA-event time of the block's K1 launch at N crops."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
N = int(os.environ.get("N", "256"))
m = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision="bf16", max_batch=N)
m.set_option("streams", 1); m.set_option("k1_variant", 4)
x = np.random.default_rng(0).integers(0, 256, (N, 224, 224, 3), dtype=np.uint8)
def t_block(blk):
    m.get_angle(x)
    m.enable_profile(True)
    for _ in range(3):
        m.get_angle(x)
    st = m.read_profile(); m.enable_profile(False)
    return {s["name"]: s["ms"] / s["launches"] for s in st}["b%02d.k1" % blk]
# (th, tw, r, cc, nb, n_epi, nt)
CANDS = {
    2: [(8, 8, 4, 48, 1, 8, 768), (8, 8, 4, 48, 1, 4, 768), (8, 4, 4, 96, 1, 8, 768), (8, 8, 4, 32, 1, 8, 768)],
    3: [(14, 14, 7, 48, 1, 4, 768), (14, 14, 7, 48, 1, 8, 768), (14, 14, 4, 48, 1, 4, 768)],
    4: [(7, 7, 4, 48, 1, 8, 768), (7, 7, 4, 48, 1, 4, 768), (7, 7, 7, 48, 1, 8, 768)],
    5: [(14, 14, 7, 80, 1, 4, 768), (14, 14, 7, 48, 1, 4, 768), (14, 14, 7, 80, 1, 8, 768), (14, 14, 4, 80, 1, 4, 768)],
    6: [(7, 7, 4, 80, 1, 8, 768), (7, 7, 4, 80, 1, 4, 768), (7, 7, 7, 80, 1, 4, 768), (7, 7, 4, 48, 1, 8, 768)],
    7: [(14, 14, 7, 96, 1, 4, 768), (14, 14, 7, 96, 1, 8, 768), (14, 14, 7, 160, 1, 4, 768), (14, 14, 4, 96, 1, 4, 768)],
    9: [(14, 14, 7, 48, 1, 4, 768), (14, 14, 7, 48, 1, 8, 768), (14, 14, 4, 48, 1, 4, 768), (14, 14, 7, 32, 1, 4, 768)],
    10: [(14, 14, 7, 48, 1, 4, 768), (14, 14, 7, 48, 1, 8, 768), (14, 14, 4, 48, 1, 4, 768)],
    12: [(7, 7, 4, 48, 1, 4, 768), (7, 7, 4, 48, 1, 8, 768), (7, 7, 7, 48, 1, 4, 768)],
    13: [(7, 7, 4, 64, 2, 4, 768), (7, 7, 7, 64, 2, 4, 768), (7, 7, 4, 128, 1, 4, 768), (7, 7, 4, 48, 2, 4, 768)],
    16: [(7, 7, 7, 96, 2, 4, 768), (7, 7, 7, 192, 1, 4, 768), (7, 7, 4, 64, 2, 4, 768), (7, 7, 4, 128, 1, 4, 768)],
}
for blk in [int(b) for b in os.environ.get("BLOCKS", "2,3,4,5,6,7,9,10,12,13,16").split(",")]:
    base = None
    for c in CANDS[blk]:
        if not m.set_k1w_plan(blk, *c):
            print("block %2d  plan %s  does not fit" % (blk, c), flush=True)
            continue
        try:
            t = t_block(blk)
        except Exception as e:
            print("block %2d  plan %s  FAILED %s" % (blk, c, e), flush=True)
            continue
        base = base or t
        print("block %2d  (th,tw,r,cc,nb,epi,nt)=%-28s %.4f ms  x%.2f" % (blk, c, t, base / t), flush=True)
