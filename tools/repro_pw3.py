"""pw_tc3 (use_tc=5) vs pw_tc2 (use_tc=2) through the conv1x1 debug hook: where do they differ?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import whenet_b200
def bf(x): return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()
net = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision="bf16", max_batch=8)
for K, N, hw, crops, res in ((32, 16, 12544, 70, False), (32, 16, 12544, 256, False), (96, 24, 3136, 200, False), (144, 24, 3136, 200, True), (144, 40, 784, 600, False)):
    rng = np.random.default_rng(K + N)
    M = crops * hw
    A = bf(rng.standard_normal((M, K))); W = bf(rng.standard_normal((K, N)) / np.sqrt(K)); bias = rng.standard_normal(N).astype(np.float32)
    gate = rng.uniform(0.1, 1.0, (crops, K)).astype(np.float32)
    resid = bf(rng.standard_normal((M, N))) if res else None
    try:
        a = net.debug_conv1x1(A, W, bias, gate=gate, resid=resid, hw=hw, swish=False, use_tc=5)
    except Exception as e:
        print(K, N, hw, crops, "pw_tc3 refused:", str(e)[:100]); continue
    b = net.debug_conv1x1(A, W, bias, gate=gate, resid=resid, hw=hw, swish=False, use_tc=2)
    bad = ~(np.isclose(a, b, rtol=0, atol=0) | (np.isnan(a) & np.isnan(b)))
    rows = np.nonzero(bad.any(axis=1))[0]
    print(K, N, hw, crops, "mismatching rows", rows.size, "of", M, "nan in pw_tc3:", int(np.isnan(a).sum()), flush=True)
    if rows.size:
        t = rows // 128
        print("   tiles-in-crop hist (first 12):", np.unique((rows % hw) // 128, return_counts=True)[0][:12], np.unique((rows % hw) // 128, return_counts=True)[1][:12])
        print("   cols bad:", np.nonzero(bad.any(axis=0))[0][:20], "max abs diff", np.nanmax(np.abs(a - b)), "rows%128 sample", (rows % hw % 128)[:10])
