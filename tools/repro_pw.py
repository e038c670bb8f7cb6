"""Repro helper: pw_tc2 (use_tc=2) vs K2 (use_tc=3) on one gated late-project shape, several times; prints where NaNs (unwritten rows) appear."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import whenet_b200
def bf(x):
    import torch
    return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()
net = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision="bf16", max_batch=8)
K, N, hw = 1152, 320, 49
rng = np.random.default_rng(K + N)
M = 2 * 148 * 128 + 3 * 128 + 77
A = bf(rng.standard_normal((M, K))); W = bf(rng.standard_normal((K, N)) / np.sqrt(K))
bias = rng.standard_normal(N).astype(np.float32)
gate = rng.uniform(0.1, 1.0, ((M + hw - 1) // hw, K)).astype(np.float32)
order = [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "3232322").strip()]
for fam in order:
    o = net.debug_conv1x1(A, W, bias, gate=gate, resid=None, hw=hw, swish=False, use_tc=fam)
    bad = np.isnan(o).any(axis=1)
    rows = np.nonzero(bad)[0]
    print("family", fam, "nan rows", int(bad.sum()), (int(rows.min()), int(rows.max())) if rows.size else None,
          "tiles", sorted(set((rows // 128).tolist()))[:20], flush=True)
