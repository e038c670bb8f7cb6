"""BASELINE configs[0]/[1]: single Sample crop (N=1) and N=32 latency, per precision, with and without CUDA-graph replay.
Prints one JSON object; angles are checked against the committed float64 golden values."""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import whenet_b200
GOLD = os.path.join(ROOT, "tests", "golden")
crops = np.load(os.path.join(GOLD, "sample_crops.npy"))
g = json.load(open(os.path.join(GOLD, "golden.json")))["samples"]
ref = np.array([[s["yaw"], s["pitch"], s["roll"]] for s in g])
out = {}
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for prec in ("fp32", "bf16", "fp16"):
    m = whenet_b200.WHENet(whenet_b200.weights.DEFAULT_NPZ, device=0, precision=prec, max_batch=32)
    m.set_stream(stream.cuda_stream)
    ang = np.stack(m.get_angle(crops), axis=1)
    res = {"angles": ang.round(4).tolist(), "max_abs_err_deg_vs_oracle64": float(np.abs(ang - ref).max())}
    for n in (1, 32):
        x = torch.from_numpy(np.repeat(crops[:1], n, axis=0)).cuda()
        y = torch.empty((n, 3), dtype=torch.float32, device="cuda")
        for graph in (0, 1):
            m.set_option("graph", graph)
            for _ in range(20):
                m.forward_device(x, y)
            torch.cuda.synchronize()
            ts = []
            for _ in range(200):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record(stream); m.forward_device(x, y); e1.record(stream)
                torch.cuda.synchronize()
                ts.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
            res["n%d_graph%d" % (n, graph)] = {"device_ms_median": statistics.median(t[0] for t in ts),
                                                "wall_ms_median": statistics.median(t[1] for t in ts),
                                                "crops_per_s": n / (statistics.median(t[1] for t in ts) * 1e-3)}
            assert np.array_equal(y.cpu().numpy()[0], ang[0]), "graph replay changed the result"
        m.set_option("graph", 0)
    # host path, as the reference's demo.py would call it
    for _ in range(5):
        m.get_angle(crops[:1])
    ts = []
    for _ in range(100):
        t0 = time.perf_counter(); m.get_angle(crops[:1]); ts.append((time.perf_counter() - t0) * 1e3)
    res["n1_get_angle_host_ms_median"] = statistics.median(ts)
    out[prec] = res
    m.close()
print(json.dumps(out))
