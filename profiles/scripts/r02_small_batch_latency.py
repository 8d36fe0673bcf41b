"""Latency of ONE fclb_distance_batch_dev call (C2 mix, f32, inputs resident) against the batch size: wall clock around the
call (it returns after the device has finished), median of 20 calls after 5 warm-up calls."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "mind-fcl_b200"))
import torch
import fclb200 as fclb
import scenes

fclb.init(0)
dev = torch.device("cuda:0")
for n in (1_000, 10_000, 100_000, 1_000_000):
    shapes, pairs, p1, p2 = scenes.config_c2(n, np.float32)
    table = fclb.shapes_upload(shapes)
    d_pairs = torch.from_numpy(pairs.view(np.uint32).reshape(n, 2).view(np.int32)).to(dev)
    d_p1, d_p2 = torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)
    dist = torch.empty(n, dtype=torch.float32, device=dev)
    w1 = torch.empty(n, 3, dtype=torch.float32, device=dev)
    w2 = torch.empty(n, 3, dtype=torch.float32, device=dev)
    ok = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ts = []
    for i in range(25):
        t0 = time.perf_counter()
        fclb.distance_batch_dev(table, d_pairs, d_p1, d_p2, n, fclb.F32, dist, w1, w2, ok)
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[5:])
    print("n = %8d   median %.3f ms   min %.3f ms   %.3e q/s" % (n, 1e3 * ts[len(ts) // 2], 1e3 * ts[0], n / ts[len(ts) // 2]), flush=True)
    fclb.release(table)
