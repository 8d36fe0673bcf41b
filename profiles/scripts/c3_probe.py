import sys, time
sys.path.insert(0, "mind-fcl_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import fclb200 as fclb, scenes, bench
wl = bench.make_workload("c3", 1_000_000, "f32", seed=0)
dev = torch.device("cuda:0")
wl.setup(fclb, torch, dev)
for i in range(6):
    torch.cuda.synchronize(); t=time.perf_counter()
    fclb.bvh_collide_batch_dev(wl.handles[0], wl.handles[1], wl.d_p1, wl.d_p2, wl.n, wl.st, wl.req, wl.d_out)
    t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    print("call %.3f ms (+sync %.3f)  kernel %.3f ms" % ((t1-t)*1e3, (t2-t1)*1e3, fclb.last_kernel_ms()), fclb.bvh_last_visit_counts())
