"""Tuning aid: ms per plan of the single-problem configs (C3 Panda 400 iterations, C1 planar fp64 500 iterations).
usage: python scratch/time_plan.py [env=val ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for kv in sys.argv[1:]:
    k, _, v = kv.partition("=")
    os.environ[k] = v
import torch
import bench
dev = torch.device('cuda:0')
for name, iters in (("panda", 400), ("planar_shipped", 500)):
    w = bench.workload(name, 1)
    pl = bench.build_planner(w, 1, dev)
    obs = {"obstacle_spheres": torch.tensor(w["spheres"], dtype=torch.float32, device=dev)} if w["spheres"] is not None else {}
    pl.optimize(opt_iters=3, return_samples=False, **obs)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = pl.optimize(opt_iters=iters, return_samples=False, **obs)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-15s %s  %d iterations: %.3f ms per plan (%.2f us / iteration)  mean cost %.6e" % (name, " ".join(sys.argv[1:]), iters, best, best / iters * 1e3,
                                                                                 float(out[4].double().mean())), flush=True)
