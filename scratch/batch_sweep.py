"""Per-iteration latency / throughput of the fused Panda loop as a function of the number of problems B
(shows where the thread-block-cluster modes take over).  python scratch/batch_sweep.py"""
import sys
import torch
sys.path.insert(0, '.')
import bench

dev = torch.device('cuda:0')
print('%6s %10s %14s %16s' % ('B', 'ms/iter', 'traj-samples/s', 'us/iter/problem'))
for B in (list(map(int, sys.argv[1:])) or [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]):
    w = bench.workload('panda', B)
    pl = bench.build_planner(w, B, dev)
    obs = {'obstacle_spheres': torch.tensor(w['spheres'], dtype=torch.float32, device=dev)}
    iters = 50 if B <= 256 else 10
    pl.optimize(opt_iters=3, return_samples=False, **obs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pl.optimize(opt_iters=iters, return_samples=False, **obs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print('%6d %10.4f %14.4e %16.2f' % (B, ms, B * 4 * 512 / (ms * 1e-3), ms * 1e3 / B))
