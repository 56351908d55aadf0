"""Quick A/B timing of the fused loop on the GPU box: python scratch/quick_time.py [panda|planar] [B] [env=val ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for kv in sys.argv[3:]:
    k, _, v = kv.partition("=")
    os.environ[k] = v
import torch
import bench
workload = sys.argv[1] if len(sys.argv) > 1 else "panda"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device('cuda:0')
w = bench.workload(workload, B)
pl = bench.build_planner(w, B, dev)
obs = {'obstacle_spheres': torch.tensor(w['spheres'], dtype=torch.float32, device=dev)} if workload == 'panda' else {}
pl.optimize(opt_iters=2, return_samples=False, **obs)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pl.optimize(opt_iters=5, return_samples=False, **obs)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 5)
print("%-7s B=%d %s  %.3f ms/iter" % (workload, B, " ".join(sys.argv[3:]), best), flush=True)
