"""Host-side cost of one optimize() call (the reference's usage: a Python loop of single-iteration calls)."""
import cProfile
import pstats
import sys
import torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda:0')
w = bench.workload('panda', 1)
pl = bench.build_planner(w, 1, dev)
obs = {'obstacle_spheres': torch.tensor(w['spheres'], dtype=torch.float32, device=dev)}
for _ in range(20):
    pl.optimize(**obs)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(400):
    pl.optimize(**obs)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
