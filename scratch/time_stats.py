"""Tuning aid: time the fused statistics launch of the split-particle mode (sgpmp_iterate_stats) for a rank's share of the samples.
usage: python scratch/time_stats.py S_local [env=val ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
S_loc = int(sys.argv[1])
for kv in sys.argv[2:]:
    k, _, v = kv.partition("=")
    os.environ[k] = v
import torch
import bench
from stoch_gpmp_b200 import ops
dev = torch.device('cuda:0')
B = 64
w = bench.workload("panda", B)
pl = bench.build_planner(w, B, dev)
obs = {"obstacle_spheres": torch.tensor(w["spheres"], dtype=torch.float32, device=dev)}
desc = pl._desc(obs)
sh = ops.make_shape(B, w["G"], w["K"], S_loc, w["T"], w["n_dof"], torch.float32, 0, sample_gid0=0)
for _ in range(3):
    ops.iterate_stats(sh, desc, pl._tables, pl._means, 1, 0)
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(10):
        ops.iterate_stats(sh, desc, pl._tables, pl._means, 1, it)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
print("stats launch: 64 problems, S_local=%d %s: %.4f ms" % (S_loc, " ".join(sys.argv[2:]), best), flush=True)
