"""Tuning aid: time K3 (sgpmp_cost) alone.  usage: SGPMP_LIB=scratch/variants/x.so python scratch/time_cost.py [panda|planar] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stoch_gpmp_b200 import ops
workload = sys.argv[1] if len(sys.argv) > 1 else "planar"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device('cuda:0')
w = bench.workload(workload, B)
pl = bench.build_planner(w, B, dev)
obs = {"obstacle_spheres": torch.tensor(w["spheres"], dtype=torch.float32, device=dev)} if w["spheres"] is not None else {}
desc, sh, tab = pl._desc(obs), pl._shape(), pl._tables
means = pl._means.clone()
xs = ops.sample(sh, tab, means, seed=1, draw=0)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
ts = []
for rep in range(13):
    flush.fill_(0.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c = ops.cost(sh, desc, tab, xs, means)
    e1.record()
    torch.cuda.synchronize()
    if rep >= 3:
        ts.append(e0.elapsed_time(e1))
ts.sort()
ntraj = B * w["G"] * w["K"] * w["S"]
M = w["T"] * 2 * w["n_dof"]
ms = ts[len(ts) // 2]
print("%-22s %-7s B=%d  K3 %.4f ms  %.0f GB/s  checksum %.6e" % (os.path.basename(os.environ.get("SGPMP_LIB", "default")), workload, B, ms,
                                                         ntraj * (M + 1) * 4 / ms / 1e6, float(c.double().sum())), flush=True)
