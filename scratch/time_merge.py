"""Tuning aid: time the merge + update launch of the split-particle mode (sgpmp_merge_apply_stats) for R gathered rank blocks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stoch_gpmp_b200 import ops
dev = torch.device('cuda:0')
B = 64
w = bench.workload("panda", B)
pl = bench.build_planner(w, B, dev)
NP, T, d = w["G"] * w["K"], w["T"], 2 * w["n_dof"]
sh = ops.make_shape(B, w["G"], w["K"], 64, T, w["n_dof"], torch.float32)
for R in (1, 2, 8):
    stats = torch.randn(R, B, NP, T * d + 2, device=dev).abs() * 1e-3
    means = pl._means.clone()
    for _ in range(3):
        ops.merge_apply_stats(sh, pl._tables, 0.1, stats, means)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(10):
            ops.merge_apply_stats(sh, pl._tables, 0.1, stats, means)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print("merge + update launch: 64 problems, R=%d: %.4f ms (incl. ~10 us of host per call)" % (R, best), flush=True)
