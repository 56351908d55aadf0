import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for kv in sys.argv[1:]:
    k, _, v = kv.partition("="); os.environ[k] = v
import torch, bench
dev = torch.device('cuda:0')
for name, iters in (("planar_shipped", 500), ("panda", 400)):
    w = bench.workload(name, 1)
    pl = bench.build_planner(w, 1, dev)
    obs = {'obstacle_spheres': torch.tensor(w['spheres'], dtype=torch.float32, device=dev)} if name == 'panda' else {}
    pl.optimize(opt_iters=3, return_samples=False, **obs); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pl.optimize(opt_iters=iters, return_samples=False, **obs); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(name, " ".join(sys.argv[1:]), "%d iterations: %.2f ms (%.1f us/iter)" % (iters, best, best / iters * 1e3), flush=True)
