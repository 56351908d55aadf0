import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for kv in sys.argv[1:]:
    k, _, v = kv.partition("="); os.environ[k] = v
import torch
src = open(os.path.join(os.path.dirname(__file__), "..", "examples", "panda_environment.py")).read()
head = src[:src.index("    opt_iters = 400")]
ns = {"__name__": "__main__", "__file__": "examples/panda_environment.py"}
t0 = time.perf_counter()
exec(compile(head, "panda_head", "exec"), ns)
torch.cuda.synchronize()
print("construction: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
planner, obs = ns["planner"], ns["obs"]
for k in range(4):
    t0 = time.perf_counter(); planner.optimize(**obs); torch.cuda.synchronize()
    print("call %d: %.2f ms" % (k, (time.perf_counter() - t0) * 1e3), flush=True)
