"""Where does a single-iteration optimize() call spend its time?  (the reference's usage: a Python loop of calls)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import runpy, torch
# build the planner of the shipped example without running its loop
src = open(os.path.join(os.path.dirname(__file__), "..", "examples", "panda_environment.py")).read()
head = src[:src.index("    opt_iters = 400")] if "    opt_iters = 400" in src else None
ns = {"__name__": "__main__", "__file__": "examples/panda_environment.py"}
exec(compile(head, "panda_head", "exec"), ns)
planner, obs = ns["planner"], ns["obs"]
for mode in ("1", "0"):
    os.environ["SGPMP_LOWLAT"] = mode
    for _ in range(20):
        planner.optimize(**obs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(400):
        planner.optimize(**obs)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("LOWLAT=%s: host enqueue %.1f us/call, gpu %.1f us/call, wall incl. sync %.1f us/call" % (
        mode, (t1 - t0) / 400 * 1e6, e0.elapsed_time(e1) / 400 * 1e3, (t2 - t0) / 400 * 1e6), flush=True)
