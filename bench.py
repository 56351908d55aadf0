#!/usr/bin/env python
"""bench.py — StochGPMP hot-path benchmark (driver contract + tier keys: roofline, cpu_baseline, e2e).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload panda|planar]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

metric   trajectory-samples/s per optimize() iteration (BASELINE.json), whole job over all ranks.
step     ONE optimize() iteration of the hot path over the whole problem batch (the reference's examples
         call optimize() with opt_iters=1 in a loop: examples/panda_environment.py:141-146).
workload N=1: BASELINE.json configs[3], "Panda 7-DoF batched: 4096 problems x 4 goals x 512 particles x
         traj_len 64" (fits one B200).  N>1: the same 4096 problems PER GPU (weak scaling), sharded by
         problem with no data-path collective; RNG streams keyed by global problem ids.
value    inputs resident in HBM, fused kernel only (CUDA events around each step, L2 flushed between steps).
e2e      same metric through the public API (StochGPMPBatch.optimize) from HOST buffers: per step the
         observation (obstacle spheres) is copied H2D from pinned memory and the plan (particle means) is
         read back D2H, both inside the timed region (host clock).  The batch is sharded into --e2e-shards
         StochGPMPBatch objects on their own streams so that a shard's D2H overlaps the next shard's kernel;
         the host waits for every shard at the end of every step.
The reference arm (--impl reference) times the reference's CPU algorithm (oracle/reference_port.py, pinned to
the real reference by tests/test_reference_port.py; the reference itself is pure Python and does not travel
to the GPU box) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "trajectory_samples_per_s_per_iteration"
UNIT = "traj-samples/s"


# ------------------------------------------------------------------------------------------------ workload
def workload(name, B):
    """Shapes, sigmas and synthetic inputs (C4 Panda batched / C2 planar batched)."""
    import numpy as np
    from stoch_gpmp_b200 import scenarios as sc
    if name == "panda":
        start, goals, spheres = sc.panda_batch(B, G=4, O=5, seed0=0)
        return dict(name="panda_7dof_batched", n_dof=7, T=64, dt=0.05, G=4, K=1, S=512, temperature=1.0, step_size=0.1,
                    start=start, goals=goals, spheres=spheres, sig=sc.PANDA_SIGMAS, cost=sc.PANDA_COST, n_links=11, O=5)
    if name == "planar":
        start, goals = sc.planar_batch(B, G=4, seed0=0)
        return dict(name="planar_2dof_batched", n_dof=2, T=64, dt=0.02, G=4, K=1, S=256, temperature=1.0, step_size=0.5,
                    start=start, goals=goals, spheres=None, sig=sc.PLANAR_SIGMAS, cost=sc.PLANAR_COST, n_links=0, O=0)
    raise SystemExit("unknown workload %r" % name)


def algorithmic_flops_per_traj(w):
    """ALGORITHMIC FP32 flops per trajectory-sample-iteration (FMA = 2), SURVEY §8(d) constants:
    banded sampling 16 T n; GP factor 10 (T-1) n; start+goal 6 d; IS dot 2 M; softmax/update 2 M + 10;
    Panda: FK (specialised z-axis chain) (7*36 + 3*18)(T-1); sphere RBF 10 L O (T-1);
    planar: map index arithmetic 10 (T-1).  MUFU ops (sin/cos, exp, Box-Muller) are counted separately."""
    T, n = w["T"], w["n_dof"]
    d, M = 2 * n, 2 * n * T
    f = 16 * T * n + 10 * (T - 1) * n + 6 * d + 2 * M + 2 * M + 10
    mufu = 0
    if w["spheres"] is not None:
        f += (7 * 36 + 3 * 18) * (T - 1) + 10 * w["n_links"] * w["O"] * (T - 1)
        mufu += 14 * (T - 1) + w["n_links"] * w["O"] * (T - 1)
    else:
        f += 10 * (T - 1)
    mufu += 2 * M          # Box-Muller: ~2 MUFU-class ops per normal (log, sqrt, sin, cos shared by a pair)
    return f, mufu


def build_planner(w, B, dev, problem_offset=0, seed=0):
    import torch
    from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoalPrior
    from stoch_gpmp_b200.costs.fields import LinkDistanceField
    from stoch_gpmp_b200.planner import StochGPMPBatch
    from stoch_gpmp_b200.robots import PandaFK
    ta = {"device": dev, "dtype": torch.float32}
    n, T, K, S = w["n_dof"], w["T"], w["K"], w["S"]
    s = torch.tensor(w["start"], **ta)
    g = torch.tensor(w["goals"], **ta)
    cl = [CostGP(n, T, s, w["dt"], dict(sigma_start=w["cost"]["sigma_start"], sigma_gp=w["cost"]["sigma_gp"]), ta),
          CostGoalPrior(n, T, multi_goal_states=g, num_particles_per_goal=K, num_samples=S,
                        sigma_goal_prior=w["cost"]["sigma_goal_prior"], tensor_args=ta)]
    FK = None
    if w["spheres"] is not None:
        FK = PandaFK()
        cl.append(CostCollision(n, T, field=LinkDistanceField(tensor_args=ta), sigma_coll=w["cost"]["sigma_coll"]))
    else:
        import random
        import numpy as np
        from stoch_gpmp_b200.envs.map_generator import generate_obstacle_map
        maps = []
        for b in range(16):     # 16 distinct 200x200 maps cycled over the problems
            random.seed(1000 + b)
            np.random.seed(1000 + b)
            maps.append(generate_obstacle_map(map_dim=[20, 20], cell_size=0.1, random_gen=True, num_obst=15,
                                              rand_limits=[[-7.5, 7.5], [-7.5, 7.5]], rand_rect_shape=[2, 2], tensor_args=ta)[0])
        cl.append(CostCollision(n, T, field=[maps[b % 16] for b in range(B)], sigma_coll=w["cost"]["sigma_coll"]))
    cost = CostComposite(n, T, cl, FK=FK, tensor_args=ta)
    return StochGPMPBatch(num_particles_per_goal=K, num_samples=S, traj_len=T, opt_iters=1, dt=w["dt"], n_dof=n,
                          step_size=w["step_size"], temperature=w["temperature"], start_state=s, multi_goal_states=g,
                          initial_particle_means="const_vel", cost=cost, seed=seed, tensor_args=ta,
                          problem_offset=problem_offset, **w["sig"])


# -------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 9 for k in range(4) if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_time(w, n_problems, iters, warmup):
    """Seconds per iteration of ONE problem for the reference's CPU algorithm (oracle/reference_port.py)."""
    import numpy as np
    import torch
    from oracle import fk as OFK
    from oracle.reference_port import time_port
    spec = dict(n_dof=w["n_dof"], T=w["T"], dt=w["dt"], G=w["G"], K=w["K"], S=w["S"], temperature=w["temperature"],
                step_size=w["step_size"], start=np.asarray(w["start"][0]), goals=np.asarray(w["goals"][0]),
                cost_sigma_start=w["cost"]["sigma_start"], cost_sigma_gp=w["cost"]["sigma_gp"],
                sigma_goal_prior=w["cost"]["sigma_goal_prior"], sigma_coll=w["cost"]["sigma_coll"], **w["sig"])
    fk = None
    dtype = torch.float32
    if w["spheres"] is not None:
        spec["spheres"] = np.asarray(w["spheres"][0])
        fk = OFK.fk_all_links_torch()
    else:
        spec["sigma_coll"] = None        # planar CPU sample: GP + goal factors (the reference needs fp64 here)
        dtype = torch.float64
    return time_port(spec, dtype, n_problems, iters, warmup, fk=fk)


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = workload(args.workload, 1)
    t0 = time.time()
    sec, cores = cpu_reference_time(w, 1, args.steps, args.warmup)
    ntraj = w["G"] * w["K"] * w["S"]
    val = ntraj / sec
    sample = "%d timed optimize() iterations (+%d warm-up) of 1 problem of the workload (G=%d,K=%d,S=%d,T=%d,n=%d), %s" % (
        args.steps, args.warmup, w["G"], w["K"], w["S"], w["T"], w["n_dof"], "fp32" if w["spheres"] is not None else "fp64")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if w["spheres"] is not None else "f64", "data": "synthetic",
            "config": {"workload": w["name"], "problems_per_step": 1, "goals": w["G"], "particles_per_goal": w["K"] * w["S"],
                       "traj_len": w["T"], "n_dof": w["n_dof"], "note": "reference CPU algorithm (oracle/reference_port.py; the "
                       "pure-Python reference does not travel to the GPU box), torch CPU, all host threads"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def probe_peaks(dev):
    """Measured FP32-FMA and MUFU pipe peaks (the denominators MEASURED_PEAKS.json lacks)."""
    import torch
    from stoch_gpmp_b200 import _lib
    lib = _lib.load()
    scratch = torch.zeros(16, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    blocks = torch.cuda.get_device_properties(dev).multi_processor_count * 8
    for mode, name, per_thread_iter in ((0, "fp32_tflops", 128 * 2), (1, "mufu_tops", 64)):
        iters = 2000
        _lib.check(lib.sgpmp_probe(mode, blocks, 200, ctypes.c_void_p(scratch.data_ptr()), st), "probe")
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.sgpmp_probe(mode, blocks, iters, ctypes.c_void_p(scratch.data_ptr()), st), "probe")
            e1.record()
            torch.cuda.synchronize()
            best = max(best, blocks * 256 * iters * per_thread_iter / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out[name] = best
    return out


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from stoch_gpmp_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B = args.problems_per_gpu
    w = workload(args.workload, B * world)
    lo = rank * B
    wl = dict(w, start=w["start"][lo:lo + B], goals=w["goals"][lo:lo + B],
              spheres=None if w["spheres"] is None else w["spheres"][lo:lo + B])
    pl = build_planner(wl, B, dev, problem_offset=lo, seed=0)
    obs = {}
    sph_host = sph_dev = None
    if w["spheres"] is not None:
        sph_host = torch.tensor(wl["spheres"], dtype=torch.float32).pin_memory()
        sph_dev = sph_host.to(dev)
        obs = {"obstacle_spheres": sph_dev}
    ntraj_rank = B * w["G"] * w["K"] * w["S"]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---------------- warm-up
    for _ in range(max(args.warmup, 3)):
        pl.optimize(return_samples=False, **obs)
    barrier()

    # ---------------- timed: device-resident inputs ("value")
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    evs = []
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1.0)                                  # L2 flush (126 MB L2 < 256 MiB), outside the event pair
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pl.optimize(return_samples=False, **obs)
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    launches = _lib.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)

    # ---------------- timed: end to end from host buffers ("e2e")
    # Public API only: the batch is sharded into E2E_SHARDS StochGPMPBatch objects (problem_offset keeps every RNG stream where it
    # was, results are bit-identical to the single batch: tests/test_gpu_planner.py::test_fused_problem_sharding_invariance), one
    # CUDA stream each.  Per step every shard copies its inputs H2D, runs optimize() and copies its plan D2H on its own stream, so
    # the 58.7 MB D2H of shard k rides under the kernel of shard k+1; the host waits for ALL shards before the next step starts
    # ("the user has the plan on the host every step").
    n_sh = max(1, min(args.e2e_shards, B))
    while B % n_sh:
        n_sh -= 1
    Bs = B // n_sh
    from stoch_gpmp_b200.parallel import StreamShards
    planners, host_in = [], []
    for k in range(n_sh):
        a = k * Bs
        wk = dict(w, start=wl["start"][a:a + Bs], goals=wl["goals"][a:a + Bs], spheres=None if wl["spheres"] is None else wl["spheres"][a:a + Bs])
        planners.append(build_planner(wk, Bs, dev, problem_offset=lo + a, seed=0) if n_sh > 1 else pl)
        if wl["spheres"] is not None:
            host_in.append(torch.tensor(wk["spheres"], dtype=torch.float32).pin_memory())
        else:      # planar: the per-step host input is the start/goal set of every problem
            host_in.append(torch.tensor(np.concatenate([wk["start"].reshape(Bs, -1), wk["goals"].reshape(Bs, -1)], 1), dtype=torch.float32).pin_memory())
    shards = StreamShards(planners, obs_key="obstacle_spheres" if wl["spheres"] is not None else None, host_inputs=host_in)
    h2d = sum(h.numel() for h in host_in) * 4
    d2h = sum(m.numel() for m in shards.host_means) * 4

    def e2e_step():
        shards.step()
        shards.wait()                                       # the user has the whole plan on the host every step

    torch.cuda.synchronize()
    for _ in range(3):
        e2e_step()
    barrier()
    launches_e2e0 = _lib.launch_count()
    e2e_wall0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_wall = time.perf_counter() - e2e_wall0               # host clock: every step ends with a host-side wait on all streams
    launches_e2e = _lib.launch_count() - launches_e2e0
    barrier()
    e2e_ms = e2e_wall * 1e3

    # ---------------- max over ranks
    t = torch.tensor([dev_ms, e2e_ms, wall * 1e3, e2e_wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms, e2e_wall_ms = [float(x) for x in t.tolist()]
    if rank != 0:
        return
    ms_per_step = dev_ms / args.steps
    value = ntraj_rank * world / (ms_per_step * 1e-3)
    e2e_value = ntraj_rank * world * args.steps / (max(e2e_ms, e2e_wall_ms) * 1e-3)

    # ---------------- roofline of the dominant kernel (the fused iteration kernel: one launch per step)
    peaks = probe_peaks(dev)
    measured = {}
    try:
        measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic, ncu_facts = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1", "traffic.json"))).get(w["name"])
        if tj and tj["problems"] == B:
            traffic = tj["dram_read_bytes"] + tj["dram_write_bytes"]
            ncu_facts = {k: tj[k] for k in ("thread_instructions_per_traj_sample", "issue_active_pct", "pipe_fma_pct", "pipe_alu_pct", "pipe_xu_pct")}
    except Exception:
        pass
    flops, mufu = algorithmic_flops_per_traj(w)
    ach_tf = ntraj_rank * flops / (ms_per_step * 1e-3) / 1e12
    ach_mufu = ntraj_rank * mufu / (ms_per_step * 1e-3) / 1e12
    roof = {"bound": "fp32", "kernel": "sgpmp::iterate_kernel<float,2,%d,%d,%d,1>" % (w["n_dof"], 128 if (w["spheres"] is None and B * w["G"] * w["K"] >= 10 * 148) else 256, 1 if w["spheres"] is not None else 0), "achieved": ach_tf,
            "peak": peaks["fp32_tflops"], "unit": "TFLOP/s", "frac": ach_tf / peaks["fp32_tflops"], "traffic": traffic,
            "traffic_source": "profiles/r1/traffic.json (ncu --set full capture of this command)" if traffic else None,
            "ncu": ncu_facts,
            "peak_source": "FP32 FMA probe kernel timed in this run (MEASURED_PEAKS.json has no FP32-pipe entry; nominal 74.4)",
            "algorithmic_flops_per_traj_sample": flops, "launch_ms": ms_per_step,
            "mufu": {"achieved_tops": ach_mufu, "peak_tops": peaks["mufu_tops"], "frac": ach_mufu / peaks["mufu_tops"],
                     "algorithmic_mufu_per_traj_sample": mufu},
            "binding": ("mufu" if mufu / peaks["mufu_tops"] > flops / peaks["fp32_tflops"] else "fp32"),
            "binding_note": "lower-bound time = max(flops/FP32 peak, MUFU ops/MUFU peak) from the ALGORITHMIC counts of SURVEY 8(d) "
                            "(+2 MUFU per Box-Muller normal); the implementation trades MUFU for FMA-pipe polynomials and "
                            "structural savings, so the executed pipe utilisations are the ncu figures",
            "hbm_note": "fused kernel: HBM traffic is O(B*NP*M) per step; materialised 3-kernel dataflow would move %.1f GB/step"
                        % (3 * 2 * w["n_dof"] * w["T"] * 4 * ntraj_rank / 1e9),
            "measured_hbm_gbs": measured.get("hbm_gbs")}

    # ---------------- CPU baseline (bounded sample, rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sec, cores = cpu_reference_time(w, 2, 2, 1)
        cpu = {"value": w["G"] * w["K"] * w["S"] / sec, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "2 problems x (1 warm-up + 2 timed) optimize() iterations of the workload's per-problem shape, "
                         "run sequentially (the reference has no problem-batch axis); value = NP*S / mean s per iteration",
               "s_per_iteration_per_problem": sec}

    # ---------------- "ms per plan" of ONE problem (the metric's second half): all iterations in one fused launch
    plan_iters = 400 if w["spheres"] is not None else 500
    w1 = dict(w, start=w["start"][:1], goals=w["goals"][:1], spheres=None if w["spheres"] is None else w["spheres"][:1])
    p1 = build_planner(w1, 1, dev, problem_offset=0, seed=0)
    obs1 = {"obstacle_spheres": torch.tensor(w1["spheres"], dtype=torch.float32, device=dev)} if w["spheres"] is not None else {}
    p1.optimize(opt_iters=3, return_samples=False, **obs1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    p1.optimize(opt_iters=plan_iters, return_samples=False, **obs1)
    e1.record()
    torch.cuda.synchronize()
    plan_ms_single = e0.elapsed_time(e1)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": w["name"], "problems_per_gpu": B, "problems_total": B * world, "goals": w["G"],
                       "particles_per_goal": w["K"] * w["S"], "K": w["K"], "S": w["S"], "traj_len": w["T"], "n_dof": w["n_dof"],
                       "traj_samples_per_step": ntraj_rank * world, "opt_iters_per_step": 1, "parallelism": "problem-sharded x%d, no collective" % world,
                       "l2": "256 MiB write between timed steps (outside the CUDA-event pair)", "prior": "fp64 factor, fp32 run-time",
                       "rng": "in-kernel Philox4x32-10"},
            "clocks": clk, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                                   "ms_per_step": max(e2e_ms, e2e_wall_ms) / args.steps, "shards_per_gpu": n_sh,
                                   "launches_per_step": launches_e2e / args.steps,
                                   "note": "StochGPMPBatch per shard on its own stream: H2D inputs, optimize(), D2H plan; host waits for all shards every step"},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "wall_ms_per_step_incl_flush": wall_ms / args.steps,
            "ms_per_plan": {"iterations": plan_iters, "single_problem_ms": plan_ms_single,
                            "amortised_over_batch_ms": ms_per_step * plan_iters / B,
                            "note": "single problem = one StochGPMP (B=1), all iterations enqueued by one optimize() call (low-latency three-kernel form, csrc/sgpmp_lowlat.cu)"}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="panda", choices=["panda", "planar"])
    ap.add_argument("--problems-per-gpu", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-shards", type=int, default=8, help="StochGPMPBatch shards (CUDA streams) per GPU in the end-to-end measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
