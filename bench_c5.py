#!/usr/bin/env python
"""C5 (BASELINE.json configs[4]): prior-construction and sampling sweep, traj_len 64..1024 x n_dof 2..14, fp64, and the
split-particle mode (one problem's samples divided over ranks, one all_gather of per-particle statistics per iteration).

    python bench_c5.py                                   # single GPU sweep: K1 prior factor + K2 sampling, fp64
    torchrun --nproc-per-node R bench_c5.py --split      # split-particle optimize() iteration at R ranks vs one rank

One JSON line per measurement (CUDA events, median of 5 after 2 warm-ups).  The reference cannot run this regime at all
beyond M ~ 4096: its dense [NP, M, M] precision is 6.6 GB at T = 1024, n = 14.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def timeit(torch, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def sweep():
    import torch
    from stoch_gpmp_b200 import ops
    from stoch_gpmp_b200.planner import prior_blocks
    dev = torch.device("cuda", 0)
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6545.6) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6545.6
    NP, S = 4, 512
    for T in (64, 128, 256, 512, 1024):
        D, O = prior_blocks(T, 0.02, 1e-3, 3.0, 1e-3)            # planar sampling prior of the shipped example (PD-sensitive)
        Dt = torch.tensor([D], dtype=torch.float64, device=dev)
        Ot = torch.tensor([O], dtype=torch.float64, device=dev)
        tables, bad = ops.prior_factor(Dt, Ot)
        assert int(bad[0]) == 0
        k1 = timeit(torch, lambda: ops.prior_factor(Dt, Ot))
        for n in (2, 4, 7, 14):
            d = 2 * n
            B = max(1, min(64, (1 << 28) // (NP * S * T * d * 8)))           # keep the sample buffer <= 256 MiB... x B problems
            sh = ops.make_shape(B, NP, 1, S, T, n, torch.float64)
            means = torch.zeros(B, NP, T, d, dtype=torch.float64, device=dev)
            x = ops.sample(sh, tables[0].contiguous(), means, seed=1, draw=0)
            var = float(x[0, 0, T // 2, 0].var())
            ms = timeit(torch, lambda: ops.sample(sh, tables[0].contiguous(), means, seed=1, draw=0))
            nbytes = B * NP * S * T * d * 8
            print(json.dumps({"config": "C5", "T": T, "n_dof": n, "M": T * d, "dtype": "f64", "problems": B, "particles": NP, "samples": S,
                              "k1_prior_factor_ms": k1, "k2_sample_ms": ms, "k2_written_gbs": nbytes / (ms * 1e-3) / 1e9,
                              "k2_frac_of_hbm": nbytes / (ms * 1e-3) / 1e9 / hbm, "mid_state_sample_variance": var,
                              "reference_dense_precision_bytes_per_particle": (T * d) ** 2 * 8}), flush=True)


def split():
    import torch
    import torch.distributed as dist
    import bench
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    for B in (1, 64):
        w = bench.workload("panda", B)
        pl = bench.build_planner(w, B, dev)
        obs = {"obstacle_spheres": torch.tensor(w["spheres"], dtype=torch.float32, device=dev)}
        iters = 20

        def run_split():
            pl.optimize_split(opt_iters=iters, **obs)

        def run_single():      # the same separate-kernel iteration on one rank (sample -> cost -> local stats -> apply), all S samples
            for _ in range(iters):
                pl.sample_and_eval(**obs)
        ms_split = timeit(torch, run_split, reps=3, warm=1) / iters
        t = torch.tensor([ms_split], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms_one = timeit(torch, run_single, reps=3, warm=1) / iters
            print(json.dumps({"config": "C5 split-particle", "workload": w["name"], "problems": B, "ranks": world, "samples_per_rank": w["S"] // world,
                              "ms_per_iteration_split": float(t[0]), "ms_per_iteration_one_rank_sample_plus_cost": ms_one,
                              "exchange": "one all_gather of NP*(M+2) reals per iteration (%d bytes per problem)" % (w["G"] * w["K"] * (w["T"] * 2 * w["n_dof"] + 2) * 4)}), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--split", action="store_true")
    args = ap.parse_args()
    split() if args.split else sweep()
