#!/usr/bin/env python
"""Per-kernel roofline measurements of the standalone (materialising) path: K1 prior, K2 sample, K3 cost,
K4 update, and the split-mode statistics kernels — the numbers DESIGN.md §4 quotes next to the fused kernel.

    python bench_kernels.py [--workload panda|planar] [--problems 1024] [--reps 10]

Prints one JSON line per kernel: algorithmic bytes / flops per launch, CUDA-event time
(median of --reps, L2 flushed between launches), achieved GB/s or TFLOP/s and the fraction of the measured
peak (MEASURED_PEAKS.json HBM copy bandwidth; FP32 FMA probe for the FP32-bound kernel).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="panda", choices=["panda", "planar"])
    ap.add_argument("--problems", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    import bench
    from stoch_gpmp_b200 import ops

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B = args.problems
    w = bench.workload(args.workload, B)
    pl = bench.build_planner(w, B, dev)
    n, T, G, K, S = w["n_dof"], w["T"], w["G"], w["K"], w["S"]
    NP, d = G * K, 2 * n
    M = T * d
    ntraj = B * NP * S
    obs = {"obstacle_spheres": torch.tensor(w["spheres"], dtype=torch.float32, device=dev)} if w["spheres"] is not None else {}
    desc = pl._desc(obs)
    sh = pl._shape()
    tab = pl._tables
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    peaks = bench.probe_peaks(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def timeit(fn):
        ts = []
        for _ in range(3):
            fn()
        for _ in range(args.reps):
            flush.fill_(0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    def report(name, ms, bytes_=None, flops=None, note=""):
        line = {"kernel": name, "workload": w["name"], "problems": B, "traj_samples": ntraj, "ms": ms, "note": note}
        if bytes_ is not None:
            gbs = bytes_ / (ms * 1e-3) / 1e9
            line.update(bound="hbm", algorithmic_bytes=bytes_, achieved_gbs=gbs, peak_gbs=hbm, frac=gbs / hbm)
        if flops is not None:
            tf = flops / (ms * 1e-3) / 1e12
            line.update(bound="fp32", algorithmic_flops=flops, achieved_tflops=tf, peak_tflops=peaks["fp32_tflops"], frac=tf / peaks["fp32_tflops"])
        print(json.dumps(line), flush=True)

    means = pl._means.clone()
    xs, eps = ops.sample(sh, tab, means, seed=1, draw=0, want_eps=True)
    # K2 with in-kernel RNG: writes M*w per sample
    out = torch.empty_like(xs)

    def k2_rng():
        ops._lib.check(ops._lib.load().sgpmp_sample(ops.C.byref(sh), ops._ptr(tab), ops._ptr(means), None, 1, 0, ops._ptr(out), None, ops._stream()), "k2")
    report("K2 sample (in-kernel Philox)", timeit(k2_rng), bytes_=ntraj * M * 4, note="writes M*4 B per trajectory sample")

    def k2_inj():
        ops._lib.check(ops._lib.load().sgpmp_sample(ops.C.byref(sh), ops._ptr(tab), ops._ptr(means), ops._ptr(eps), 1, 0, ops._ptr(out), None, ops._stream()), "k2")
    report("K2 sample (injected eps)", timeit(k2_inj), bytes_=2 * ntraj * M * 4, note="reads eps + writes samples")

    # K2 as the reference formulates it (dense L @ eps), on the tensor cores: the comparison point for the banded recurrence
    if B * NP <= 65535 and T % 2 == 0:
        L1 = ops.prior_dense_L(tab, 1, torch.float32)
        bf16_tf0 = (json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", 1663.8) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1663.8)

        def k2_dense():
            ops._lib.check(ops._lib.load().sgpmp_sample_dense_tc(ops.C.byref(sh), ops._ptr(L1), ops._ptr(means), ops._ptr(eps), ops._ptr(out), ops._stream()), "k2d")
        ms = timeit(k2_dense)
        useful = ntraj * n * 2.0 * (2 * T) ** 2 / 2          # lower-triangular L: half of the dense product is structurally zero
        tiles_m = (2 * T + 127) // 128
        issued = 3.0 * B * NP * n * ((S + 127) // 128) * sum(2.0 * 128 * 128 * (min(2 * T, 128 * (m + 1)) + 31) // 32 * 32 for m in range(tiles_m))
        print(json.dumps({"kernel": "K2 sample, dense-L variant (tcgen05 kind::tf32, 3xTF32, injected eps)", "workload": w["name"], "problems": B,
                          "traj_samples": ntraj, "ms": ms, "issued_tflops": issued / (ms * 1e-3) / 1e12, "tf32_peak_tflops_est": bf16_tf0 / 2,
                          "frac_of_tf32_peak_issued": issued / (ms * 1e-3) / 1e12 / (bf16_tf0 / 2),
                          "algorithmic_bytes": 2 * ntraj * M * 4, "achieved_gbs": 2 * ntraj * M * 4 / (ms * 1e-3) / 1e9, "frac_of_hbm": 2 * ntraj * M * 4 / (ms * 1e-3) / 1e9 / hbm,
                          "note": "same inputs and outputs as 'K2 sample (injected eps)' above; %d flop per sample issued against %d for the banded recurrence" % (round(issued / ntraj), 16 * T * n)}), flush=True)

    costs = torch.empty(B, NP, S, device=dev)

    def k3():
        ops._lib.check(ops._lib.load().sgpmp_cost(ops.C.byref(sh), ops.C.byref(desc), ops._ptr(tab), ops._ptr(xs), ops._ptr(means), ops._ptr(costs), None, ops._stream()), "k3")
    f_cost, _ = bench.algorithmic_flops_per_traj(w)
    f_cost -= 16 * T * n + 2 * M + 10      # minus sampling recurrence and softmax/update
    ms3 = timeit(k3)
    report("K3 cost", ms3, flops=ntraj * f_cost, note="FP32-bound for Panda")
    report("K3 cost (as HBM stream)", ms3, bytes_=ntraj * (M + 1) * 4, note="reads M*4 B per trajectory sample")

    grad = torch.empty_like(means)
    wts = torch.empty_like(costs)
    mu2 = means.clone()

    def k4():
        ops._lib.check(ops._lib.load().sgpmp_update(ops.C.byref(sh), float(w["temperature"]), float(w["step_size"]), ops._ptr(costs), ops._ptr(xs), ops._ptr(mu2), ops._ptr(grad), ops._ptr(wts), ops._stream()), "k4")
    report("K4 update", timeit(k4), bytes_=ntraj * (M + 2) * 4, note="reads samples + costs, writes weights")

    def k_stats():
        ops.local_stats(sh, w["temperature"], costs, eps)
    report("local_stats (split mode)", timeit(k_stats), bytes_=ntraj * (M + 1) * 4)

    # weighted covariance diagnostic (tcgen05): the first Bc problems only (the [M, M] output per particle is 3.2 MB for Panda)
    Bc = min(B, 16)
    shc = ops.make_shape(Bc, G, K, S, T, n, torch.float32)
    xc, mc, wc = xs[:Bc].contiguous(), means[:Bc].contiguous(), torch.softmax(-costs[:Bc] / float(w["temperature"]) * 1e-3, dim=-1).contiguous()
    covb = torch.empty(Bc, NP, M, M, device=dev)
    bf16_tf = (json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", 1663.8) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1663.8)
    for tc, nm in ((1, "weighted covariance (tcgen05 kind::tf32, 3xTF32)"), (0, "weighted covariance (CUDA cores, fp64 accumulate)")):
        def kcov(tc=tc):
            ops._lib.check(ops._lib.load().sgpmp_weighted_cov(ops.C.byref(shc), ops._ptr(xc), ops._ptr(mc), ops._ptr(wc), ops._ptr(covb), tc, ops._stream()), "cov")
        ms = timeit(kcov)
        fl = 2.0 * Bc * NP * M * M * S
        tiles = (M + 127) // 128
        issued = 3 * 2.0 * Bc * NP * (tiles * (tiles + 1) // 2) * 128 * 128 * (-(-S // 32) * 32) if tc else fl   # upper-triangle tiles, 3xTF32
        by = Bc * NP * (M * S + M * M + S) * 4
        print(json.dumps({"kernel": nm, "workload": w["name"], "problems": Bc, "particles": Bc * NP, "M": M, "S": S, "ms": ms,
                          "useful_tflops": fl / (ms * 1e-3) / 1e12, "issued_tflops": issued / (ms * 1e-3) / 1e12,
                          "tf32_peak_tflops_est": bf16_tf / 2, "frac_of_tf32_peak_issued": (issued / (ms * 1e-3) / 1e12) / (bf16_tf / 2) if tc else None,
                          "algorithmic_bytes": by, "achieved_gbs": by / (ms * 1e-3) / 1e9, "frac_of_hbm": by / (ms * 1e-3) / 1e9 / hbm,
                          "note": "diagnostic without reference counterpart; tf32 peak estimated as half the measured bf16 GEMM peak"}), flush=True)

    # K1: latency
    from stoch_gpmp_b200.planner import prior_blocks
    for TT in (64, 1024):
        D, O = prior_blocks(TT, w["dt"], w["sig"]["sigma_start_sample"], w["sig"]["sigma_gp_sample"], w["sig"]["sigma_goal_sample"])
        Dt = torch.tensor([D], dtype=torch.float64, device=dev)
        Ot = torch.tensor([O], dtype=torch.float64, device=dev)
        report("K1 prior factor T=%d" % TT, timeit(lambda: ops.prior_factor(Dt, Ot)), note="latency-bound, one thread per prior (includes two small torch allocations)")


if __name__ == "__main__":
    main()
