"""Kernel-tuning aid: build libsgpmp variants with extra -D flags into scratch/variants/<name>.so (run here, no GPU),
then time the fused loop with each of them on the GPU box:

    python scratch/variants.py build base: unroll5:-DSGPMP_SPH_UNROLL=5
    gpurun -- python scratch/variants.py time [panda|planar] [B]
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "variants")


def build(specs):
    from stoch_gpmp_b200 import build as B
    import concurrent.futures as cf
    os.makedirs(OUT, exist_ok=True)
    nvcc = B._nvcc()

    def one(spec):
        name, _, flags = spec.partition(":")
        flags = [f for f in flags.split(",") if f]
        d = os.path.join(OUT, name + "_obj")
        os.makedirs(d, exist_ok=True)
        objs = []
        for src in B._sources():
            obj = os.path.join(d, os.path.basename(src)[:-3] + ".o")
            r = subprocess.run([nvcc] + B.NVCC_FLAGS + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
            if r.returncode:
                raise RuntimeError(r.stderr)
            objs.append(obj)
        lib = os.path.join(OUT, name + ".so")
        r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-cudart", "static"],
                           capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr)
        return lib
    with cf.ThreadPoolExecutor(max_workers=4) as ex:
        for lib in ex.map(one, specs):
            print("built", lib)


def time_one(workload, B):
    import torch
    import bench
    dev = torch.device('cuda:0')
    w = bench.workload(workload, B)
    pl = bench.build_planner(w, B, dev)
    obs = {'obstacle_spheres': torch.tensor(w['spheres'], dtype=torch.float32, device=dev)} if workload == 'panda' else {}
    pl.optimize(opt_iters=2, return_samples=False, **obs)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pl.optimize(opt_iters=5, return_samples=False, **obs)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    print("%-24s %-7s B=%d  %.3f ms/iter" % (os.path.basename(os.environ.get("SGPMP_LIB", "default")), workload, B, best), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "time":
        workload = sys.argv[2] if len(sys.argv) > 2 else "panda"
        B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
        for lib in sorted(glob.glob(os.path.join(OUT, "*.so"))):
            env = dict(os.environ, SGPMP_LIB=lib)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", workload, str(B)], env=env)
    elif sys.argv[1] == "one":
        time_one(sys.argv[2], int(sys.argv[3]))
