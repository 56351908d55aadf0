"""Fast kernel-variant builder: recompile ONLY the listed sources with extra -D flags and link with the base objects of
stoch_gpmp_b200/_C (build the base first).   python scratch/fastvar.py name:src1.cu+src2.cu:-DX=1,-DY=2 ...
Then on the GPU box:   python scratch/fastvar.py time [panda|planar] [B]"""
import glob, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE); sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "variants")
def build(specs):
    from stoch_gpmp_b200 import build as B
    import concurrent.futures as cf
    os.makedirs(OUT, exist_ok=True)
    nvcc = B._nvcc()
    base = {os.path.basename(s)[:-3]: os.path.join(B.OUT_DIR, os.path.basename(s)[:-3] + ".o") for s in B._sources()}
    def one(spec):
        name, srcs, flags = (spec.split(":") + ["", ""])[:3]
        flags = [f for f in flags.split(",") if f]
        objs = dict(base)
        for src in srcs.split("+"):
            stem = src[:-3]
            obj = os.path.join(OUT, "%s_%s.o" % (name, stem))
            r = subprocess.run([nvcc] + B.NVCC_FLAGS + flags + ["-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
            if r.returncode: raise RuntimeError(r.stderr)
            objs[stem] = obj
        lib = os.path.join(OUT, name + ".so")
        r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + list(objs.values()) + ["-cudart", "static"], capture_output=True, text=True)
        if r.returncode: raise RuntimeError(r.stderr)
        return lib
    with cf.ThreadPoolExecutor(max_workers=6) as ex:
        for lib in ex.map(one, specs): print("built", lib)
if __name__ == "__main__":
    if sys.argv[1] == "time":
        wl = sys.argv[2] if len(sys.argv) > 2 else "panda"; Bn = sys.argv[3] if len(sys.argv) > 3 else "4096"
        for lib in sorted(glob.glob(os.path.join(OUT, "*.so"))):
            print(os.path.basename(lib), end="  ", flush=True)
            subprocess.run([sys.executable, os.path.join(HERE, "quick_time.py"), wl, Bn], env=dict(os.environ, SGPMP_LIB=lib))
    else:
        build(sys.argv[1:])
