"""Build the in-tree CUDA library (sm_100a only) — `python -m stoch_gpmp_b200.build`.

nvcc cross-compiles without a GPU.  Each .cu is compiled to an object in parallel, then linked into
stoch_gpmp_b200/_C/libsgpmp.so (git-ignored, travels to the GPU box with the tree).
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libsgpmp.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--threads", "2", "-I", INCLUDE]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found (set $NVCC)")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "stoch_gpmp_b200.h")]
    return sorted(files)


def source_hash():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS[:6]).encode())
    for f in _deps():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_fresh():
    stamp = LIB + ".hash"
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == source_hash()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libsgpmp.so.  Returns the library path."""
    if not force and is_fresh():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(LIB + ".hash", "w") as fh:
        fh.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
