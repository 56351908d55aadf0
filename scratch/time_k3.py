"""Tuning aid: time the standalone kernels with an OLDER build of the library (symbols added since are skipped).
usage: SGPMP_LIB=scratch/variants/x.so python scratch/time_k3.py"""
import ctypes
import os
import subprocess
import sys
sys.path.insert(0, '.')
import stoch_gpmp_b200._lib as L
lib = ctypes.CDLL(L.lib_path())
for k in list(L.SIGNATURES):
    if not hasattr(lib, k):
        del L.SIGNATURES[k]
sys.argv = ['bench_kernels.py', '--workload', sys.argv[1] if len(sys.argv) > 1 else 'panda']
import bench_kernels
bench_kernels.main()
