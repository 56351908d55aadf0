#!/bin/bash
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in d1 d2 d3 d4 d6 d8; do
  SGPMP_LIB=scratch/variants/$v.so python scratch/time_cost.py planar 1024
  SGPMP_LIB=scratch/variants/$v.so python scratch/time_cost.py panda 1024
done 2>&1 | grep K3 | tee gpurun_out/r2d_k3_ring_depth.txt
python examples/panda_environment.py 2>&1 | tail -3
