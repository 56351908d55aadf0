#!/bin/bash
export PYTHONPATH=$PWD
for v in d2 d3 d4 d6 d8; do
  SGPMP_LIB=scratch/variants/$v.so python scratch/time_cost.py planar 1024
  SGPMP_LIB=scratch/variants/$v.so python scratch/time_cost.py planar 4096
done 2>&1 | grep K3 | tee gpurun_out/r2i_k3_lean_ring_depth.txt
timeout 200 ncu --set full --clock-control none -k regex:cost_kernel -c 1 -o gpurun_out/r2i_k3_planar_lean -f \
    python scratch/time_cost.py planar 1024 > gpurun_out/r2i_ncu_k3.log 2>&1
ls -la gpurun_out/*.ncu-rep
