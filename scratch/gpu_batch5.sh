#!/bin/bash
export PYTHONPATH=$PWD
for B in 4096 1024 512 256 128; do
  python scratch/quick_time.py panda $B SGPMP_SPLIT_CFG=4,4
  python scratch/quick_time.py panda $B SGPMP_SPLIT_CFG=8,8
done 2>&1 | grep "ms/iter" | tee gpurun_out/r2f_cfg88.txt
