#!/bin/bash
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for B in 4096 1024 256; do
  for W in 4 8; do python scratch/quick_time.py planar $B SGPMP_STATE_WARPS=$W; done
done 2>&1 | grep "ms/iter" | tee gpurun_out/r2g_planar_warps.txt
python scratch/quick_time.py panda 512
python scratch/quick_time.py panda 64
python scratch/quick_time.py panda 64 SGPMP_SPLIT_CFG=4,4
