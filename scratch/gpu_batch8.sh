#!/bin/bash
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scratch/time_cost.py planar 1024
python scratch/time_cost.py planar 4096
python scratch/time_plan.py
for S in 256 64; do
  python scratch/time_stats.py $S SGPMP_SPLIT_CFG=8,8
  python scratch/time_stats.py $S SGPMP_SPLIT_CFG=4,4
  python scratch/time_stats.py $S SGPMP_SPLIT_CFG=2,2
done
