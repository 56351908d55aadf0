#!/bin/bash
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scratch/time_plan.py SGPMP_PDL=0
python scratch/time_plan.py SGPMP_PDL=1
SGPMP_PDL=0 python examples/panda_environment.py 2>&1 | grep plan
SGPMP_PDL=1 python examples/panda_environment.py 2>&1 | grep plan
python scratch/time_cost.py planar 1024; python scratch/time_cost.py panda 1024
