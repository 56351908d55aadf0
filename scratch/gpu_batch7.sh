#!/bin/bash
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python scratch/time_cost.py panda 1024
python scratch/time_plan.py
python bench_kernels.py --workload panda 2>/dev/null | grep '"K3 cost"' | cut -c1-400
