#!/bin/bash
# round-2 session-2 batch 1: parity, C5 sweep after the K1 prefetch, ncu of K3-planar and of the fp64 tiled sampler, example loop
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench_c5.py > gpurun_out/r2c_c5.jsonl 2> gpurun_out/r2c_c5.err
python examples/panda_environment.py 2>&1 | tail -2 > gpurun_out/r2c_example_panda.txt
python examples/planar_environment.py 2>&1 | tail -2 > gpurun_out/r2c_example_planar.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:cost_kernel -c 1 -o gpurun_out/r2c_k3_planar -f \
    python bench_kernels.py --workload planar --reps 1 > gpurun_out/r2c_ncu_k3.log 2>&1
cat gpurun_out/r2c_example_panda.txt gpurun_out/r2c_example_planar.txt
ls -la gpurun_out
