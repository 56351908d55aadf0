#!/bin/bash
# final single-GPU pipeline of round 2 (session 2): parity, bench line, launch list, full capture of the dominant kernel, standalone kernels
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2z_gputest.txt
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2z_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2z_launch_bench.log 2>&1
python bench_kernels.py --workload panda > gpurun_out/r2z_kernels_panda.jsonl 2>/dev/null
python bench_kernels.py --workload planar > gpurun_out/r2z_kernels_planar.jsonl 2>/dev/null
python bench_c5.py > gpurun_out/r2z_c5.jsonl 2>/dev/null
python examples/panda_environment.py 2>&1 | tail -3 > gpurun_out/r2z_example_panda.txt
python examples/planar_environment.py 2>&1 | tail -3 > gpurun_out/r2z_example_planar.txt
cat gpurun_out/r2z_gputest.txt; tail -c 600 gpurun_out/r2z_bench_n1.json; ls -la gpurun_out
