for cfg in 4,4 4,2 4,1 8,4 8,2 2,1 6,2; do python scratch/quick_time.py panda 4096 SGPMP_SPLIT_CFG=$cfg; done
for cfg in 4,2 4,1 2,1; do python scratch/quick_time.py panda 4096 SGPMP_SPLIT_CFG=$cfg SGPMP_LIB=scratch/variants/ts3.so; done
python -m pytest tests -m gpu -q 2>&1 | tail -5
