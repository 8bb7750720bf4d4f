python tools/gpu_check.py 40000 warm 2>&1 | tail -20
python tools/gpu_check.py 40000 cold 2>&1 | tail -20
