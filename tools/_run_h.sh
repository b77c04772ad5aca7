python -m pytest tests/test_gpu_cat.py tests/test_gpu_env.py -x -q 2>&1 | tail -12
python tools/quick_gpu_cat.py 2>&1 | tail -4
python tools/quick_gpu.py 2>&1 | grep "n="
