python -m pytest tests/test_gpu_parity.py tests/test_gpu_fp64.py -x -q 2>&1 | tail -8
python tools/diag_e2e.py 4096,8192,32768
