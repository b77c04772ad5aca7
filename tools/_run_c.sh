python tools/diag_fp64.py 8192 96 1.0 D
SEED=5 python tools/diag_fp64.py 8192 96 0.3 D
SEED=7 python tools/diag_fp64.py 8192 96 1.0 D
DECIM=4 python tools/diag_fp64.py 8192 24 1.0 D
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
