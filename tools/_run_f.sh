python -m pytest tests/test_gpu_env.py -x -q -k "rsl_rl or graph_rollout" 2>&1 | tail -15
