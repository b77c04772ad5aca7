python -m pytest tests/test_gpu_env.py -x -q -k "rsl_rl or graph_rollout or rsl_task" 2>&1 | tail -5
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -3 gpurun_out/r2n_bench.err
H1V2_GRAPH_LEARNER=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-big --no-e2e > gpurun_out/r2n_bench_eager_learner.json 2> /dev/null
