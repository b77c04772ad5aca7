python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python bench.py > gpurun_out/r2_bench_4096.json 2> gpurun_out/r2_bench_4096.err
python bench.py --task rsl --no-cpu-baseline --no-ppo > gpurun_out/r2_bench_rsl_4096.json 2> /dev/null
python bench.py --task cat --no-cpu-baseline --no-ppo > gpurun_out/r2_bench_cat_4096.json 2> /dev/null
python tools/sweep.py > gpurun_out/r2_sweep.json 2> gpurun_out/r2_sweep.err
