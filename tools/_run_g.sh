python bench.py --impl reference > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err
python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -c 600 gpurun_out/r2l_bench.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
