#!/bin/bash
# One GPU-box pass that refreshes every measured artefact of a round (run under gpurun from the repo root):
# parity suite, bench lines (Flat 4096 + 32768, Rsl 4096 + 32768, reference arm), full ncu captures at both env counts.
tag=${1:-rX}
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench_4096.json 2> gpurun_out/${tag}_bench_4096.err
python bench.py --envs 32768 --no-big > gpurun_out/${tag}_bench_32768.json 2> gpurun_out/${tag}_bench_32768.err
python bench.py --task rsl --no-cpu-baseline > gpurun_out/${tag}_bench_rsl_4096.json 2> gpurun_out/${tag}_bench_rsl_4096.err
python bench.py --task rsl --envs 32768 --no-big --no-cpu-baseline > gpurun_out/${tag}_bench_rsl_32768.json 2> gpurun_out/${tag}_bench_rsl_32768.err
python bench.py --task cat --no-cpu-baseline > gpurun_out/${tag}_bench_cat_4096.json 2> gpurun_out/${tag}_bench_cat_4096.err
for n in 4096 32768; do
  ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 30 --launch-count 1 -f -o gpurun_out/prof_${tag}_${n} python tools/prof_target.py $n 40 > gpurun_out/${tag}_ncu_${n}.log 2>&1
done
cat gpurun_out/${tag}_tests.log; for f in 4096 32768 rsl_4096 rsl_32768; do python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value']/1e6,2), 'M  e2e', round(d['e2e']['value']/1e6,2) if d.get('e2e') else None, 'big', (d.get('at_32768_envs_per_gpu') or {}).get('value'), 'solver', d.get('solver'))
"; done
