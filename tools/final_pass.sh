#!/bin/bash
# Final pass of a session (run under gpurun from the repo root): parity suite, bench lines of every id, reference arm, ncu launch list of the bench command
tag=r4
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench_4096.json 2> gpurun_out/${tag}_bench_4096.err
python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
python bench.py --task rsl --no-cpu-baseline > gpurun_out/${tag}_bench_rsl_4096.json 2> gpurun_out/${tag}_bench_rsl_4096.err
python bench.py --task cat --no-cpu-baseline > gpurun_out/${tag}_bench_cat_4096.json 2> gpurun_out/${tag}_bench_cat_4096.err
python bench.py --task rough --no-cpu-baseline > gpurun_out/${tag}_bench_rough.json 2> gpurun_out/${tag}_bench_rough.err
bash tools/launch_list.sh ${tag} > gpurun_out/${tag}_launch_list.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
cat gpurun_out/${tag}_tests.log gpurun_out/${tag}_smoke.log
for f in 4096 rsl_4096 cat_4096 rough reference_arm; do python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value']/1e6,3), 'M  e2e', round(d['e2e']['value']/1e6,2) if d.get('e2e') else None, 'big', (d.get('at_32768_envs_per_gpu') or {}).get('value'), 'ppo', (d.get('ppo') or {}).get('value'), 'roofline', (d.get('roofline') or {}).get('frac'))
"; done
