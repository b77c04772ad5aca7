#!/bin/bash
# ncu launch list of a short bench run (B200_PROFILING.md: --metrics gpu__time_duration.sum --clock-control none), after the same
# command has exited 0 without ncu.  Writes gpurun_out/<tag>_launches.csv and a per-kernel summary.
tag=${1:-rX}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-big --no-ppo ${2:+--task $2}"   # optional 2nd argument: flat / rsl / cat
$CMD > gpurun_out/${tag}_launch_plain.log 2>&1 || { echo "bench failed without ncu"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launch_ncu.log 2>&1
python - "$tag" "$CMD" <<'PY'
import csv, collections, sys
tag, cmd = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in agg.values())
with open(f"gpurun_out/{tag}_launches_summary.txt", "w") as f:
    f.write(f"# ncu launch list of `{cmd}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n")
    f.write("# kernel | launches | total us | share   (fma_peak_kernel = the FP32-peak micro-benchmark, outside every timed region; one step = ONE step_kernel<1,0,0,Q> launch, Q = 1: the mirror-lane instantiation of 8 envs per warp; CaT: step_kernel<1,1,0,Q> + cat_apply_kernel)\n")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k} | {n} | {v:.1f} | {100 * v / tot:.1f}%\n")
print(open(f"gpurun_out/{tag}_launches_summary.txt").read())
PY
