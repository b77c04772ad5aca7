for n in 4096 32768; do
python tools/cat_target.py $n > gpurun_out/cat_plain_$n.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"step_kernel|cat_apply" --csv --log-file gpurun_out/r2cat_warm_$n.csv python tools/cat_target.py $n > /dev/null 2>&1
python - $n <<'PY'
import csv, sys, collections
n = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/r2cat_warm_{n}.csv")) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    agg.setdefault(r[ki].split("(")[0], []).append(v)
for k, v in agg.items():
    v = v[10:]
    print(f"n={n} {k}: {len(v)} launches, mean {sum(v)/len(v):.1f} us, min {min(v):.1f}")
PY
done
