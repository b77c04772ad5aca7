set -x
tools/launch_list.sh r2 flat > gpurun_out/r2_launch_flat.txt 2>&1
tools/launch_list.sh r2cat cat > gpurun_out/r2_launch_cat.txt 2>&1
for n in 4096 32768; do
  python tools/prof_target.py $n 40 flat flush > gpurun_out/prof_plain_$n.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 30 --launch-count 1 -f -o gpurun_out/prof_r2a_$n python tools/prof_target.py $n 40 flat flush > gpurun_out/prof_ncu_$n.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
