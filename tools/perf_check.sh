#!/bin/bash
# GPU-box check used after every kernel change: parity suite, throughput at 4096/32768 envs, a light ncu metrics pass.
# usage (under gpurun): bash tools/perf_check.sh <tag> [full]
tag=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_tests.log
python tools/quick_gpu.py > gpurun_out/${tag}_quick.log 2>&1
M=sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,smsp__sass_inst_executed_op_shared_ld.sum,smsp__sass_inst_executed_op_shared_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
if [ "$2" == "full" ]; then
  ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 30 --launch-count 1 -f -o gpurun_out/prof_${tag} python tools/prof_target.py 32768 40 > gpurun_out/${tag}_ncu.log 2>&1
else
  ncu --metrics $M --clock-control none -k regex:step_kernel --launch-skip 30 --launch-count 1 python tools/prof_target.py 32768 40 > gpurun_out/${tag}_ncu.log 2>&1
fi
cat gpurun_out/${tag}_tests.log gpurun_out/${tag}_quick.log; grep -E "^\s+(sm__|gcc__|smsp__|gpu__|launch__|l1tex__)" gpurun_out/${tag}_ncu.log
