for cfg in "1 1" "0 0" "1 1"; do set -- $cfg
  echo "== GRAPH_ROLLOUT=$1 GRAPH_LEARNER=$2"
  H1V2_GRAPH_ROLLOUT=$1 H1V2_GRAPH_LEARNER=$2 bash tools/run_train_unmodified.sh 4096 400 > gpurun_out/r2_train_w_$1$2.log 2>&1
  grep -E "it (1|100|200|300|400)/400 " gpurun_out/r2_train_w_$1$2.log
done
