#!/bin/bash
# compute-sanitizer evidence (SURVEY section 5 "race detection"): memcheck, racecheck, initcheck, synccheck over every
# kernel path at a small env count.  Output: gpurun_out/r2_sanitizer_<tool>.txt (copy the summaries to profiles/).
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py 96 > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done|Error|hazard" gpurun_out/r2_sanitizer_$tool.txt | head -8
done
