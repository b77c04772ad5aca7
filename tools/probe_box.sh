#!/bin/bash
# What the GPU box has (VERDICT r1 item 1b / SURVEY App. E.8): python packages of the reference's CPU path, host topology.
echo "== python imports"
for m in mujoco gymnasium rsl_rl isaaclab isaacsim onnx onnxruntime onnxscript warp tensorboard hydra omegaconf; do
  python - <<PY 2>&1 | tail -1
try:
    import $m
    print("$m: OK", getattr($m, "__version__", "?"))
except Exception as e:
    print("$m: absent (%s: %s)" % (type(e).__name__, e))
PY
done
echo "== host"
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|L3|L2" ; (numactl -H 2>/dev/null || echo "numactl absent"); cat /sys/devices/system/node/online 2>/dev/null
free -g | head -2
echo "== gpu"
nvidia-smi --query-gpu=name,pci.bus_id,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv
nvidia-smi topo -m 2>/dev/null | head -20
