#!/bin/bash
# BASELINE configs[3] (scaled to the GPUs at hand): env-sharded PPO, the reference's unmodified train.py under torchrun;
# rsl_rl's multi-GPU mode keys off WORLD_SIZE, the device is passed per rank because train.py has no --distributed flag.
REF=$PWD/baseline/_ref
export PYTHONPATH=$PWD/h1v2_isaac_b200/shims:$PWD:$REF/packages/biped_tasks:$REF/packages/biped_assets:$REF/scripts/rsl_rl
mkdir -p gpurun_out/train_run_ddp && cd gpurun_out/train_run_ddp
cat > launch.py <<'PY'
import os, sys, runpy
r = os.environ["LOCAL_RANK"]
sys.argv = [sys.argv[1]] + sys.argv[2:] + ["--device", f"cuda:{r}", f"agent.device=cuda:{r}"]
runpy.run_path(sys.argv[0], run_name="__main__")
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${1:-2} --master-addr 127.0.0.1 --master-port 29533 launch.py $REF/scripts/rsl_rl/train.py --task Isaac-Velocity-Flat-H12_12dof-v0 --num_envs ${2:-4096} --max_iterations ${3:-20} --headless
