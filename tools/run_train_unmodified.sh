#!/bin/bash
# Evidence run (BASELINE configs[2]): the reference's UNMODIFIED scripts/rsl_rl/train.py on the B200 backend.
# Needs a copy of the reference's python packages under baseline/_ref (git-ignored; made by hand for this run, removed after).
REF=$PWD/baseline/_ref
export PYTHONPATH=$PWD/h1v2_isaac_b200/shims:$PWD:$REF/packages/biped_tasks:$REF/packages/biped_assets:$REF/scripts/rsl_rl
mkdir -p gpurun_out/train_run && cd gpurun_out/train_run
md5sum $REF/scripts/rsl_rl/train.py
python $REF/scripts/rsl_rl/train.py --task ${3:-Isaac-Velocity-Flat-H12_12dof-v0} --num_envs ${1:-4096} --max_iterations ${2:-8} --headless
