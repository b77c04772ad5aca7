#!/bin/bash
# Evidence run: the reference's UNMODIFIED scripts/clean_rl/train.py (in-tree CleanRL PPO, utils/cleanrl/ppo.py) on the CaT id.
# Needs a temporary, git-ignored copy of the reference's python under baseline/_ref in which the CaT ids' entry point is switched
# from the class object CaTEnv to "h1v2_isaac_b200.env:H1v2CaTEnv" -- the one-line binding of INTEGRATION.md:
#   sed -i 's/entry_point=CaTEnv,/entry_point="h1v2_isaac_b200.env:H1v2CaTEnv",/' \
#       baseline/_ref/packages/biped_tasks/biped_tasks/tasks/locomotion/velocity/config/h12_12dof/__init__.py
REF=$PWD/baseline/_ref
export PYTHONPATH=$PWD/h1v2_isaac_b200/shims:$PWD:$REF/packages/biped_tasks:$REF/packages/biped_assets:$REF/scripts/clean_rl
mkdir -p gpurun_out/cat_train && cd gpurun_out/cat_train
md5sum $REF/scripts/clean_rl/train.py $REF/packages/biped_tasks/biped_tasks/utils/cleanrl/ppo.py
grep -n "H1v2CaTEnv" $REF/packages/biped_tasks/biped_tasks/tasks/locomotion/velocity/config/h12_12dof/__init__.py
python $REF/scripts/clean_rl/train.py --task Isaac-Velocity-CaT-Flat-H12_12dof-v0 --num_envs ${1:-4096} --num_iterations ${2:-100} --headless
echo "train.py exit code $?"
rm -rf logs/*/*/*/*.pt logs/*/*/*/model* 2>/dev/null
