#!/bin/bash
# Evidence run: the reference's UNMODIFIED scripts/rsl_rl/play.py on the B200 backend, after tools/run_train_unmodified.sh
# (same cwd, so it finds the checkpoint): loads the last model, exports policy.pt / policy.onnx, rolls the policy out.
# --video makes the reference's loop stop after --video_length steps (RecordVideo itself is a no-op here: no renderer).
REF=$PWD/baseline/_ref
export PYTHONPATH=$PWD/h1v2_isaac_b200/shims:$PWD:$REF/packages/biped_tasks:$REF/packages/biped_assets:$REF/scripts/rsl_rl
cd gpurun_out/train_run || exit 1
md5sum $REF/scripts/rsl_rl/play.py
python $REF/scripts/rsl_rl/play.py --task ${2:-Isaac-Velocity-Flat-H12_12dof-v0} --num_envs ${1:-256} --headless --video --video_length ${3:-300}
echo "play.py exit code $?"
find logs -name "policy.*" | head
