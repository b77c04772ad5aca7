"""Freeze the instrumented-oracle operation count in profiles/roofline.json.  TEST INFRASTRUCTURE.
usage: python oracle/flopcount/count.py [--write]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import count_flops, task_config

cfg = task_config("flat")
# (1) binding figure: the double-support standing state -- zero actions (targets = default pose), after the drop from the
#     reset height has settled (2 s), no observation noise influence on the physics
stand = count_flops(cfg, 16, steps=50, settle=100)
# (2) the bench workload: N(0,1) actions from reset, the first 50 steps (falls and resets included)
rng = np.random.default_rng(0)
acts = [rng.normal(size=(64, 12)).astype(np.float32) for _ in range(60)]
rand = count_flops(cfg, 64, steps=50, actions=lambda s: acts[s % 60])
print("standing:", json.dumps(stand)); print("random  :", json.dumps(rand))
if "--write" in sys.argv:
    path = os.path.join(ROOT, "profiles", "roofline.json")
    d = json.load(open(path))
    # bench.py's roofline uses the count of the workload it times (random actions: most envs fall, few contact rows); the
    # standing figure SURVEY 8(d) names is 4.4x larger (8 sole corners x 4 pyramid edges in the dense Hessian) and would
    # overstate what the kernel does on the bench workload, so it is kept beside it, not used as the numerator
    d["fp32_flops_per_env_step"] = round(rand["flops"])
    d["fp32_flops_per_env_step_standing"] = round(stand["flops"])
    d.pop("fp32_flops_per_env_step_random_actions", None)
    d["how"] = ("instrumented oracle (SURVEY 8(d)): oracle/h1v2_oracle.c compiled with `double` replaced by a counting scalar "
                "(oracle/flopcount/counted_double.h, outputs bit-identical to the plain build), add + mul + div + sqrt + transcendental "
                "per env-step (4 substeps of dense 18-dof MuJoCo-semantics dynamics + managers), double-support standing state (zero actions, "
                "settled 2 s) = fp32_flops_per_env_step_standing; fp32_flops_per_env_step is the same count on the workload bench.py times (N(0,1) actions from reset, "
                "50 steps, falls and resets included) and is the numerator of bench.py's FP32 roofline. "
                "oracle/flopcount/count.py --write regenerates both.")
    d["instrumented_oracle_counts"] = {"standing": stand, "random_actions": rand}
    json.dump(d, open(path, "w"), indent=1)
