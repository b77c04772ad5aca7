"""The instrumented oracle behind the FP32 roofline: oracle/h1v2_oracle.c compiled as C++ with `double` replaced by a counting
scalar (oracle/flopcount/counted_double.h).  It must compute exactly what the plain build computes, and the operation count frozen in
profiles/roofline.json must be what the committed recipe (oracle/flopcount/count.py) produces."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_counted_build_is_bit_identical_and_reproduces_the_frozen_count():
    import oracle.oracle as oo
    cfg = oo.task_config("flat")
    rng = np.random.default_rng(0)
    acts = [rng.normal(size=(64, 12)).astype(np.float32) for _ in range(60)]
    # same outputs as the plain build, step by step
    plain = oo.Oracle(cfg, 64, seed=3, threads=1)
    plain.observe()
    want = [plain.step(acts[s]) for s in range(6)]
    saved = (oo._lib, oo._SO)
    oo.build()
    oo._lib, oo._SO = None, oo._SO_COUNTED
    try:
        counted = oo.Oracle(cfg, 64, seed=3, threads=1)
        counted.observe()
        for s in range(6):
            got = counted.step(acts[s])
            assert all(np.array_equal(a, b) for a, b in zip(got, want[s])), s
        del counted
    finally:
        oo._lib, oo._SO = saved
    # the frozen figure: N(0,1) actions from reset, 64 envs x 50 steps (oracle/flopcount/count.py)
    c = oo.count_flops(cfg, 64, steps=50, actions=lambda s: acts[s % 60])
    frozen = json.load(open(os.path.join(ROOT, "profiles", "roofline.json")))
    assert abs(c["flops"] - frozen["fp32_flops_per_env_step"]) <= 0.001 * frozen["fp32_flops_per_env_step"]
    assert c["fma"] == 0 and c["mul"] > c["add"] > 10 * c["div"] > 0  # -ffp-contract=off on both builds; a dynamics step is multiply-add work
    assert frozen["fp32_flops_per_env_step_standing"] > 3 * frozen["fp32_flops_per_env_step"]  # double support: 32 contact rows in the dense Hessian
