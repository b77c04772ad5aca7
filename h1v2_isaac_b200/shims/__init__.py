"""Import shims for machines without isaaclab / isaaclab_tasks / isaaclab_rl / gymnasium / rsl_rl.

Two ways to activate them for the reference's UNMODIFIED scripts/rsl_rl/train.py:
    PYTHONPATH=<repo>/h1v2_isaac_b200/shims:<repo>  python scripts/rsl_rl/train.py --task Isaac-Velocity-Flat-H12_12dof-v0 ...
or, from Python,  `import h1v2_isaac_b200.shims as s; s.install()`  before importing biped_tasks.
A real installation of any of these packages always wins: install() only fills the gaps."""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys

_DIR = os.path.dirname(os.path.abspath(__file__))
PACKAGES = ("gymnasium", "isaaclab", "isaaclab_tasks", "isaaclab_rl", "rsl_rl")


def install(force: bool = False) -> list[str]:
    """Make the shim packages importable as top-level modules for every package that is not really installed."""
    installed = []
    missing = [p for p in PACKAGES if force or (p not in sys.modules and importlib.util.find_spec(p) is None)]
    if missing and _DIR not in sys.path:
        sys.path.append(_DIR)  # appended: real packages found earlier on the path keep priority
    for p in missing:
        importlib.import_module(p)
        installed.append(p)
    return installed
