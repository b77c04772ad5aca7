"""isaaclab_rl stand-in: the rsl_rl cfg classes and RslRlVecEnvWrapper (isaaclab_rl 2.1.0 semantics)."""
from __future__ import annotations

import sys

from h1v2_isaac_b200.shims._lenient import install_finder, make_lenient

install_finder()
from . import rsl_rl  # noqa: E402,F401
