"""rsl_rl stand-in (rsl-rl-lib 2.3.3 surface used by the reference: runners.OnPolicyRunner).  The PPO actor-critic MLP
stays in PyTorch (north star); multi-GPU runs all-reduce gradients over NCCL exactly where rsl_rl 2.3.3 does."""
__version__ = "2.3.3+h1v2_b200_shim"
from . import runners  # noqa: F401
