"""Stand-in for the parts of isaaclab 2.1.0 that the reference imports (SURVEY.md section 8(b) shim list).

Only importable through `h1v2_isaac_b200.shims.install()` or by putting `h1v2_isaac_b200/shims` on PYTHONPATH, and only
meant for machines where the real isaaclab is absent.  `isaaclab.envs.ManagerBasedRLEnv` here IS the B200 backend env, so
the reference's gym registration (config/h12_12dof/__init__.py:40-49, entry_point "isaaclab.envs:ManagerBasedRLEnv")
resolves to it unchanged."""
from __future__ import annotations

import argparse
import os
import pickle
import sys

from h1v2_isaac_b200.shims._configclass import MISSING, configclass
from h1v2_isaac_b200.shims._lenient import Placeholder, install_finder, make_lenient

install_finder()
__version__ = "2.1.0+h1v2_b200_shim"
_me = sys.modules[__name__]


# ---------------------------------------------------------------- utils
def print_dict(val, nesting: int = -4, start: bool = True):
    if isinstance(val, dict):
        if not start:
            print("")
        nesting += 4
        for k in val:
            print(nesting * " ", end="")
            print(k, end=": ")
            print_dict(val[k], nesting, start=False)
    else:
        print(val)


def _plain(o):
    if hasattr(o, "to_dict") and not isinstance(o, type):
        o = o.to_dict()
    if isinstance(o, dict):
        return {str(k): _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    if isinstance(o, (int, float, str, bool)) or o is None:
        return o
    return repr(o)


def dump_yaml(filename: str, data, sort_keys: bool = False):
    import yaml
    if not filename.endswith("yaml"):
        filename += ".yaml"
    os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
    with open(filename, "w") as f:
        yaml.safe_dump(_plain(data), f, default_flow_style=False, sort_keys=sort_keys)


def dump_pickle(filename: str, data):
    if not filename.endswith("pkl"):
        filename += ".pkl"
    os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
    with open(filename, "wb") as f:
        try:
            pickle.dump(data, f)
        except Exception:
            pickle.dump(_plain(data), f)


utils = make_lenient("isaaclab.utils", configclass=configclass)
make_lenient("isaaclab.utils.dict", print_dict=print_dict)
make_lenient("isaaclab.utils.io", dump_yaml=dump_yaml, dump_pickle=dump_pickle)
make_lenient("isaaclab.utils.configclass", configclass=configclass)
utils.configclass = configclass  # upstream re-exports the decorator over the submodule of the same name


@configclass
class NoiseCfg:
    func = None
    operation: str = "add"


@configclass
class AdditiveUniformNoiseCfg(NoiseCfg):
    n_min: float = -1.0
    n_max: float = 1.0


@configclass
class NoiseModelCfg:
    class_type = None
    noise_cfg = MISSING


make_lenient("isaaclab.utils.noise", NoiseCfg=NoiseCfg, AdditiveUniformNoiseCfg=AdditiveUniformNoiseCfg, UniformNoiseCfg=AdditiveUniformNoiseCfg,
             NoiseModelCfg=NoiseModelCfg)
make_lenient("isaaclab.utils.math")
make_lenient("isaaclab.utils.modifiers")


# ---------------------------------------------------------------- managers (cfg classes only; the managers themselves are fused into the kernel)
class SceneEntityCfg:
    """isaaclab.managers.SceneEntityCfg: name is positional in the reference (SceneEntityCfg("robot", joint_names=[...]))."""

    def __init__(self, name=MISSING, joint_names=None, joint_ids=slice(None), fixed_tendon_names=None, fixed_tendon_ids=slice(None),
                 body_names=None, body_ids=slice(None), object_collection_names=None, object_collection_ids=slice(None),
                 preserve_order=False):
        self.name, self.joint_names, self.joint_ids = name, joint_names, joint_ids
        self.fixed_tendon_names, self.fixed_tendon_ids = fixed_tendon_names, fixed_tendon_ids
        self.body_names, self.body_ids = body_names, body_ids
        self.object_collection_names, self.object_collection_ids = object_collection_names, object_collection_ids
        self.preserve_order = preserve_order

    def to_dict(self):
        return {k: (v if not isinstance(v, slice) else "slice(None)") for k, v in vars(self).items()}

    def __repr__(self):
        return f"SceneEntityCfg({self.to_dict()})"


@configclass
class ManagerTermBaseCfg:
    func = MISSING
    params: dict = {}


@configclass
class RewardTermCfg(ManagerTermBaseCfg):
    weight: float = MISSING


@configclass
class TerminationTermCfg(ManagerTermBaseCfg):
    time_out: bool = False


@configclass
class CurriculumTermCfg(ManagerTermBaseCfg):
    pass


@configclass
class EventTermCfg(ManagerTermBaseCfg):
    mode: str = MISSING
    interval_range_s = None
    is_global_time: bool = False
    min_step_count_between_reset: int = 0


@configclass
class ObservationTermCfg(ManagerTermBaseCfg):
    modifiers = None
    noise = None
    clip = None
    scale = None
    history_length: int = 0
    flatten_history_dim: bool = True


@configclass
class ObservationGroupCfg:
    concatenate_terms: bool = True
    enable_corruption: bool = False
    history_length = None
    flatten_history_dim: bool = True


@configclass
class ActionTermCfg:
    class_type = None
    asset_name: str = MISSING
    debug_vis: bool = False
    clip = None


@configclass
class CommandTermCfg:
    class_type = None
    resampling_time_range = MISSING
    debug_vis: bool = False


@configclass
class RecorderTermCfg:
    class_type = None


managers = make_lenient("isaaclab.managers", SceneEntityCfg=SceneEntityCfg, ManagerTermBaseCfg=ManagerTermBaseCfg, RewardTermCfg=RewardTermCfg,
                        TerminationTermCfg=TerminationTermCfg, CurriculumTermCfg=CurriculumTermCfg, EventTermCfg=EventTermCfg,
                        ObservationTermCfg=ObservationTermCfg, ObservationGroupCfg=ObservationGroupCfg, ActionTermCfg=ActionTermCfg,
                        CommandTermCfg=CommandTermCfg, RecorderTermCfg=RecorderTermCfg)
make_lenient("isaaclab.managers.manager_term_cfg", ManagerTermBaseCfg=ManagerTermBaseCfg, RewardTermCfg=RewardTermCfg,
             TerminationTermCfg=TerminationTermCfg, CurriculumTermCfg=CurriculumTermCfg, EventTermCfg=EventTermCfg,
             ObservationTermCfg=ObservationTermCfg, ObservationGroupCfg=ObservationGroupCfg, ActionTermCfg=ActionTermCfg,
             CommandTermCfg=CommandTermCfg)
make_lenient("isaaclab.managers.manager_base")


# ---------------------------------------------------------------- assets / actuators / scene / sim
@configclass
class AssetBaseCfg:
    @configclass
    class InitialStateCfg:
        pos: tuple = (0.0, 0.0, 0.0)
        rot: tuple = (1.0, 0.0, 0.0, 0.0)

    class_type = None
    prim_path: str = MISSING
    spawn = None
    init_state = InitialStateCfg()
    collision_group: int = 0
    debug_vis: bool = False


@configclass
class ArticulationCfg(AssetBaseCfg):
    @configclass
    class InitialStateCfg(AssetBaseCfg.InitialStateCfg):
        lin_vel: tuple = (0.0, 0.0, 0.0)
        ang_vel: tuple = (0.0, 0.0, 0.0)
        joint_pos: dict = {".*": 0.0}
        joint_vel: dict = {".*": 0.0}

    init_state = InitialStateCfg()
    soft_joint_pos_limit_factor: float = 1.0
    actuators: dict = MISSING


make_lenient("isaaclab.assets", AssetBaseCfg=AssetBaseCfg, ArticulationCfg=ArticulationCfg)
make_lenient("isaaclab.assets.articulation", ArticulationCfg=ArticulationCfg)


@configclass
class ActuatorBaseCfg:
    class_type = None
    joint_names_expr = MISSING
    effort_limit = None
    velocity_limit = None
    effort_limit_sim = None
    velocity_limit_sim = None
    stiffness = MISSING
    damping = MISSING
    armature = None
    friction = None


@configclass
class IdealPDActuatorCfg(ActuatorBaseCfg):
    pass


@configclass
class ImplicitActuatorCfg(ActuatorBaseCfg):
    pass


@configclass
class DCMotorCfg(IdealPDActuatorCfg):
    saturation_effort: float = MISSING


@configclass
class DelayedPDActuatorCfg(IdealPDActuatorCfg):
    min_delay: int = 0
    max_delay: int = 0


make_lenient("isaaclab.actuators", ActuatorBaseCfg=ActuatorBaseCfg, IdealPDActuatorCfg=IdealPDActuatorCfg, ImplicitActuatorCfg=ImplicitActuatorCfg,
             DCMotorCfg=DCMotorCfg, DelayedPDActuatorCfg=DelayedPDActuatorCfg)


@configclass
class InteractiveSceneCfg:
    num_envs: int = MISSING
    env_spacing: float = MISSING
    lazy_sensor_update: bool = True
    replicate_physics: bool = True
    filter_collisions: bool = True


make_lenient("isaaclab.scene", InteractiveSceneCfg=InteractiveSceneCfg)


@configclass
class SimulationCfg:
    physics_prim_path: str = "/physicsScene"
    device: str = "cuda:0"
    dt: float = 1.0 / 60.0
    render_interval: int = 1
    gravity: tuple = (0.0, 0.0, -9.81)
    enable_scene_query_support: bool = False
    use_fabric: bool = True
    disable_contact_processing: bool = False
    physx = Placeholder()
    physics_material = Placeholder(static_friction=0.5, dynamic_friction=0.5, restitution=0.0)
    render = Placeholder()


make_lenient("isaaclab.sim", SimulationCfg=SimulationCfg)
make_lenient("isaaclab.sensors")
make_lenient("isaaclab.sensors.patterns")
make_lenient("isaaclab.terrains")
make_lenient("isaaclab.terrains.config")
make_lenient("isaaclab.terrains.config.rough")
make_lenient("isaaclab.terrains.terrain_generator_cfg")


# ---------------------------------------------------------------- app
class _App:
    def close(self):
        pass

    def is_running(self):
        return True


class AppLauncher:
    """No Omniverse Kit to boot: keeps the CLI surface of isaaclab.app.AppLauncher (scripts/rsl_rl/train.py:30,42-43)."""

    def __init__(self, launcher_args=None, **kwargs):
        self.app = _App()

    @staticmethod
    def add_app_launcher_args(parser: argparse.ArgumentParser) -> None:
        g = parser.add_argument_group("app_launcher arguments", description="Arguments for the AppLauncher (shim).")
        g.add_argument("--headless", action="store_true", default=True)
        g.add_argument("--livestream", type=int, default=-1)
        g.add_argument("--enable_cameras", action="store_true", default=False)
        g.add_argument("--device", type=str, default=None, help="device the simulation runs on, e.g. cuda:0")
        g.add_argument("--verbose", action="store_true")
        g.add_argument("--experience", type=str, default="")
        g.add_argument("--kit_args", type=str, default="")


make_lenient("isaaclab.app", AppLauncher=AppLauncher)

from . import envs  # noqa: E402,F401  (defines isaaclab.envs, isaaclab.envs.mdp, ...)
