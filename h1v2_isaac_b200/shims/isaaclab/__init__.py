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


# ---------------------------------------------------------------- terrains / ray caster (cfg classes the Rough id's flatten_cfg reads)
@configclass
class SubTerrainBaseCfg:
    """isaaclab.terrains.SubTerrainBaseCfg / HfTerrainBaseCfg fields [UPSTREAM 2.1.0]."""
    function = None
    proportion: float = 1.0
    size: tuple = (10.0, 10.0)
    flat_patch_sampling = None
    border_width: float = 0.0
    horizontal_scale: float = 0.1
    vertical_scale: float = 0.005
    slope_threshold = None


@configclass
class HfRandomUniformTerrainCfg(SubTerrainBaseCfg):
    """height_field/hf_terrains_cfg.py [UPSTREAM]; used by packages/biped_tasks/biped_tasks/utils/mdp/terrains.py:20-25."""
    noise_range: tuple = MISSING
    noise_step: float = MISSING
    downsampled_scale = None


@configclass
class TerrainGeneratorCfg:
    """isaaclab.terrains.TerrainGeneratorCfg [UPSTREAM]; instantiated by utils/mdp/terrains.py:11-28."""
    seed = None
    curriculum: bool = False
    size: tuple = MISSING
    border_width: float = 0.0
    border_height: float = 1.0
    num_rows: int = 1
    num_cols: int = 1
    color_scheme: str = "none"
    horizontal_scale: float = 0.1
    vertical_scale: float = 0.005
    slope_threshold = 0.75
    sub_terrains: dict = MISSING
    difficulty_range: tuple = (0.0, 1.0)
    use_cache: bool = False
    cache_dir: str = "/tmp/isaaclab/terrains"


@configclass
class TerrainImporterCfg:
    """isaaclab.terrains.TerrainImporterCfg [UPSTREAM]; instantiated by V/velocity_env_cfg.py:40-58."""
    class_type = None
    collision_group: int = -1
    prim_path: str = MISSING
    num_envs: int = 1
    terrain_type: str = "generator"
    terrain_generator = None
    usd_path = None
    env_spacing = None
    visual_material = None
    physics_material = None
    max_init_terrain_level = None
    debug_vis: bool = False


@configclass
class GridPatternCfg:
    """isaaclab.sensors.patterns.GridPatternCfg [UPSTREAM]: rays over torch.arange(-size/2, size/2 + 1e-9, resolution), "xy" ordering."""
    func = None
    resolution: float = MISSING
    size: tuple = MISSING
    direction: tuple = (0.0, 0.0, -1.0)
    ordering: str = "xy"


@configclass
class RayCasterCfg:
    """isaaclab.sensors.RayCasterCfg [UPSTREAM]; instantiated by V/velocity_env_cfg.py:61-68."""

    @configclass
    class OffsetCfg:
        pos: tuple = (0.0, 0.0, 0.0)
        rot: tuple = (1.0, 0.0, 0.0, 0.0)

    class_type = None
    prim_path: str = MISSING
    update_period: float = 0.0
    history_length: int = 0
    debug_vis: bool = False
    mesh_prim_paths: list = MISSING
    offset: OffsetCfg = OffsetCfg()
    attach_yaw_only: bool = MISSING
    pattern_cfg = MISSING
    max_distance: float = 1e6
    drift_range: tuple = (0.0, 0.0)
    visualizer_cfg = None


make_lenient("isaaclab.sensors", RayCasterCfg=RayCasterCfg)
make_lenient("isaaclab.sensors.patterns", GridPatternCfg=GridPatternCfg)
make_lenient("isaaclab.sensors.ray_caster", RayCasterCfg=RayCasterCfg)
make_lenient("isaaclab.terrains", TerrainImporterCfg=TerrainImporterCfg, TerrainGeneratorCfg=TerrainGeneratorCfg, HfRandomUniformTerrainCfg=HfRandomUniformTerrainCfg)
make_lenient("isaaclab.terrains.config")
make_lenient("isaaclab.terrains.terrain_generator_cfg", TerrainGeneratorCfg=TerrainGeneratorCfg)


class _RoughTerrainsModule(type(_me)):
    """isaaclab.terrains.config.rough: ROUGH_TERRAINS_CFG resolves, at first use, to the reference's OWN in-tree generator cfg of
    that name (packages/biped_tasks/biped_tasks/utils/mdp/terrains.py:11-28, one random-rough height field) -- upstream's cfg of the
    same name also holds mesh stairs / boxes / slopes, which the B200 backend does not simulate (flatten_cfg refuses them by name).
    Without biped_tasks, the same values come from h1v2_isaac_b200.tasks."""

    def __getattr__(self, name):
        if name != "ROUGH_TERRAINS_CFG":
            raise AttributeError(name)
        try:
            import importlib
            cfg = importlib.import_module("biped_tasks.utils.mdp.terrains").ROUGH_TERRAINS_CFG
        except Exception:  # noqa: BLE001
            from h1v2_isaac_b200.tasks import rough_terrains_cfg
            cfg = rough_terrains_cfg()
        setattr(self, name, cfg)
        return cfg


_rough = _RoughTerrainsModule("isaaclab.terrains.config.rough")
sys.modules[_rough.__name__] = _rough
sys.modules["isaaclab.terrains.config"].rough = _rough


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
