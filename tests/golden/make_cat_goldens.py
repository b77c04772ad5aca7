"""Golden sequence for the Constraints-as-Terminations tail, produced by the REFERENCE's own code:
the ten constraint functions the CaT cfg uses (packages/biped_tasks/biped_tasks/utils/cat/constraints.py, parameters of
config/h12_12dof/cat_env_cfg.py:336-431) and the CaT probability class (utils/cat/constraint_manager.py:23-86), driven for T
steps on random states through a stand-in env.  Stateful parts are exercised: the Polyak running maxima of every constraint
column, the swing-height tracker of foot_clearance, the cross-env gather of no_move.  Also modify_constraint_p
(utils/cat/curriculums.py:20-42).  Run where /root/reference exists:  python tests/golden/make_cat_goldens.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "h1v2_isaac_b200", "shims"), os.path.join(REF, "packages", "biped_tasks"), os.path.join(REF, "packages", "biped_assets")]
import isaaclab  # noqa: E402,F401
from isaaclab.managers import SceneEntityCfg  # noqa: E402

from biped_tasks.utils.cat import constraints as C  # noqa: E402
from biped_tasks.utils.cat import curriculums as CUR  # noqa: E402
from biped_tasks.utils.cat.constraint_manager import CaT  # noqa: E402

rng = np.random.default_rng(20261019)
N, T, STEP_DT, DEADZONE = 128, 12, 0.02, 0.2
lo = np.array([-0.43, -3.14, -0.43, -0.26, -0.897334, -0.261799, -0.43, -3.14, -3.14, -0.26, -0.897334, -0.261799])
hi = np.array([0.43, 2.5, 3.14, 2.05, 0.523598, 0.261799, 0.43, 2.5, 0.43, 2.05, 0.523598, 0.261799])
mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * 0.9
soft = np.stack([mid - half, mid + half], -1).astype(np.float32)
vel_lim = np.full(12, 100.0, np.float32)
eff_lim = np.array([220, 220, 220, 360, 45, 45] * 2, np.float32)

# cat_env_cfg.py:336-431: (name, function, max_p, kwargs built below)
ORDER = ["contact", "joint_position_limits", "joint_velocity_limits", "joint_torque_limits", "foot_contact_force", "no_move", "base_orientation",
         "base_height", "foot_contact", "foot_clearance"]
MAX_P = {"contact": 1.0, **{k: 0.25 for k in ORDER[1:]}}
ill = SceneEntityCfg("contact_forces"); ill.body_ids = [2, 3, 4, 5]
feet = SceneEntityCfg("contact_forces"); feet.body_ids = [0, 1]
jall = SceneEntityCfg("robot"); jall.joint_ids = list(range(12))
rob = SceneEntityCfg("robot")
footpos = SceneEntityCfg("robot"); footpos.body_ids = [0, 1]

inp = {k: [] for k in ("joint_pos", "joint_vel", "torque", "force_hist", "gravity", "root_z", "cmd", "foot_z", "contact_time")}
out = {f"raw_{k}": [] for k in ORDER}
out.update({f"prob_{k}": [] for k in ORDER}); out.update({f"rmax_{k}": [] for k in ORDER}); out["cstr_prob"] = []
robot_data = types.SimpleNamespace(soft_joint_pos_limits=torch.from_numpy(soft)[None].repeat(N, 1, 1), joint_vel_limits=torch.from_numpy(vel_lim)[None].repeat(N, 1),
                                   joint_effort_limits=torch.from_numpy(eff_lim)[None].repeat(N, 1))
sensor_data = types.SimpleNamespace()


class _Sensor:
    data = sensor_data

    def compute_first_contact(self, dt, abs_tol=1.0e-8):  # isaaclab 2.1.0 ContactSensor.compute_first_contact
        return (self.data.current_contact_time > 0.0) * (self.data.current_contact_time < (dt + abs_tol))


env = types.SimpleNamespace(scene={"robot": types.SimpleNamespace(data=robot_data), "contact_forces": _Sensor()}, step_dt=STEP_DT, num_envs=N, device="cpu",
                            action_manager=types.SimpleNamespace(action_term_dim=[12]), command_manager=types.SimpleNamespace())
cat = CaT(tau=0.95, min_p=0.0)
cat._device = torch.device("cpu")
contact_time = np.zeros((N, 2), np.float32)
for t in range(T):
    qj = (mid + rng.uniform(-1.1, 1.1, (N, 12)) * half).astype(np.float32)
    qd = (rng.normal(size=(N, 12)) * rng.choice([2.0, 8.0, 60.0], (N, 1))).astype(np.float32)
    tau = (rng.normal(size=(N, 12)) * eff_lim * 0.6).astype(np.float32)
    Fh = (rng.normal(size=(N, 3, 6, 3)) * np.exp(rng.uniform(np.log(0.05), np.log(1500.0), (N, 1, 6, 1)))).astype(np.float32)
    Fh[(rng.random((N, 1, 6, 1)) < np.array([0.3, 0.3, 0.93, 0.93, 0.95, 0.95]).reshape(1, 1, 6, 1)).repeat(3, 1).repeat(3, 3)] = 0.0
    grav = rng.normal(size=(N, 3)) * np.array([0.08, 0.08, 0.0]) + np.array([0, 0, -1.0]); grav = (grav / np.linalg.norm(grav, axis=1, keepdims=True)).astype(np.float32)
    root_z = rng.normal(1.0, 0.04, N).astype(np.float32)
    cmd = rng.uniform(-1, 1, (N, 3)).astype(np.float32); cmd[rng.random(N) < (0.5 if t != 7 else 0.0)] *= 0.1  # step 7: no env inside the dead zone
    if t == 7:
        cmd[(np.abs(cmd) < DEADZONE).all(axis=1), 0] = 0.5
    foot_z = rng.uniform(0.03, 0.25, (N, 2)).astype(np.float32)
    in_contact = np.linalg.norm(Fh[:, -1, :2], axis=-1) > 1.0
    contact_time = np.where(in_contact, contact_time + STEP_DT, 0.0).astype(np.float32)
    robot_data.joint_pos, robot_data.joint_vel, robot_data.applied_torque = map(torch.from_numpy, (qj, qd, tau))
    robot_data.projected_gravity_b = torch.from_numpy(grav); robot_data.root_pos_w = torch.from_numpy(np.stack([0 * root_z, 0 * root_z, root_z], 1))
    robot_data.body_link_pos_w = torch.from_numpy(np.stack([0 * foot_z, 0 * foot_z, foot_z], -1))
    sensor_data.net_forces_w_history = torch.from_numpy(Fh); sensor_data.current_contact_time = torch.from_numpy(contact_time)
    env.command_manager.get_command = lambda name, c=cmd: torch.from_numpy(c)
    raw = {
        "contact": C.contact(env, ill), "joint_position_limits": C.joint_position_limits(env, jall),
        "joint_velocity_limits": C.joint_velocity_limits(env, jall), "joint_torque_limits": C.joint_torque_limits(env, jall),
        "foot_contact_force": C.foot_contact_force(env, 750.0, feet), "no_move": C.no_move(env, DEADZONE, 6.0, jall),
        "base_orientation": C.base_orientation(env, 0.1, rob), "base_height": C.base_height(env, 1.0, 0.05, rob),
        "foot_contact": C.foot_contact(env, feet), "foot_clearance": C.foot_clearance(env, 0.1, DEADZONE, footpos, feet)}
    for k in ORDER:  # ConstraintManager.compute (constraint_manager.py:213-229)
        cat.add(k, raw[k], MAX_P[k])
    for k, v in (("joint_pos", qj), ("joint_vel", qd), ("torque", tau), ("force_hist", Fh), ("gravity", grav), ("root_z", root_z), ("cmd", cmd), ("foot_z", foot_z),
                 ("contact_time", contact_time)):
        inp[k].append(v)
    for k in ORDER:
        r = raw[k].float(); r = r.unsqueeze(1) if r.ndim == 1 else r
        out[f"raw_{k}"].append(r.numpy().copy()); out[f"prob_{k}"].append(cat.probs[k].numpy().copy()); out[f"rmax_{k}"].append(cat.running_maxes[k].numpy().copy())
    out["cstr_prob"].append(cat.get_probs().numpy().copy())
ramp = []
for counter in (0, 1000, 60000, 120000, 500000):
    cm = types.SimpleNamespace(get_term_cfg=lambda n: types.SimpleNamespace(max_p=None), set_term_cfg=lambda n, c: None)
    ramp.append(CUR.modify_constraint_p(types.SimpleNamespace(common_step_counter=counter, constraint_manager=cm), None, "x", 24 * 5000, 0.25))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cat_sequence.npz"), soft_limits=soft, vel_limits=vel_lim, effort_limits=eff_lim,
                    order=np.array(ORDER), max_p=np.array([MAX_P[k] for k in ORDER], np.float32), ramp_counter=np.array([0, 1000, 60000, 120000, 500000]),
                    ramp_max_p=np.array(ramp, np.float64), **{k: np.stack(v) for k, v in inp.items()}, **{k: np.stack(v) for k, v in out.items()})
print("wrote cat_sequence.npz", {k: out[f'raw_{k}'][0].shape for k in ORDER}, "ramp", ramp)
