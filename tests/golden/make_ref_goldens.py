"""Golden vectors produced by the REFERENCE's own Python, imported from /root/reference in the build container
(the reference cannot travel to the GPU box; these small fixtures can).  Run:  python tests/golden/make_ref_goldens.py

  feet_air_time.npz     packages/biped_tasks/.../velocity/mdp/rewards.py:13-35 (feet_air_time) and :38-62
                        (feet_air_time_positive_biped), evaluated on random contact-timer states through a stand-in env
                        that exposes exactly the attributes those functions read.
  obs_history.npz       packages/biped_tasks/biped_tasks/utils/history/circular_buffer.py:22-137 (CircularBuffer: append,
                        first-push back-fill, reset, buffer()) driven term by term and flattened/concatenated as
                        utils/history/observation_manager.py:335-355 does, on a random sample sequence with resets;
                        projected gravity of each sample from packages/biped_deploy/biped_deploy/controllers/rl.py:86-95.
  deploy_obs.npz        packages/biped_deploy/biped_deploy/controllers/rl.py:34-121 ObservationHandler.get_observations on the
                        same kind of sequence (single env): the deployment side's view of the same 450-vector layout.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "h1v2_isaac_b200", "shims"), os.path.join(REF, "packages", "biped_tasks"),
                os.path.join(REF, "packages", "biped_assets"), os.path.join(REF, "packages", "biped_deploy")]
OUT = os.path.join(ROOT, "tests", "golden")
rng = np.random.default_rng(20261018)

# ------------------------------------------------------------------ feet_air_time (reference functions, torch CPU)
import isaaclab  # noqa: E402,F401  (shim: lets the reference module import)
from biped_tasks.tasks.locomotion.velocity.mdp import rewards as ref_rewards  # noqa: E402
from isaaclab.managers import SceneEntityCfg  # noqa: E402

N, STEP_DT = 4096, 0.02


class _Sensor:
    def __init__(self, cur_air, last_air, cur_con):
        self.data = types.SimpleNamespace(current_air_time=cur_air, last_air_time=last_air, current_contact_time=cur_con)

    def compute_first_contact(self, dt, abs_tol=1.0e-8):  # isaaclab 2.1.0 ContactSensor.compute_first_contact
        return (self.data.current_contact_time > 0.0) * (self.data.current_contact_time < (dt + abs_tol))


# timers as a 200 Hz contact sensor produces them: multiples of 5 ms, a foot is either in the air or in contact
steps_air = rng.integers(0, 200, (N, 2)); steps_con = rng.integers(1, 200, (N, 2)); in_con = rng.random((N, 2)) < 0.55
cur_air = np.where(in_con, 0.0, steps_air * 0.005).astype(np.float32)
cur_con = np.where(in_con, steps_con * 0.005, 0.0).astype(np.float32)
cur_con[: N // 4][in_con[: N // 4]] = np.float32(rng.integers(1, 5, in_con[: N // 4].sum()) * 0.005)  # plenty of fresh touch-downs
last_air = (rng.integers(0, 200, (N, 2)) * 0.005).astype(np.float32)
last_con = (rng.integers(0, 200, (N, 2)) * 0.005).astype(np.float32)
cmd = rng.uniform(-1, 1, (N, 3)).astype(np.float32); cmd[: N // 8, :2] *= 0.05  # some below the 0.1 moving threshold
env = types.SimpleNamespace(step_dt=STEP_DT, scene=types.SimpleNamespace(sensors={"contact_forces": _Sensor(*map(torch.from_numpy, (cur_air, last_air, cur_con)))}),
                            command_manager=types.SimpleNamespace(get_command=lambda name: torch.from_numpy(cmd)))
scfg = SceneEntityCfg("contact_forces", body_names=".*ankle_roll_link"); scfg.body_ids = [0, 1]
out = {}
for thr in (0.4, 0.5):
    out[f"l2_thr{thr}"] = ref_rewards.feet_air_time(env, "base_velocity", scfg, thr).numpy()
    out[f"biped_thr{thr}"] = ref_rewards.feet_air_time_positive_biped(env, "base_velocity", thr, scfg).numpy()
np.savez_compressed(os.path.join(OUT, "feet_air_time.npz"), cur_air=cur_air, last_air=last_air, cur_con=cur_con, last_con=last_con, cmd=cmd, **out)

# ------------------------------------------------------------------ observation history (reference CircularBuffer)
from biped_tasks.utils.history.circular_buffer import CircularBuffer  # noqa: E402

sys.modules.setdefault("onnxruntime", types.ModuleType("onnxruntime"))
from biped_deploy.controllers.rl import ObservationHandler  # noqa: E402

NE, T, H = 16, 32, 10
DIMS = [3, 3, 3, 12, 12, 12]  # base_ang_vel, projected_gravity, velocity_commands, joint_pos, joint_vel, actions
quat = rng.normal(size=(T, NE, 4)); quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
oh = ObservationHandler([], [], H, np.zeros(12), {})
grav = np.zeros((T, NE, 3))
for t in range(T):
    for e in range(NE):
        oh.state = {"base_orientation": quat[t, e]}
        grav[t, e] = oh.projected_gravity()
ang = rng.normal(size=(T, NE, 3)); cmdv = rng.uniform(-1, 1, (T, NE, 3)); qrel = rng.normal(size=(T, NE, 12)) * 0.3
qvel = rng.normal(size=(T, NE, 12)) * 3; act = rng.normal(size=(T, NE, 12))
samples = [a.astype(np.float32) for a in (ang, grav, cmdv, qrel, qvel, act)]
reset = rng.random((T, NE)) < 0.08; reset[0] = True
bufs = [CircularBuffer(H, NE, "cpu") for _ in DIMS]
obs = np.zeros((T, NE, 45 * H), np.float32)
for t in range(T):
    ids = np.nonzero(reset[t])[0].tolist()
    for b, s in zip(bufs, samples):
        if t > 0 and ids:
            b.reset(ids)
        b.append(torch.from_numpy(s[t]))
    obs[t] = torch.cat([b.buffer(1).reshape(NE, -1) for b in bufs], dim=-1).numpy()  # observation_manager.py:335-355
np.savez_compressed(os.path.join(OUT, "obs_history.npz"), quat=quat.astype(np.float32), ang=samples[0], grav=samples[1], cmd=samples[2], qrel=samples[3],
                    qvel=samples[4], act=samples[5], reset=reset, obs=obs)

# ------------------------------------------------------------------ deployment-side ObservationHandler (single env)
names = ["base_ang_vel", "projected_gravity", "generated_commands", "joint_pos_rel", "joint_vel_rel", "last_action"]
q0 = np.array([0, -0.16, 0, 0.36, -0.2, 0] * 2)
handler = ObservationHandler(names, [1.0] * 6, H, q0, {"lower": np.array([0.0, -0.5, -1.0]), "upper": np.array([1.0, 0.5, 1.0]), "velocity_deadzone": 0.0})
Td = 25
d_quat = quat[:Td, 0]; d_ang = ang[:Td, 0]; d_q = qrel[:Td, 0] + q0; d_qd = qvel[:Td, 0]; d_act = act[:Td, 0]; d_cmd = rng.uniform(-1, 1, (Td, 3))
d_obs = np.zeros((Td, 45 * H), np.float32)
for t in range(Td):
    state = {"base_orientation": d_quat[t], "base_angular_vel": d_ang[t], "qpos": d_q[t], "qvel": d_qd[t]}
    d_obs[t] = handler.get_observations(state, d_act[t], d_cmd[t])
np.savez_compressed(os.path.join(OUT, "deploy_obs.npz"), quat=d_quat, ang=d_ang, q=d_q, qd=d_qd, act=d_act, cmd_unit=d_cmd, obs=d_obs,
                    cmd_lower=np.array([0.0, -0.5, -1.0]), cmd_upper=np.array([1.0, 0.5, 1.0]))
# the same through the format of the Rsl id / the shipped env.yaml (scripts/deploy/policies/demo_rsl/env.yaml): history 6,
# gyro x0.25, joint velocity x0.05, command ranges +-1
H6 = 6
handler6 = ObservationHandler(names, [0.25, 1.0, 1.0, 1.0, 0.05, 1.0], H6, q0, {"lower": np.array([-1.0, -1.0, -1.0]), "upper": np.array([1.0, 1.0, 1.0]), "velocity_deadzone": 0.0})
d_obs6 = np.zeros((Td, 45 * H6), np.float32)
for t in range(Td):
    state = {"base_orientation": d_quat[t], "base_angular_vel": d_ang[t], "qpos": d_q[t], "qvel": d_qd[t]}
    d_obs6[t] = handler6.get_observations(state, d_act[t], d_cmd[t])
np.savez_compressed(os.path.join(OUT, "deploy_obs_rsl.npz"), quat=d_quat, ang=d_ang, q=d_q, qd=d_qd, act=d_act, cmd_unit=d_cmd, obs=d_obs6,
                    cmd_lower=np.array([-1.0, -1.0, -1.0]), cmd_upper=np.array([1.0, 1.0, 1.0]))
print("wrote feet_air_time.npz obs_history.npz deploy_obs.npz deploy_obs_rsl.npz")

# ------------------------------------------------------------------ dead-zone command class (Rsl id, velocity_deadzone = 0)
# biped_tasks/utils/mdp/commands.py:41-96 UniformVelocityCommandWithDeadzone._update_command, the reference's own method, run on
# an instance that holds exactly the attributes the method reads (no simulator): what survives of a command after k calls.
import json  # noqa: E402

from biped_tasks.utils.mdp import commands as ref_commands  # noqa: E402

torch.manual_seed(20261018)
ND, CALLS, PHYS_DT, EP_S = 4096, 12, 0.005, 20.0
term = object.__new__(ref_commands.UniformVelocityCommandWithDeadzone)
term.cfg = types.SimpleNamespace(heading_command=False, heading_control_stiffness=1.0, ranges=types.SimpleNamespace(ang_vel_z=(-1.0, 1.0)))
term.velocity_deadzone = 0.0          # C12/rsl_env_cfg.py:98
term.dt, term.max_episode_length_s = PHYS_DT, EP_S  # commands.py:37-38
term.vel_command_b = torch.from_numpy(rng.uniform(-1, 1, (ND, 3)).astype(np.float32))
term.is_standing_env = torch.zeros(ND, dtype=torch.bool); term.is_standing_env[:64] = True  # the override must NOT zero these
first = term.vel_command_b.clone()
zero_xy, flips = [], 0
for _ in range(CALLS):
    before = term.vel_command_b[:, 2].clone()
    term._update_command()
    zero_xy.append(int(((term.vel_command_b[:, 0] == 0) & (term.vel_command_b[:, 1] == 0)).sum()))
    flips += int((term.vel_command_b[:, 2] == -before).sum())
    assert torch.equal(term.vel_command_b[:, 2].abs(), before.abs())
# long run for the flip probability (it is 2.5e-4 per call)
for _ in range(2000):
    before = term.vel_command_b[:, 2].clone()
    term._update_command()
    flips += int((term.vel_command_b[:, 2] == -before).sum())
# positive dead zone (the class default 0.1, the CaT cfg's 0.2): the balancing branch keeps exactly n // 2 envs inside.  _resample
# belongs to the upstream base class (absent here); the stand-in redraws the command uniformly over the cfg ranges as upstream does.
def _balanced_counts(deadzone, calls=30):
    t = object.__new__(ref_commands.UniformVelocityCommandWithDeadzone)
    t.cfg = term.cfg; t.velocity_deadzone = deadzone; t.dt, t.max_episode_length_s = PHYS_DT, EP_S
    t.vel_command_b = torch.from_numpy(rng.uniform(-1, 1, (ND, 3)).astype(np.float32))
    t._resample = lambda ids: t.vel_command_b.__setitem__(ids, torch.from_numpy(rng.uniform(-1, 1, (len(ids), 3)).astype(np.float32)))
    counts = [int((torch.norm(t.vel_command_b[:, :2], dim=1) < deadzone).sum())]
    for _ in range(calls):
        t._update_command()
        counts.append(int((torch.norm(t.vel_command_b[:, :2], dim=1) < deadzone).sum()))
    return counts


balanced = {str(dz): _balanced_counts(dz) for dz in (0.1, 0.2)}
json.dump({"n_envs": ND, "calls": CALLS, "zero_xy_after_call": zero_xy, "yaw_flips": flips, "flip_trials": ND * (CALLS + 2000),
           "in_deadzone_count_before_and_after_each_call": balanced,
           "standing_envs_yaw_untouched_by_zeroing": bool(torch.equal(term.vel_command_b[:64, 2].abs(), first[:64, 2].abs())),
           "physics_dt": PHYS_DT, "max_episode_length_s": EP_S},
          open(os.path.join(OUT, "deadzone_command.json"), "w"), indent=1)
print("wrote deadzone_command.json", zero_xy, flips)

# ------------------------------------------------------------------ contact / limit idioms (in-tree constraint functions)
# biped_tasks/utils/cat/constraints.py holds in-tree bodies of the idioms the Flat / Rsl terms use upstream:
#   contact (:86-99)                 any_b max_h |F_hb| > 1.0            == mdp.illegal_contact(threshold=1.0)     (termination)
#   foot_contact_force (:161-168)    max_h |F_hb| - limit                == the per-body argument of mdp.contact_forces' clip
#   joint_position_limits (:22-31)   max(lo - q, q - hi)                 == the per-joint argument of mdp.joint_pos_limits' clip
# Evaluated on random sensor histories / joint positions through a stand-in env.
from biped_tasks.utils.cat import constraints as ref_cstr  # noqa: E402

NC = 2048
Fh = rng.normal(size=(NC, 3, 6, 3)).astype(np.float32)               # net_forces_w_history [N, H=3, B=6 sensor slots, 3]
scale = np.exp(rng.uniform(np.log(0.02), np.log(3000.0), (NC, 1, 6, 1))).astype(np.float32)
Fh *= scale
Fh[(rng.random((NC, 1, 6, 1)) < 0.75).repeat(3, 1).repeat(3, 3)] = 0.0            # bodies in the air
near = rng.random((NC, 6)) < 0.1                                                   # norms right at the 1.0 N threshold
Fh = Fh.transpose(0, 2, 1, 3).copy()                                               # [N, B, H, 3] to index by (env, body)
Fh[near] = 0.0
Fh[near, 2] = (np.array([1.0, 0.0, 0.0], np.float32) * (1.0 + rng.choice([-1e-4, 1e-4], (int(near.sum()), 1)))).astype(np.float32)
Fh = Fh.transpose(0, 2, 1, 3).copy()
lo = np.array([-0.43, -3.14, -0.43, -0.26, -0.897334, -0.261799, -0.43, -3.14, -3.14, -0.26, -0.897334, -0.261799])
hi = np.array([0.43, 2.5, 3.14, 2.05, 0.523598, 0.261799, 0.43, 2.5, 0.43, 2.05, 0.523598, 0.261799])
mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * 0.9                                   # A/robots/h12.py:56 soft_joint_pos_limit_factor
soft = np.stack([mid - half, mid + half], -1).astype(np.float32)
qj = (mid + rng.uniform(-1.15, 1.15, (NC, 12)) * half).astype(np.float32)           # some beyond the soft limits
sensor = types.SimpleNamespace(data=types.SimpleNamespace(net_forces_w_history=torch.from_numpy(Fh)))
robot = types.SimpleNamespace(data=types.SimpleNamespace(joint_pos=torch.from_numpy(qj), soft_joint_pos_limits=torch.from_numpy(soft)[None].repeat(NC, 1, 1)))
cenv = types.SimpleNamespace(scene={"contact_forces": sensor, "robot": robot})
ill = SceneEntityCfg("contact_forces"); ill.body_ids = [2, 3, 4, 5]
feet = SceneEntityCfg("contact_forces"); feet.body_ids = [0, 1]
jall = SceneEntityCfg("robot"); jall.joint_ids = list(range(12))
np.savez_compressed(os.path.join(OUT, "contact_limit_idioms.npz"), force_hist=Fh, joint_pos=qj, soft_limits=soft,
                    illegal=ref_cstr.contact(cenv, ill).numpy(), foot_force_800=ref_cstr.foot_contact_force(cenv, 800.0, feet).numpy(),
                    pos_limit=ref_cstr.joint_position_limits(cenv, jall).numpy())
print("wrote contact_limit_idioms.npz")
