"""CPU restatement (numpy, float32 like the reference's torch code) of the Constraints-as-Terminations tail -- SURVEY.md 8(f) rank 3.

TEST INFRASTRUCTURE, like everything under oracle/: only tests/ may import it.  It is the pinned specification the CUDA path of the CaT
variant (the constraint columns inside step_kernel<.., CAT> + cat_apply_kernel, csrc/h1v2_cat.cuh; h1v2_cat_step) is checked against in
tests/test_gpu_cat.py.  Pinned by
tests/golden/cat_sequence.npz, which tests/golden/make_cat_goldens.py produced by running the reference's own functions.

Reference (paths relative to packages/biped_tasks/biped_tasks/):
  utils/cat/constraints.py                     the constraint bodies; each function below cites its lines
  utils/cat/constraint_manager.py:23-86        CaT: running maxima, termination probabilities
  utils/cat/constraint_manager.py:213-229      ConstraintManager.compute: statistics per term
  utils/cat/curriculums.py:20-42               modify_constraint_p
  utils/cat/cat_env.py:147-153                 reward *= 1 - p ; dones = p
  tasks/locomotion/velocity/config/h12_12dof/cat_env_cfg.py:336-431   which constraints, with which parameters

Several of these operations couple the envs of one process inside a step (column maxima over all envs, the gather of no_move);
they are written here exactly as the reference computes them, oddities included.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def hist_norm_max(force_hist: np.ndarray, bodies) -> np.ndarray:
    """max over the history of |F| per body: the idiom of constraints.py:91-98,167-168,178-187.  force_hist [N,H,B,3] -> [N,len(bodies)]"""
    return np.linalg.norm(force_hist[:, :, bodies].astype(F), axis=-1).astype(F).max(axis=1)


def contact(force_hist, bodies):  # constraints.py:86-99
    return (hist_norm_max(force_hist, bodies) > F(1.0)).any(axis=1)


def joint_position_limits(joint_pos, soft_limits):  # constraints.py:22-31
    return np.maximum(soft_limits[:, 0] - joint_pos, joint_pos - soft_limits[:, 1]).astype(F)


def joint_velocity_limits(joint_vel, limits):  # constraints.py:34-43
    return (np.abs(joint_vel) - limits).astype(F)


def joint_torque_limits(torque, limits):  # constraints.py:46-55
    return (np.abs(torque) - limits).astype(F)


def foot_contact_force(force_hist, feet, limit):  # constraints.py:161-168
    return (hist_norm_max(force_hist, feet) - F(limit)).astype(F)


def foot_contact(force_hist, feet):  # constraints.py:171-194: number of feet in contact is not 1 or 2
    k = (hist_norm_max(force_hist, feet) > F(1.0)).sum(axis=1)
    return ((k < 1) | (k > 2)).astype(F)


def no_move(cmd, joint_vel, velocity_deadzone, joint_vel_limit):  # constraints.py:197-235
    """Rows of the envs whose whole command is inside the dead zone are gathered, and that block is TILED over the batch: row i of
    the result is |joint_vel| - limit of the (i mod K)-th such env, whatever env i itself does.  All zeros when K == 0."""
    inactive = (np.abs(cmd[:, :3]) < F(velocity_deadzone)).all(axis=1)
    n = cmd.shape[0]
    if inactive.sum() == 0:
        return np.zeros((n, joint_vel.shape[1]), F)
    block = (np.abs(joint_vel[inactive]) - F(joint_vel_limit)).astype(F)
    reps = n // block.shape[0] + 1
    return np.tile(block, (reps, 1))[:n]


def base_orientation(projected_gravity, limit):  # constraints.py:102-108
    return (np.linalg.norm(projected_gravity[:, :2].astype(F), axis=1).astype(F) - F(limit)).astype(F)


def base_height(root_z, height, std):  # constraints.py:256-272
    return ((root_z < F(height - std)) | (root_z > F(height + std))).astype(F)


def first_contact(contact_time, dt, abs_tol=1.0e-8):  # isaaclab 2.1.0 ContactSensor.compute_first_contact
    return (contact_time > 0.0) & (contact_time < F(dt + abs_tol))


class FootClearance:  # constraints.py:275-308 (the tracker lives on the asset's data object upstream)
    def __init__(self):
        self.swing_max_height = None

    def __call__(self, foot_z, touchdown, cmd, min_height, velocity_deadzone):
        if self.swing_max_height is None:
            self.swing_max_height = np.zeros_like(foot_z, dtype=F)
        violation = (F(min_height) - self.swing_max_height) * touchdown.astype(F)
        self.swing_max_height = np.where(~touchdown, np.maximum(self.swing_max_height, foot_z), F(0.0)).astype(F)
        active = (np.abs(cmd[:, :3]) > F(velocity_deadzone)).any(axis=1).astype(F)[:, None]
        return (violation * active).astype(F)


class CaT:  # constraint_manager.py:23-86
    def __init__(self, tau=0.95, min_p=0.0):
        self.tau, self.min_p = F(tau), F(min_p)
        self.running_maxes, self.probs = {}, {}

    def add(self, name, constraint, max_p):
        c = np.asarray(constraint, F)
        c = c[:, None] if c.ndim == 1 else c
        cmax = np.maximum(c.max(axis=0, keepdims=True), F(1e-6))  # over ALL envs of this step
        if name in self.running_maxes:
            self.running_maxes[name] = (self.running_maxes[name] * self.tau + (F(1.0) - self.tau) * cmax).astype(F)
        else:
            self.running_maxes[name] = cmax.astype(F)
        probs = np.zeros_like(c)
        mask = c > 0
        norm = (c / self.running_maxes[name]).astype(F)
        probs[mask] = (self.min_p + np.clip(norm[mask], 0.0, 1.0) * (F(max_p) - self.min_p)).astype(F)
        self.probs[name] = probs
        return probs

    def get_probs(self):
        return np.concatenate(list(self.probs.values()), axis=1).max(axis=1)


def manager_statistics(probs: dict):  # constraint_manager.py:221-227: what one step adds to the per-term episode sums
    return {k: ((p.max(axis=1) > 0).astype(F), p.max(axis=1)) for k, p in probs.items()}


def constrained_reward(reward, cstr_prob, reset):  # cat_env.py:147-153,166-169
    """reward_buf = reward * (1 - p); dones = p, and 1 for the envs that reset."""
    dones = cstr_prob.astype(F).copy()
    dones[reset] = F(1.0)
    return (reward * (F(1.0) - cstr_prob)).astype(F), dones


def modify_constraint_p(common_step_counter, num_steps, init_max_p):  # curriculums.py:20-42
    progress = min(common_step_counter / num_steps, 1.0)
    t_start, t_end = 20, 1 / init_max_p
    return 1 / (t_start + progress * (t_end - t_start))
