"""ctypes wrapper of the CPU oracle (oracle/h1v2_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from h1v2_isaac_b200._capi import (  # struct layouts are the shared ABI, not code under test
    H1v2Config, H1v2State, LOG_DIM, NJ, OBS_TERM_DIM, STATE_FIELDS, READ_ONLY_STATE, state_field_count,
)

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_build", "libh1v2_oracle.so")
_lib = None


_SO_COUNTED = os.path.join(_DIR, "_build", "libh1v2_oracle_counted.so")
_SRCS = [os.path.join(_DIR, "h1v2_oracle.c"), os.path.join(_DIR, "h1v2_oracle.h"), os.path.join(_DIR, "flopcount", "counted_double.h"),
         os.path.join(_DIR, "..", "h1v2_isaac_b200", "csrc", "h1v2_config.cpp"), os.path.join(_DIR, "..", "include", "h1v2_b200.h")]


def build(force: bool = False) -> str:
    newest = max(os.path.getmtime(f) for f in _SRCS if os.path.exists(f))
    if force or not os.path.exists(_SO) or not os.path.exists(_SO_COUNTED) or min(os.path.getmtime(_SO), os.path.getmtime(_SO_COUNTED)) < newest:
        subprocess.check_call(["make", "-C", _DIR, "--quiet"])
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.h1v2o_create.argtypes = [C.POINTER(H1v2Config), C.c_int32, C.c_uint64, C.POINTER(C.c_void_p)]
        L.h1v2o_destroy.argtypes = [C.c_void_p]
        L.h1v2o_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.h1v2o_obs_dim.argtypes = [C.c_void_p]
        L.h1v2o_max_episode_length.argtypes = [C.c_void_p]
        L.h1v2o_max_episode_length.restype = C.c_int64
        L.h1v2o_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.h1v2o_observe.argtypes = [C.c_void_p, C.c_void_p]
        L.h1v2o_step.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.h1v2o_step_injected.argtypes = [C.c_void_p] + [C.c_void_p] * 12
        L.h1v2o_get_state.argtypes = [C.c_void_p, C.POINTER(H1v2State)]
        L.h1v2o_set_state.argtypes = [C.c_void_p, C.POINTER(H1v2State)]
        L.h1v2o_get_episode_length.argtypes = [C.c_void_p, C.c_void_p]
        L.h1v2o_set_episode_length.argtypes = [C.c_void_p, C.c_void_p]
        L.h1v2o_get_log.argtypes = [C.c_void_p, C.c_void_p]
        L.h1v2o_set_reward_weights.argtypes = [C.c_void_p, C.c_void_p]
        L.h1v2o_solver_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.h1v2o_activation_margin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.h1v2o_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.h1v2o_rng4.argtypes = [C.c_uint64, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        L.h1v2o_fk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.h1v2o_mass_matrix.argtypes = [C.POINTER(H1v2Config), C.c_void_p, C.c_void_p]
        L.h1v2o_bias.argtypes = [C.POINTER(H1v2Config), C.c_void_p, C.c_void_p, C.c_void_p]
        L.h1v2o_physics_step.argtypes = [C.POINTER(H1v2Config), C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.h1v2o_total_energy.argtypes = [C.POINTER(H1v2Config), C.c_void_p, C.c_void_p]
        L.h1v2o_total_energy.restype = C.c_double
        L.h1v2o_wrap_to_pi.argtypes = [C.c_float]
        L.h1v2o_wrap_to_pi.restype = C.c_float
        L.h1v2o_terrain_dims.argtypes = [C.c_void_p, C.c_void_p]
        L.h1v2o_get_terrain.argtypes = [C.c_void_p, C.c_void_p]
        L.h1v2o_set_terrain.argtypes = [C.c_void_p, C.c_void_p]
        L.h1v2o_terrain_level_mean.argtypes = [C.c_void_p]
        L.h1v2o_terrain_level_mean.restype = C.c_float
        L.h1v2o_terrain_query.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.h1v2o_tri_margin.argtypes = [C.c_void_p, C.c_void_p]
        for f in ("h1v2_default_config", "h1v2_rsl_config", "h1v2_cat_config", "h1v2_rough_config"):  # the oracle library's own copy (h1v2_config.cpp, host only)
            getattr(L, f).argtypes = [C.POINTER(H1v2Config)]
        _lib = L
    return _lib


def task_config(task: str = "flat") -> H1v2Config:
    """Resolved cfg of the Flat / Rsl / CaT id from the ORACLE library (so that the reference arm of bench.py and the CPU tests
    never load the CUDA library)."""
    cfg = H1v2Config()
    rc = getattr(lib(), {"flat": "h1v2_default_config", "rsl": "h1v2_rsl_config", "cat": "h1v2_cat_config", "rough": "h1v2_rough_config"}[task])(C.byref(cfg))
    assert rc == 0
    return cfg


def count_flops(cfg: H1v2Config, n_envs: int, steps: int, actions=None, settle: int = 0, seed: int = 3) -> dict:
    """Exact double-precision operation counts of the oracle per env-step (the binding FLOP figure of the FP32
    roofline).  Runs libh1v2_oracle_counted.so -- this file's source compiled with `double` replaced by a counting scalar
    (flopcount/counted_double.h), outputs bit-identical to the plain build -- single-threaded.  actions: None = zero actions
    (the double-support standing state once `settle` steps have passed), else a callable step -> [n,12] float32."""
    global _lib, _SO
    build()
    saved = (_lib, _SO)
    _lib, _SO = None, _SO_COUNTED
    try:
        L = lib()
        L.h1v2o_flops_get.argtypes = [C.c_void_p]
        orc = Oracle(cfg, n_envs, seed=seed, threads=1)
        orc.observe()
        zero = np.zeros((n_envs, NJ), np.float32)
        for s in range(settle):
            orc.step(zero if actions is None else actions(s))
        L.h1v2o_flops_reset()
        for s in range(steps):
            orc.step(zero if actions is None else actions(settle + s))
        c = np.zeros(7, np.uint64)
        L.h1v2o_flops_get(_p(c))
        it, _ = orc.solver_stats()
        del orc
    finally:
        _lib, _SO = saved
    per = c.astype(np.float64) / (n_envs * steps)
    names = ["add", "mul", "div", "fma", "sqrt", "transcendental", "compare"]
    out = {k: float(v) for k, v in zip(names, per)}
    out["flops"] = float(per[0] + per[1] + per[2] + 2 * per[3] + per[4] + per[5])  # compares are not counted as flops
    out["mean_newton_iters_last_substep"] = float(it.mean())
    return out


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


_NP = {C.c_float: np.float32, C.c_int32: np.int32}


class Oracle:
    """Host twin of h1v2_isaac_b200.backend.H1v2Sim (same method names and array layouts)."""

    def __init__(self, cfg: H1v2Config, n_envs: int, seed: int = 42, threads: int = 1):
        self.cfg = cfg.copy()
        self.n = n_envs
        self._h = C.c_void_p()
        rc = lib().h1v2o_create(C.byref(self.cfg), n_envs, seed, C.byref(self._h))
        assert rc == 0
        lib().h1v2o_set_threads(self._h, threads)
        self.obs_dim = lib().h1v2o_obs_dim(self._h)
        self.max_episode_length = lib().h1v2o_max_episode_length(self._h)
        self.history = cfg.history_length

    def __del__(self):
        if getattr(self, "_h", None):
            lib().h1v2o_destroy(self._h)
            self._h = None

    def reset(self, env_ids=None):
        if env_ids is None:
            lib().h1v2o_reset(self._h, None, 0)
        else:
            ids = np.ascontiguousarray(env_ids, dtype=np.int64)
            lib().h1v2o_reset(self._h, _p(ids), len(ids))

    def observe(self) -> np.ndarray:
        obs = np.zeros((self.n, self.obs_dim), np.float32)
        lib().h1v2o_observe(self._h, _p(obs))
        return obs

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        assert a.shape == (self.n, NJ)
        obs = np.zeros((self.n, self.obs_dim), np.float32)
        rew = np.zeros(self.n, np.float32)
        term = np.zeros(self.n, np.uint8)
        trunc = np.zeros(self.n, np.uint8)
        lib().h1v2o_step(self._h, _p(a), _p(obs), _p(rew), _p(term), _p(trunc))
        return obs, rew, term.astype(bool), trunc.astype(bool)

    def step_injected(self, actions, post: dict):
        """Control step with the physics loop replaced by the given post-physics values (identical-state tail parity).
        post keys: pre_reset_qpos, pre_reset_qvel, pre_reset_timers, slot_force_hist, applied_torque, joint_acc and, optionally,
        foot_vel (without it the oracle computes the foot velocities from the injected state with its own kinematics)."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        f = {k: np.ascontiguousarray(post[k], dtype=np.float32) for k in
             ("pre_reset_qpos", "pre_reset_qvel", "pre_reset_timers", "slot_force_hist", "applied_torque", "joint_acc", "foot_vel") if k in post}
        obs = np.zeros((self.n, self.obs_dim), np.float32)
        rew = np.zeros(self.n, np.float32)
        term = np.zeros(self.n, np.uint8)
        trunc = np.zeros(self.n, np.uint8)
        lib().h1v2o_step_injected(self._h, _p(a), _p(f["pre_reset_qpos"]), _p(f["pre_reset_qvel"]), _p(f["pre_reset_timers"]),
                                  _p(f["slot_force_hist"]), _p(f["applied_torque"]), _p(f["joint_acc"]), _p(f["foot_vel"]) if "foot_vel" in f else None,
                                  _p(obs), _p(rew), _p(term), _p(trunc))
        return obs, rew, term.astype(bool), trunc.astype(bool)

    def get_state(self, names=None) -> dict:
        out, st = {}, H1v2State()
        for name, _, ct in STATE_FIELDS:
            if names is not None and name not in names:
                continue
            cnt = state_field_count(name, self.history)
            arr = np.zeros((self.n, cnt), _NP[ct])
            out[name] = arr
            setattr(st, name, arr.ctypes.data_as(C.POINTER(ct)))
        lib().h1v2o_get_state(self._h, C.byref(st))
        return out

    def set_state(self, state: dict):
        st, keep = H1v2State(), []
        for name, _, ct in STATE_FIELDS:
            if name in state and name not in READ_ONLY_STATE:
                cnt = state_field_count(name, self.history)
                arr = np.ascontiguousarray(np.asarray(state[name]).reshape(self.n, cnt), dtype=_NP[ct])
                keep.append(arr)
                setattr(st, name, arr.ctypes.data_as(C.POINTER(ct)))
        lib().h1v2o_set_state(self._h, C.byref(st))

    @property
    def episode_length(self) -> np.ndarray:
        out = np.zeros(self.n, np.int64)
        lib().h1v2o_get_episode_length(self._h, _p(out))
        return out

    @episode_length.setter
    def episode_length(self, v):
        a = np.ascontiguousarray(v, dtype=np.int64)
        lib().h1v2o_set_episode_length(self._h, _p(a))

    def set_reward_weights(self, w):
        a = np.ascontiguousarray(w, dtype=np.float32)
        assert a.size == len(self.cfg.rew_weight)
        lib().h1v2o_set_reward_weights(self._h, _p(a))

    def log(self) -> np.ndarray:
        out = np.zeros(LOG_DIM, np.float32)
        lib().h1v2o_get_log(self._h, _p(out))
        return out

    def activation_margin(self):
        """(contact, limit): smallest distance to an activation boundary at any substep start of the last step."""
        c = np.zeros(self.n, np.float64)
        l = np.zeros(self.n, np.float64)
        lib().h1v2o_activation_margin(self._h, _p(c), _p(l))
        return c, l

    # ---- rough terrain ----
    def terrain(self) -> np.ndarray:
        d = np.zeros(2, np.int32)
        assert lib().h1v2o_terrain_dims(self._h, _p(d)) == 0, "no terrain (cfg.terrain_enable)"
        out = np.zeros((int(d[0]), int(d[1])), np.float32)
        lib().h1v2o_get_terrain(self._h, _p(out))
        return out

    def set_terrain(self, heights):
        a = np.ascontiguousarray(heights, dtype=np.float32)
        assert lib().h1v2o_set_terrain(self._h, _p(a)) == 0

    def terrain_level_mean(self) -> float:
        return float(lib().h1v2o_terrain_level_mean(self._h))

    def terrain_query(self, level: int, ttype: int, lx: float, ly: float):
        h = np.zeros(1)
        n = np.zeros(3)
        assert lib().h1v2o_terrain_query(self._h, level, ttype, lx, ly, _p(h), _p(n)) == 0
        return float(h[0]), n

    def tri_margin(self) -> np.ndarray:
        out = np.zeros(self.n, np.float64)
        lib().h1v2o_tri_margin(self._h, _p(out))
        return out

    def solver_stats(self):
        it = np.zeros(self.n, np.int32)
        res = np.zeros(self.n, np.float64)
        lib().h1v2o_solver_stats(self._h, _p(it), _p(res))
        return it, res


# ---- building blocks for known-answer tests ----
def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().h1v2o_philox(_p(c), _p(k), _p(o))
    return o


def rng4(seed, env, step, stream, block):
    u = np.zeros(4, np.float32)
    lib().h1v2o_rng4(seed, env, step, stream, block, _p(u))
    return u


def fk(qpos):
    q = np.ascontiguousarray(qpos, np.float64)
    R = np.zeros((13, 3, 3))
    x = np.zeros((13, 3))
    lib().h1v2o_fk(_p(q), _p(R), _p(x))
    return R, x


def mass_matrix(cfg, qpos):
    q = np.ascontiguousarray(qpos, np.float64)
    M = np.zeros((18, 18))
    lib().h1v2o_mass_matrix(C.byref(cfg), _p(q), _p(M))
    return M


def bias(cfg, qpos, qvel):
    q = np.ascontiguousarray(qpos, np.float64)
    v = np.ascontiguousarray(qvel, np.float64)
    b = np.zeros(18)
    lib().h1v2o_bias(C.byref(cfg), _p(q), _p(v), _p(b))
    return b


def physics_step(cfg, qpos, qvel, ctrl, friction=None):
    q = np.array(qpos, np.float64)
    v = np.array(qvel, np.float64)
    u = np.ascontiguousarray(ctrl, np.float64)
    sf = np.zeros((6, 3))
    it = np.zeros(1, np.int32)
    res = np.zeros(1)
    lib().h1v2o_physics_step(C.byref(cfg), _p(q), _p(v), _p(u), cfg.friction if friction is None else friction,
                             _p(sf), _p(it), _p(res))
    return q, v, sf, int(it[0]), float(res[0])


def total_energy(cfg, qpos, qvel):
    q = np.ascontiguousarray(qpos, np.float64)
    v = np.ascontiguousarray(qvel, np.float64)
    return lib().h1v2o_total_energy(C.byref(cfg), _p(q), _p(v))


def wrap_to_pi(a: float) -> float:
    return lib().h1v2o_wrap_to_pi(a)
