"""H1v2Sim: thin torch-facing wrapper over the C-ABI (include/h1v2_b200.h).

PyTorch is used only for device memory and streams; all arithmetic of the step happens in the CUDA library.
There is no fallback: constructing H1v2Sim without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _capi
from ._capi import (H1v2Config, H1v2State, LOG_DIM, NJ, READ_ONLY_STATE, STATE_FIELDS, default_config, load_library,
                    state_field_count)

_TORCH_DT = {C.c_float: torch.float32, C.c_int32: torch.int32}


class _DevPtr:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


class H1v2Sim:
    def __init__(self, num_envs: int, cfg: H1v2Config | None = None, device: str | torch.device = "cuda:0",
                 seed: int = 42, diagnostics: bool = False, lib_path: str | None = None):
        self._lib = load_library(lib_path)  # lib_path: a build variant of the same ABI (tests: the double-precision build)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("H1v2Sim runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("H1v2Sim: no CUDA device visible; the CUDA kernels are the product, nothing to fall back to")
        self.cfg = (cfg or default_config()).copy()
        if diagnostics:
            self.cfg.reserved[0] = 1
        if os.environ.get("H1V2_PLAIN_KERNEL") == "1":  # A/B switch: the plain instantiations instead of the mirror-lane ones (cfg.reserved[3])
            self.cfg.reserved[3] = 1
        self.num_envs = int(num_envs)
        self._h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)  # "cuda" without an index would never compare equal to a tensor's device
        with torch.cuda.device(idx):
            rc = self._lib.h1v2_create(C.byref(self.cfg), self.num_envs, idx, seed, C.byref(self._h))
            if rc != 0:
                raise RuntimeError("h1v2_create: " + self._lib.h1v2_last_error().decode())
            self.obs_dim = self._lib.h1v2_obs_dim(self._h)
            self.history = self.cfg.history_length
            step_dt = self.cfg.sim_dt * self.cfg.decimation
            import math
            self.max_episode_length = int(math.ceil(self.cfg.episode_length_s / step_dt * (1.0 - 1e-6)))
            self.episode_length_buf = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)
            self._check(self._lib.h1v2_bind_episode_length(self._h, self.episode_length_buf.data_ptr()))
            p = C.c_void_p()
            self._check(self._lib.h1v2_get_log(self._h, C.byref(p)))
            self.log_buf = torch.as_tensor(_DevPtr(p.value, LOG_DIM), device=self.device)
            self.terrain_log_buf = None
            if self.cfg.terrain_enable:  # [0] running sum, [1] Curriculum/terrain_levels after the last step (device view)
                self._check(self._lib.h1v2_get_terrain_log(self._h, C.byref(p)))
                self.terrain_log_buf = torch.as_tensor(_DevPtr(p.value, 2), device=self.device)
            self.cat_acc_buf = None
            if self.cfg.cat_enable:  # sums behind Episode_Constraint_* (violation x100 [10], probability [10], count), device view
                self._check(self._lib.h1v2_get_cat_log(self._h, C.byref(p)))
                self.cat_acc_buf = torch.as_tensor(_DevPtr(p.value, 21), device=self.device)

    # ------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError(self._lib.h1v2_last_error().decode())

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.h1v2_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def step(self, actions: torch.Tensor):
        """One control step.  actions: float32 [N,12] on this device (external joint order)."""
        a = actions
        if a.dtype != torch.float32 or not a.is_contiguous() or a.device != self.device:
            a = a.to(device=self.device, dtype=torch.float32).contiguous()
        if a.shape != (self.num_envs, NJ):
            raise ValueError(f"actions must be [{self.num_envs},{NJ}], got {tuple(a.shape)}")
        obs = torch.empty((self.num_envs, self.obs_dim), dtype=torch.float32, device=self.device)
        rew = torch.empty(self.num_envs, dtype=torch.float32, device=self.device)
        term = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        trunc = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        self._check(self._lib.h1v2_step(self._h, a.data_ptr(), obs.data_ptr(), rew.data_ptr(), term.data_ptr(),
                                        trunc.data_ptr(), self._stream()))
        return obs, rew, term.view(torch.bool), trunc.view(torch.bool)

    def cat_step(self, actions: torch.Tensor):
        """Constraints-as-Terminations step (CaTEnv.step): -> obs, rew * (1 - p), dones (float: p, 1 on reset), truncated."""
        a = actions
        if a.dtype != torch.float32 or not a.is_contiguous() or a.device != self.device:
            a = a.to(device=self.device, dtype=torch.float32).contiguous()
        if a.shape != (self.num_envs, NJ):
            raise ValueError(f"actions must be [{self.num_envs},{NJ}], got {tuple(a.shape)}")
        obs = torch.empty((self.num_envs, self.obs_dim), dtype=torch.float32, device=self.device)
        rew = torch.empty(self.num_envs, dtype=torch.float32, device=self.device)
        dones = torch.empty(self.num_envs, dtype=torch.float32, device=self.device)
        trunc = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        self._check(self._lib.h1v2_cat_step(self._h, a.data_ptr(), obs.data_ptr(), rew.data_ptr(), dones.data_ptr(), trunc.data_ptr(), self._stream()))
        return obs, rew, dones, trunc.view(torch.bool)

    def cat_step_into(self, actions, obs, rew, dones, trunc):
        """cat_step() with caller-provided output tensors (dones float32 [N])."""
        self._check(self._lib.h1v2_cat_step(self._h, actions.data_ptr(), obs.data_ptr(), rew.data_ptr(), dones.data_ptr(), trunc.data_ptr(), self._stream()))

    def set_constraint_max_p(self, max_p) -> None:
        import numpy as np
        p = np.ascontiguousarray(max_p, dtype=np.float32)
        self._check(self._lib.h1v2_set_constraint_max_p(self._h, p.ctypes.data_as(C.POINTER(C.c_float))))
        for i in range(p.size):
            self.cfg.cat_max_p[i] = float(p[i])

    def cat_debug(self):
        """(raw [56,N], probs [56,N], running_max [56]) of the last cat step, host copies."""
        import numpy as np
        from ._capi import CSTR_COLS
        raw = np.zeros((CSTR_COLS, self.num_envs), np.float32); probs = np.zeros_like(raw); rm = np.zeros(CSTR_COLS, np.float32)
        self._check(self._lib.h1v2_cat_debug(self._h, raw.ctypes.data, probs.ctypes.data, rm.ctypes.data))
        return raw, probs, rm

    def cat_log_host(self):
        import numpy as np
        from ._capi import NUM_CSTR
        out = np.zeros(2 * NUM_CSTR + 1, np.float32)
        self._check(self._lib.h1v2_get_cat_log_host(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def step_into(self, actions, obs, rew, term, trunc):
        """Same as step() with caller-provided output tensors (no allocation; used by bench.py and CUDA graphs)."""
        self._check(self._lib.h1v2_step(self._h, actions.data_ptr(), obs.data_ptr(), rew.data_ptr(), term.data_ptr(),
                                        trunc.data_ptr(), self._stream()))

    def step_host(self, actions, obs, rew, term, trunc):
        """C-ABI call with HOST buffers (numpy or pinned CPU tensors); copies inside, synchronises."""
        def ptr(x):
            return x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data
        self._check(self._lib.h1v2_step_host(self._h, ptr(actions), ptr(obs), ptr(rew), ptr(term), ptr(trunc)))

    def cat_step_host(self, actions, obs, rew, dones, trunc):
        """h1v2_cat_step_host: the CaT step with HOST buffers (dones float32 [N])."""
        def ptr(x):
            return x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data
        self._check(self._lib.h1v2_cat_step_host(self._h, ptr(actions), ptr(obs), ptr(rew), ptr(dones), ptr(trunc)))

    # ---- rough terrain (cfg.terrain_enable) ----
    def terrain(self):
        """Host copy of the height field [grid_x, grid_y] in metres (h1v2_get_terrain)."""
        import numpy as np
        d = (C.c_int32 * 2)()
        self._check(self._lib.h1v2_terrain_dims(self._h, d))
        out = np.zeros((d[0], d[1]), np.float32)
        self._check(self._lib.h1v2_get_terrain(self._h, out.ctypes.data))
        return out

    def set_terrain(self, heights) -> None:
        """Replace the generated height field, e.g. by the grid a real isaaclab TerrainGenerator produced (h1v2_set_terrain)."""
        import numpy as np
        d = (C.c_int32 * 2)()
        self._check(self._lib.h1v2_terrain_dims(self._h, d))
        a = np.ascontiguousarray(heights, dtype=np.float32)
        if a.shape != (d[0], d[1]):
            raise ValueError(f"terrain must be [{d[0]},{d[1]}], got {a.shape}")
        self._check(self._lib.h1v2_set_terrain(self._h, a.ctypes.data))

    def check_guards(self) -> int:
        """Guard bytes around the handle's device arrays that were overwritten (0 = no out-of-bounds store); synchronises."""
        return int(self._lib.h1v2_check_guards(self._h))

    def host_path_info(self):
        """(mode, threads) of step_host: 0 = rows written by the kernel, 1 = samples over PCIe + host assembly, -1 = undecided."""
        m, t = C.c_int32(-1), C.c_int32(0)
        self._check(self._lib.h1v2_host_path_info(self._h, C.byref(m), C.byref(t)))
        return int(m.value), int(t.value)

    def host_path_rows(self) -> int:
        """Envs whose rows the kernel writes into the caller's buffer itself (N: mode rows; 0..N in mode assemble, > 0 = hybrid)."""
        return int(self._lib.h1v2_host_path_rows(self._h))

    def observe(self) -> torch.Tensor:
        obs = torch.empty((self.num_envs, self.obs_dim), dtype=torch.float32, device=self.device)
        self._check(self._lib.h1v2_observe(self._h, obs.data_ptr(), self._stream()))
        return obs

    def reset(self, env_ids: torch.Tensor | None = None):
        if env_ids is None:
            self._check(self._lib.h1v2_reset(self._h, None, 0, self._stream()))
        else:
            ids = env_ids.to(device=self.device, dtype=torch.int64).contiguous()
            if ids.numel() == 0:  # an empty tensor has a NULL data pointer, which the C ABI reads as "all envs"
                return
            self._check(self._lib.h1v2_reset(self._h, ids.data_ptr(), ids.numel(), self._stream()))
            self._keep = ids

    def random_actions(self, step: int, out: torch.Tensor | None = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.num_envs, NJ), dtype=torch.float32, device=self.device)
        self._check(self._lib.h1v2_random_actions(self._h, out.data_ptr(), step, self._stream()))
        return out

    # ------------------------------------------------------------------
    def get_state(self, names=None) -> dict:
        out, st = {}, H1v2State()
        for name, _, ct in STATE_FIELDS:
            if names is not None and name not in names:
                continue
            cnt = state_field_count(name, self.history)
            t = torch.zeros((self.num_envs, cnt), dtype=_TORCH_DT[ct], device=self.device)
            out[name] = t
            setattr(st, name, C.cast(t.data_ptr(), C.POINTER(ct)))
        self._check(self._lib.h1v2_get_state(self._h, C.byref(st), self._stream()))
        return out

    def set_state(self, state: dict):
        st, keep = H1v2State(), []
        for name, _, ct in STATE_FIELDS:
            if name in state and name not in READ_ONLY_STATE:
                cnt = state_field_count(name, self.history)
                t = torch.as_tensor(state[name]).to(device=self.device, dtype=_TORCH_DT[ct]).reshape(self.num_envs, cnt).contiguous()
                keep.append(t)
                setattr(st, name, C.cast(t.data_ptr(), C.POINTER(ct)))
        self._check(self._lib.h1v2_set_state(self._h, C.byref(st), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # keep the staging tensors alive until consumed

    def set_reward_weights(self, weights) -> None:
        """CurriculumManager's modify_reward_weight: new weights for the kernel's reward slots, effective from the next step."""
        import numpy as np
        w = np.ascontiguousarray(weights, dtype=np.float32)
        if w.size != len(self.cfg.rew_weight):
            raise ValueError(f"expected {len(self.cfg.rew_weight)} weights, got {w.size}")
        self._check(self._lib.h1v2_set_reward_weights(self._h, w.ctypes.data_as(C.POINTER(C.c_float))))
        for i in range(w.size):
            self.cfg.rew_weight[i] = float(w[i])

    def log_host(self):
        import numpy as np
        out = np.zeros(LOG_DIM, np.float32)
        self._check(self._lib.h1v2_get_log_host(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def iter_hist(self):
        import numpy as np
        out = np.zeros(32, np.float32)
        self._check(self._lib.h1v2_debug_iter_hist(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    @property
    def launch_count(self) -> int:
        return int(self._lib.h1v2_launch_count(self._h))
