"""Generate tests/golden/terrain_curriculum.npz: the reference's OWN terrain-level curriculum
(packages/biped_tasks/biped_tasks/tasks/locomotion/velocity/mdp/curriculums.py:21-52 terrain_levels_vel), imported from /root/reference
and run on a stand-in env.  What the function calls on the terrain -- TerrainImporter.update_env_origins -- is upstream isaaclab 2.1.0 and
absent here; the stand-in restates it (levels += up - down; >= max level -> torch.randint_like; clip at 0; origins re-read).
Inputs are chosen exactly representable (positions / commands on a 1/1024 grid) so that world - origin is exact in fp32.
Run in the build container:  python tests/golden/make_rough_goldens.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "h1v2_isaac_b200", "shims"), os.path.join(REF, "packages", "biped_tasks"), os.path.join(REF, "packages", "biped_assets")]

from biped_tasks.tasks.locomotion.velocity.mdp.curriculums import terrain_levels_vel  # noqa: E402
from biped_tasks.utils.mdp.terrains import ROUGH_TERRAINS_CFG  # noqa: E402

ROWS, COLS, TILE = ROUGH_TERRAINS_CFG.num_rows, ROUGH_TERRAINS_CFG.num_cols, ROUGH_TERRAINS_CFG.size[0]
N = 4096
g = torch.Generator().manual_seed(7)


class Terrain:
    """TerrainImporter stand-in [UPSTREAM 2.1.0 terrain_importer.py: _compute_env_origins_curriculum, update_env_origins]."""

    def __init__(self):
        class _C:
            terrain_generator = ROUGH_TERRAINS_CFG
        self.cfg = _C()
        r, c = torch.meshgrid(torch.arange(ROWS), torch.arange(COLS), indexing="ij")
        self.terrain_origins = torch.stack([(r + 0.5) * TILE - ROWS * TILE / 2, (c + 0.5) * TILE - COLS * TILE / 2, torch.zeros_like(r, dtype=torch.float32)], dim=-1).float()
        self.max_terrain_level = ROWS
        self.terrain_levels = torch.randint(0, ROWS, (N,), generator=g)
        self.terrain_types = torch.div(torch.arange(N), (N / COLS), rounding_mode="floor").to(torch.long)
        self.env_origins = self.terrain_origins[self.terrain_levels, self.terrain_types].clone()
        self.seen = {}

    def update_env_origins(self, env_ids, move_up, move_down):
        self.seen = {"move_up": move_up.clone(), "move_down": move_down.clone()}
        self.terrain_levels[env_ids] += 1 * move_up - 1 * move_down
        wrapped = self.terrain_levels[env_ids] >= self.max_terrain_level
        self.seen["wrapped"] = wrapped.clone()
        self.terrain_levels[env_ids] = torch.where(wrapped, torch.randint_like(self.terrain_levels[env_ids], self.max_terrain_level),
                                                   torch.clip(self.terrain_levels[env_ids], 0))
        self.env_origins[env_ids] = self.terrain_origins[self.terrain_levels[env_ids], self.terrain_types[env_ids]]


terrain = Terrain()
levels0 = terrain.terrain_levels.clone()
rel = (torch.randint(-6 * 1024, 6 * 1024, (N, 2), generator=g).float() / 1024.0)
rel[::7] = torch.round(rel[::7] * 0.1 * 1024.0) / 1024.0  # many short walks (still on the 1/1024 grid)
rel[5::11, 0] = 4.0; rel[5::11, 1] = 0.0  # exactly half a tile: NOT further than it
cmd = torch.randint(-1024, 1024, (N, 3), generator=g).float() / 1024.0
cmd[::5, :2] = 0.0  # standing commands never move down


class _Data:
    root_pos_w = torch.cat([terrain.env_origins[:, :2] + rel, torch.ones(N, 1)], dim=1)


class _Asset:
    data = _Data()


class _Scene:
    terrain = terrain
    env_origins = terrain.env_origins

    def __getitem__(self, k):
        return _Asset()


class _Cmd:
    def get_command(self, name):
        return cmd


class Env:
    scene = _Scene()
    command_manager = _Cmd()
    max_episode_length_s = 20.0


assert torch.equal(_Data.root_pos_w[:, :2] - terrain.env_origins[:, :2], rel)  # exact in fp32 by construction
mean = terrain_levels_vel(Env(), torch.arange(N))
out = dict(rel=rel.numpy(), cmd=cmd.numpy(), levels0=levels0.numpy().astype(np.int32), types=terrain.terrain_types.numpy().astype(np.int32),
           move_up=terrain.seen["move_up"].numpy(), move_down=terrain.seen["move_down"].numpy(), wrapped=terrain.seen["wrapped"].numpy(),
           levels1=terrain.terrain_levels.numpy().astype(np.int32), mean=np.float32(mean), rows=ROWS, cols=COLS, tile=TILE,
           origins_xy=terrain.terrain_origins[..., :2].numpy())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "terrain_curriculum.npz"), **out)
print("wrote terrain_curriculum.npz: up", int(out["move_up"].sum()), "down", int(out["move_down"].sum()), "wrapped", int(out["wrapped"].sum()), "mean", float(mean))
