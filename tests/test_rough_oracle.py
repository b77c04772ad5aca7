"""Rough terrain, oracle side: the CPU restatement against what the reference and upstream pin.

* terrain grid: the generator cfg the reference ships (packages/biped_tasks/biped_tasks/utils/mdp/terrains.py:11-28) -- 10 x 20 tiles of
  8 m at 0.1 m, heights {0, 5, 10, 15, 20} mm, flat 3-cell tile rims, tile origins at the highest vertex of the central 2 m patch;
* triangulation: isaaclab's convert_height_field_to_mesh restated HERE as an explicit vertex / triangle list and interpolated
  barycentrically -- an independent second statement of the same upstream function the oracle's closed form follows;
* curriculum: tests/golden/terrain_curriculum.npz, produced by the reference's OWN terrain_levels_vel
  (tasks/locomotion/velocity/mdp/curriculums.py:21-52) on a stand-in env (tests/golden/make_rough_goldens.py);
* physics: a zero height field reproduces the plane bit for bit; on a uniform slope a standing robot is held by a contact force
  that equals its weight, with the contact normal of the slope.
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rough():
    from oracle.oracle import task_config
    return task_config("rough")


def test_generated_height_field_has_the_reference_generator_cfg_statistics(rough):
    from oracle.oracle import Oracle
    orc = Oracle(rough, 64, seed=3)
    H = orc.terrain()
    assert H.shape == (10 * 80 + 1, 20 * 80 + 1)
    lev = np.rint(H / 0.005).astype(int)
    assert np.allclose(H, lev * np.float32(0.005), atol=1e-9) and set(np.unique(lev)) == {0, 1, 2, 3, 4}
    # every tile: 81 x 81 vertices, rim of 3 flat (border_width 0.25 -> int(2.5) + 1), 75 x 75 random interior
    for r, c in ((0, 0), (3, 7), (9, 19)):
        t = lev[80 * r:80 * r + 81, 80 * c:80 * c + 81]
        assert (t[:3] == 0).all() and (t[-3:] == 0).all() and (t[:, :3] == 0).all() and (t[:, -3:] == 0).all()
        inner = t[3:-3, 3:-3]
        assert inner.shape == (75, 75)
        cnt = np.bincount(inner.ravel(), minlength=5) / inner.size
        assert np.abs(cnt - 0.2).max() < 0.03  # uniform over the five levels (75 * 75 draws: sigma 0.005)
    # independent draws: neighbouring vertices are uncorrelated
    inner = lev[3:78, 3:78].astype(float)
    assert abs(np.corrcoef(inner[:-1].ravel(), inner[1:].ravel())[0, 1]) < 0.05
    # another seed gives another field; the same seed the same
    assert not np.array_equal(Oracle(rough, 4, seed=4).terrain(), H) and np.array_equal(Oracle(rough, 4, seed=3).terrain(), H)
    # tile origin height = highest vertex of the central 2 m x 2 m patch [UPSTREAM height_field/utils.py]; queried at the tile centre
    for lvl, ty in ((0, 0), (5, 12)):
        patch = H[80 * lvl + 30:80 * lvl + 50, 80 * ty + 30:80 * ty + 50]
        h, n = orc.terrain_query(lvl, ty, 0.0, 0.0)
        assert abs((h + patch.max()) - H[80 * lvl + 40, 80 * ty + 40]) < 1e-7


def _mesh_height(H, hs, x, y):
    """convert_height_field_to_mesh [UPSTREAM isaaclab/terrains/height_field/utils.py] restated as an explicit mesh: vertices on the grid,
    two triangles per cell -- (i,j) (i+1,j+1) (i,j+1) and (i,j) (i+1,j) (i+1,j+1) -- and a vertical ray cast onto it."""
    nr, nc = H.shape
    verts = np.stack([np.repeat(np.arange(nr) * hs, nc), np.tile(np.arange(nc) * hs, nr), H.ravel()], axis=1)
    tris = []
    for i in range(nr - 1):
        ind0 = np.arange(0, nc - 1) + i * nc
        ind1, ind2 = ind0 + 1, ind0 + nc
        ind3 = ind2 + 1
        tris += [np.stack([ind0, ind3, ind1], 1), np.stack([ind0, ind2, ind3], 1)]
    tris = np.concatenate(tris)
    p = np.array([x, y])
    for a, b, c in verts[tris]:
        m = np.array([b[:2] - a[:2], c[:2] - a[:2]]).T
        w = np.linalg.solve(m, p - a[:2])
        if w[0] >= -1e-12 and w[1] >= -1e-12 and w.sum() <= 1 + 1e-12:
            n = np.cross(b - a, c - a)
            n = n / np.linalg.norm(n) * np.sign(n[2])
            return a[2] + w[0] * (b[2] - a[2]) + w[1] * (c[2] - a[2]), n
    raise AssertionError("point outside the mesh")


def test_terrain_query_is_the_upstream_triangulation(rough):
    """Heights and normals of the closed-form lookup == a ray cast onto the explicitly triangulated mesh, on both triangles of a cell."""
    from oracle.oracle import Oracle
    orc = Oracle(rough, 4, seed=9)
    H = orc.terrain().astype(np.float64)
    lvl, ty = 2, 5
    i0, j0 = 80 * lvl + 20, 80 * ty + 33  # a 4 x 4 patch inside tile (2, 5)
    patch = H[i0:i0 + 5, j0:j0 + 5]
    oz = H[80 * lvl + 30:80 * lvl + 50, 80 * ty + 30:80 * ty + 50].max()
    rng = np.random.default_rng(0)
    hs = 0.1
    for _ in range(40):
        px, py = rng.uniform(0.01, 0.39, 2)
        want_h, want_n = _mesh_height(patch, hs, px, py)
        lx, ly = (20 * hs + px) - 4.0, (33 * hs + py) - 4.0  # relative to the tile centre
        h, n = orc.terrain_query(lvl, ty, lx, ly)
        assert abs((h + oz) - want_h) < 2e-7, (px, py)  # hscale is the fp32 0.1 on the oracle's side
        assert np.abs(n - want_n).max() < 1e-6
    # outside the tile grid: the flat border at world height 0
    h, n = orc.terrain_query(0, 0, -10.0, 0.0)
    assert abs(h + H[30:50, 30:50].max()) < 1e-9 and np.array_equal(n, [0, 0, 1])


def test_terrain_curriculum_equals_the_reference_function(rough):
    """terrain_levels_vel as run by the reference's own code on 4096 stand-in envs (walked distance vs half a tile, vs half the
    commanded distance; past the last level -> a random one): the oracle's reset applies the same moves, env by env."""
    from oracle.oracle import Oracle
    G = np.load(os.path.join(ROOT, "tests", "golden", "terrain_curriculum.npz"))
    n = len(G["levels0"])
    assert int(G["rows"]) == rough.terrain_rows and int(G["cols"]) == rough.terrain_cols and float(G["tile"]) == rough.terrain_tile_size
    orc = Oracle(rough, n, seed=5)
    st = orc.get_state(["root_pos", "command", "terrain_level", "terrain_type"])
    assert np.array_equal(st["terrain_type"][:, 0], G["types"])  # floor(i / (n / cols)) in fp32, as torch.div(..., rounding_mode="floor")
    assert st["terrain_level"].min() >= 0 and st["terrain_level"].max() <= 5 and len(np.unique(st["terrain_level"])) == 6  # randint(0, max_init + 1)
    pos = st["root_pos"].copy(); pos[:, :2] = G["rel"]
    orc.set_state({"root_pos": pos, "command": G["cmd"], "terrain_level": G["levels0"]})
    orc.reset(np.arange(n))
    lv = orc.get_state(["terrain_level"])["terrain_level"][:, 0]
    w = G["wrapped"]
    assert np.array_equal(lv[~w], G["levels1"][~w])
    assert (lv[w] >= 0).all() and (lv[w] < rough.terrain_rows).all() and len(np.unique(lv[w])) > 5  # sent to a random level
    up, down = G["move_up"], G["move_down"]
    assert up.sum() > 1000 and down.sum() > 1000 and not (up & down).any()
    # positions are relative to the tile origin: after the reset the env stands on its (new) tile, within the reset offsets
    p = orc.get_state(["root_pos"])["root_pos"]
    assert np.abs(p[:, :2]).max() <= 0.5 + 1e-6 and np.allclose(p[:, 2], rough.init_root_height)


def test_zero_height_field_is_the_plane(rough):
    """The rough code path on a flat field == the plane code path of the flat ids (to float rounding): same states, same forces, height scan = z - offset."""
    from oracle.oracle import Oracle
    c = rough.copy()
    c.enable_corruption = 0
    n = 48
    a = Oracle(c, n, seed=2)
    a.set_terrain(np.zeros_like(a.terrain()))
    f = c.copy()
    f.terrain_enable = 0  # plane under the same observation layout
    b = Oracle(f, n, seed=2)
    b.set_state(a.get_state(["root_pos", "root_quat", "joint_pos", "command", "heading_target", "time_left", "is_standing", "is_heading", "lag"]))
    oa, ob = a.observe(), b.observe()
    assert oa.shape == (n, 235) and np.array_equal(oa, ob)
    z = a.get_state(["root_pos"])["root_pos"][:, 2]
    assert np.allclose(oa[:, 48:], np.clip(z - 0.5, -1, 1)[:, None], atol=1e-6)
    rng = np.random.default_rng(0)
    for k in range(30):
        act = rng.normal(size=(n, 12)).astype(np.float32)
        ra, rb = a.step(act), b.step(act)
        tol = 1e-6 if k < 5 else 1e-4  # the frame-general contact rows round differently in the last bit; stiff contacts amplify that over steps
        for x, y in zip(ra, rb):
            np.testing.assert_allclose(x, y, rtol=tol, atol=tol)
    sa, sb = a.get_state(["joint_pos", "joint_vel", "root_pos", "slot_force"]), b.get_state(["joint_pos", "joint_vel", "root_pos", "slot_force"])
    for k in sa:
        np.testing.assert_allclose(sa[k], sb[k], rtol=1e-4, atol=1e-4, err_msg=k)


def _sloped(orc, sx, sy):
    H = orc.terrain()
    x = ((np.arange(H.shape[0]) * 0.1) % 8.0) - 4.0
    y = ((np.arange(H.shape[1]) * 0.1) % 8.0) - 4.0
    return (sx * x[:, None] + sy * y[None, :]).astype(np.float32)  # every tile: the plane h = sx * x + sy * y about its centre


def test_slope_contacts_live_in_the_slope_frame(rough):
    """Uniform slopes.  (1) Every foot's contact force stays inside the friction pyramid built on the SLOPE normal and MuJoCo's tangents
    for it (|f_t1| + |f_t2| <= mu f_n, reached while a foot slides) -- about the vertical it would not.  (2) Turning terrain and robot
    together by 90 degrees about z turns the solution with them: the x and the y branch of the lookup, and both tangent rows, agree.
    (3) The height scan reads the slope."""
    from oracle.oracle import Oracle
    c = rough.copy()
    c.enable_corruption = 0
    c.terrain_curriculum = 0
    c.max_delay = 0
    for i in range(6):
        c.reset_pose_range[i][0] = c.reset_pose_range[i][1] = 0.0
    n, s, mu = 8, 0.15, float(c.friction)
    a, b = Oracle(c, n, seed=1), Oracle(c, n, seed=1)
    a.set_terrain(_sloped(a, s, 0.0)); b.set_terrain(_sloped(b, 0.0, s))
    h, nrm = a.terrain_query(1, 1, 0.3, -0.2)
    oz = _sloped(a, s, 0.0)[80 + 30:80 + 50, 0].max()
    nx = np.array([-s, 0, 1]) / np.sqrt(1 + s * s)
    assert abs(h - (0.3 * s - oz)) < 1e-6 and np.allclose(nrm, nx, atol=1e-6)
    a.reset(None); b.reset(None)
    # b = a turned by +90 degrees about z: (x, y) -> (-y, x)
    st = a.get_state(["root_pos", "root_quat", "root_lin_vel", "joint_pos", "joint_vel"])
    q = st["root_quat"].astype(np.float64)
    rz = np.array([np.sqrt(0.5), 0, 0, np.sqrt(0.5)])
    qb = np.stack([rz[0] * q[:, 0] - rz[3] * q[:, 3], rz[0] * q[:, 1] - rz[3] * q[:, 2], rz[0] * q[:, 2] + rz[3] * q[:, 1], rz[0] * q[:, 3] + rz[3] * q[:, 0]], 1)
    # the tile origin sits at the highest vertex of its central patch, which differs between the two fields by nothing: same heights
    b.set_state({"root_quat": qb, "root_pos": st["root_pos"] * [1, 1, 1], "joint_pos": st["joint_pos"], "joint_vel": st["joint_vel"]})
    t1 = np.array([0.0, 1.0, 0.0]) - nx * nx[1]
    t1 /= np.linalg.norm(t1)
    t2 = np.cross(nx, t1)
    rng = np.random.default_rng(3)
    worst, sliding = 0.0, 0
    for k in range(60):
        act = (0.5 * rng.normal(size=(n, 12))).astype(np.float32)
        oa, ra, ta, _ = a.step(act)
        ob, rb, tb, _ = b.step(act)
        if k == 0:
            scan = oa[:, 48:].reshape(n, 11, 17)
            assert np.allclose(np.diff(scan, axis=2).mean(axis=(1, 2)), -s * 0.1, atol=1e-3)  # ahead (+x, uphill) the ground is closer
            assert np.allclose(np.diff(ob[:, 48:].reshape(n, 11, 17), axis=2).mean(axis=(1, 2)), -s * 0.1, atol=1e-3)
        if ta.any() or tb.any():
            break
        sa, sb = a.get_state(["joint_pos", "joint_vel", "root_lin_vel", "slot_force"]), b.get_state(["joint_pos", "joint_vel", "root_lin_vel", "slot_force"])
        tol = 1e-6 if k < 5 else 1e-3
        np.testing.assert_allclose(sa["joint_pos"], sb["joint_pos"], rtol=0, atol=tol)
        np.testing.assert_allclose(sa["joint_vel"], sb["joint_vel"], rtol=0, atol=100 * tol)
        np.testing.assert_allclose(sb["root_lin_vel"][:, 0], -sa["root_lin_vel"][:, 1], rtol=0, atol=10 * tol)
        np.testing.assert_allclose(sb["root_lin_vel"][:, 1], sa["root_lin_vel"][:, 0], rtol=0, atol=10 * tol)
        F = sa["slot_force"].reshape(n, 6, 3)[:, :2].astype(np.float64)  # the two feet
        fn = F @ nx
        on = fn > 5.0
        ratio = (np.abs(F @ t1) + np.abs(F @ t2))[on] / (mu * fn[on])
        if on.any():
            worst = max(worst, float(ratio.max())); sliding += int((ratio > 0.98).sum())
    assert k >= 20, "the robots fell before the test saw anything"
    assert worst <= 1.0 + 1e-6 and sliding > 0, (worst, sliding)


def test_play_cfg_of_the_rough_id_runs_on_the_oracle():
    """Isaac-Velocity-Rough-H12_12dof-Play-v0 (C12/rough_env_cfg.py:128-156): 50 envs on 5 x 5 tiles, every level populated from the start
    (max_init_terrain_level None), no curriculum, no observation noise, forward command 1 m/s with heading 0."""
    import json
    from h1v2_isaac_b200._capi import H1v2Config
    from oracle.oracle import Oracle, task_config
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "play_cfg_resolved.json")))["Isaac-Velocity-Rough-H12_12dof-Play-v0"]
    c = task_config("rough")
    for k, v in gold["kernel_config"].items():  # the reference's own Play cfg, flattened
        cur = getattr(c, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                if hasattr(cur[i], "__len__"):
                    for j, y in enumerate(x):
                        cur[i][j] = y
                else:
                    cur[i] = x
        else:
            setattr(c, k, v)
    assert isinstance(c, H1v2Config) and c.terrain_rows == 5 and c.terrain_cols == 5 and c.terrain_curriculum == 0 and c.enable_corruption == 0
    n = gold["num_envs"]
    orc = Oracle(c, n, seed=1)
    assert orc.terrain().shape == (5 * 80 + 1, 5 * 80 + 1)
    st = orc.get_state(["terrain_level", "terrain_type", "command"])
    assert set(np.unique(st["terrain_level"])) == {0, 1, 2, 3, 4} and np.array_equal(st["terrain_type"][:, 0], np.arange(n) // 10)
    moving = np.abs(st["command"]).sum(axis=1) > 0  # 2 % of the envs stand (rel_standing_envs)
    assert moving.sum() >= n - 3 and np.allclose(st["command"][moving, 0], 1.0) and np.all(st["command"][:, 1] == 0)
    lv0 = st["terrain_level"].copy()
    obs = orc.observe()
    rng = np.random.default_rng(0)
    for _ in range(60):
        obs2, rew, term, trunc = orc.step(rng.normal(size=(n, 12)).astype(np.float32))
    assert term.sum() + trunc.sum() >= 0 and np.isfinite(obs2).all()
    assert np.array_equal(orc.get_state(["terrain_level"])["terrain_level"], lv0)  # resets happened (random actions), levels did not move
    # no corruption: two observations of the same state are identical
    assert np.array_equal(orc.observe()[:, :48], orc.observe()[:, :48])
