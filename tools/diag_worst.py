"""Find the worst single-physics-step velocity errors (kernel vs oracle from identical states) and show what they look like;
then re-run the same states with single solver settings changed to see which one matters."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from h1v2_isaac_b200._capi import default_config
from h1v2_isaac_b200.backend import H1v2Sim
from oracle.oracle import Oracle
n, steps = 8192, 96
PHYS = ["root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel"]
SYNC = PHYS + ["last_action", "target_hist", "lag", "fresh", "command", "heading_target", "time_left", "is_standing", "is_heading", "cmd_metrics",
               "feet_timers", "episode_sums", "obs_history", "friction", "mass_add", "push_time_left"]
def mk(**kw):
    c = default_config(); c.decimation = 1; c.max_delay = 2
    for k, v in kw.items():
        if k == "ls_max": c.reserved[1] = v
        else: setattr(c, k, v)
    return c
variants = {"prod": {}, "step_tol0": {"solver_step_tolerance": 0.0}, "tol1e-7": {"solver_tolerance": 1e-7}, "ls": {"solver_ls_tolerance": 0.01, "ls_max": 50},
            "tol1e-7+ls": {"solver_tolerance": 1e-7, "solver_ls_tolerance": 0.01, "ls_max": 50}}
sims = {k: H1v2Sim(n, mk(**v), device="cuda:0", seed=3, diagnostics=True) for k, v in variants.items()}
orc = Oracle(mk(), n, seed=3, threads=16)
for s in sims.values(): s.observe()
orc.observe()
rng = np.random.default_rng(3)
worst = []
for step in range(steps):
    a = (rng.normal(size=(n, 12))).astype(np.float32)
    at = torch.from_numpy(a).cuda()
    pre = {k: v.cpu().numpy() for k, v in sims["prod"].get_state(SYNC).items()}
    outs = {}
    for name, s in sims.items():
        if name != "prod":
            s.set_state(pre); s.episode_length_buf.copy_(sims["prod"].episode_length_buf)
        _, _, tg, ug = s.step(at)
        outs[name] = ({k: v.cpu().numpy() for k, v in s.get_state(SYNC + ["solver_iters", "slot_force"]).items()}, tg.cpu().numpy())
    _, _, to, uo = orc.step(a)
    o = orc.get_state(PHYS + ["slot_force"])
    it_o, res_o = orc.solver_stats()
    mc, ml = orc.activation_margin()
    g, tg = outs["prod"]
    keep = ~(to | uo | tg) & (mc > 2e-6) & (ml > 2e-6)
    err = np.abs(g["joint_vel"] - o["joint_vel"]); err[~keep] = 0
    for e in np.argsort(err.max(axis=1))[-3:]:
        j = int(err[e].argmax())
        worst.append((float(err[e, j]), step, int(e), j, {name: float(np.abs(outs[name][0]["joint_vel"][e] - o["joint_vel"][e]).max()) for name in sims},
                      float(np.abs(g["joint_vel"][e]).max()), g["solver_iters"][e].tolist(), int(it_o[e]), float(res_o[e]), float(mc[e]), float(ml[e]),
                      np.round(o["slot_force"][e].reshape(6, 3)[:, 2], 1).tolist(), float(pre["joint_pos"][e, j]), float(o["joint_vel"][e, j])))
    orc.set_state({k: g[k] for k in SYNC}); orc.episode_length = sims["prod"].episode_length_buf.cpu().numpy()
worst.sort(key=lambda w: -w[0])
for w in worst[:16]:
    print(f"err {w[0]:.2e} step {w[1]} env {w[2]} joint {w[3]} | by variant " + " ".join(f"{k} {v:.1e}" for k, v in w[4].items()) +
          f" | max|qd| {w[5]:.1f} gpu iters(max,sum) {w[6]} oracle iters {w[7]} resid {w[8]:.1e} margins {w[9]:.1e} {w[10]:.1e} Fz slots {w[11]} q_j {w[12]:.4f} qd_j {w[13]:.3f}")
