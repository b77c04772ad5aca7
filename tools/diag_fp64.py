"""Where does the kernel-vs-oracle velocity error come from?  One physics step (decimation 1) from identical states:
   A  fp64 build, tight solver (tol 1e-10, 100 iterations, exact line search)   -> algorithm vs oracle
   B  fp64 build, production solver settings                                     -> solver tolerances alone
   C  fp32 product, tight solver                                                 -> rounding alone
   D  fp32 product, production settings                                          -> what the parity test sees
usage: python tools/diag_fp64.py [n] [steps] [action_scale]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from h1v2_isaac_b200._capi import default_config
from h1v2_isaac_b200.backend import H1v2Sim
from oracle.oracle import Oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 96
ascale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
DEC = int(os.environ.get("DECIM", "1")); SEED = int(os.environ.get("SEED", "3"))
F64 = os.environ.get("F64LIB") or os.path.join(ROOT, "h1v2_isaac_b200", "libh1v2_b200_f64.so")
PHYS = ["root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel"]
SYNC = PHYS + ["last_action", "target_hist", "lag", "fresh", "command", "heading_target", "time_left", "is_standing", "is_heading", "cmd_metrics",
               "feet_timers", "episode_sums", "obs_history", "friction", "mass_add", "push_time_left"]
def tight(c):
    c.solver_tolerance = 1e-10; c.solver_iterations = 100; c.solver_ls_tolerance = 0.01; c.reserved[1] = 50; c.solver_step_tolerance = 0.0
    return c
def run(label, lib, tight_solver, extra=None):
    c = default_config(); c.decimation = DEC; c.max_delay = min(c.max_delay, 2 * DEC)
    if tight_solver: tight(c)
    if extra: extra(c)
    sim = H1v2Sim(n, c, device="cuda:0", seed=SEED, diagnostics=True, lib_path=lib)
    co = default_config(); co.decimation = DEC; co.max_delay = min(co.max_delay, 2 * DEC)
    orc = Oracle(co, n, seed=SEED, threads=16)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(SEED)
    errs = {k: [] for k in PHYS}
    for step in range(steps):
        a = (ascale * rng.normal(size=(n, 12))).astype(np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        _, _, to, uo = orc.step(a)
        g = {k: v.cpu().numpy() for k, v in sim.get_state(SYNC).items()}
        o = orc.get_state(PHYS)
        mc, ml = orc.activation_margin()
        keep = ~(to | uo | tg.cpu().numpy()) & (mc > 2e-6) & (ml > 2e-6)
        for k in PHYS:
            errs[k].append(np.abs(g[k][keep] - o[k][keep]).max(axis=1))
        orc.set_state({k: g[k] for k in SYNC}); orc.episode_length = sim.episode_length_buf.cpu().numpy()
    e = {k: np.concatenate(v) for k, v in errs.items()}
    lg = sim.log_host()
    label = label + f" [iters/substep {lg[30] / (DEC * n):.2f} max {lg[28]:.0f} cap {lg[29]:.0f}]"
    print(f"{label:52s} env-steps {len(e['joint_vel'])}  " + "  ".join(f"{k}: max {e[k].max():.2e} q999 {np.quantile(e[k], 0.999):.2e} q99 {np.quantile(e[k], 0.99):.2e}" for k in ("joint_vel", "root_ang_vel", "root_lin_vel", "joint_pos")), flush=True)
    sim.close()
if __name__ == "__main__":
    which = sys.argv[4] if len(sys.argv) > 4 else "ABCD"
    if "A" in which: run("A fp64 kernel, tight solver", F64, True)
    if "B" in which: run("B fp64 kernel, production solver settings", F64, False)
    if "C" in which: run("C fp32 kernel, tight solver", None, True)
    if "D" in which: run("D fp32 kernel, production solver settings", None, False)
    if "E" in which:
        run("E fp32, tol 1e-6", None, False, lambda c: setattr(c, "solver_tolerance", 1e-6))
        run("E fp32, ls_tol 0.01, ls_max 12", None, False, lambda c: (setattr(c, "solver_ls_tolerance", 0.01), c.reserved.__setitem__(1, 12)))
        run("E fp32, iterations 30", None, False, lambda c: setattr(c, "solver_iterations", 30))
    if "V" in which:
        for vt in (5e-4, 3e-4, 2e-4, 1e-4, 5e-5):
            run(f"V fp32, vel_tol {vt:g}", None, False, lambda c: setattr(c, "solver_vel_tolerance", vt))
        run("V fp32, tol 3e-6", None, False, lambda c: setattr(c, "solver_tolerance", 3e-6))
        run("V fp64, vel_tol 2e-4", F64, False, lambda c: setattr(c, "solver_vel_tolerance", 2e-4))
