"""Cap hits / iteration histogram of the mirror-lane instantiation against the plain one (cfg.reserved[3] = 1), same seeds."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config, LOG_CAP_HITS, LOG_MAX_ITERS, LOG_NAN_RESETS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for plain in (1, 0):
    cfg = default_config(); cfg.reserved[3] = plain; cfg.reserved[2] = 8
    sim = H1v2Sim(n, cfg, seed=1, diagnostics=True); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    caps = 0.0; mx = []
    for i in range(300):
        sim.step(acts[i % 8]); lg = sim.log_host(); caps += lg[LOG_CAP_HITS]; mx.append(lg[LOG_MAX_ITERS])
    h = sim.iter_hist(); h = h / h.sum()
    print(f"{'plain ' if plain else 'mirror'}: cap hits per 300 steps {caps:.0f}, max-iteration log: mean {np.mean(mx):.2f}, share of steps at the cap {np.mean(np.array(mx) >= 12):.2f}, nan resets {lg[LOG_NAN_RESETS]:.0f}")
    print("   hist", h[:14].round(5))
    sim.close()
