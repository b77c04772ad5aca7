"""N>1 host logic on CPU (gloo, world_size 2): env sharding by rank and the learner's gradient all-reduce.
The simulator has no data-path collective, so what needs covering is (1) that rank r's shard IS the
slice [r*n,(r+1)*n) of one global env set -- same Philox draws as a single process with 2n envs -- and (2) that the PPO
runner keeps the replicas identical (parameter broadcast + gradient all-reduce where rsl-rl-lib 2.3.3 does them)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_world_size_2_gloo(tmp_path, cfg):
    port = _free_port()
    procs, outs = [], []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        out = tmp_path / f"r{rank}.json"
        outs.append(out)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_mp_worker.py"), str(out)], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    for p, lg in zip(procs, logs):
        assert p.returncode == 0, lg[-3000:]
    r = [json.load(open(o)) for o in outs]
    assert r[0]["world"] == r[1]["world"] == 2
    # (2) replicas stay identical: same parameters and same adaptive learning rate on both ranks
    assert abs(r[0]["param_sum"] - r[1]["param_sum"]) < 1e-6 * max(1.0, abs(r[0]["param_sum"]))
    assert abs(r[0]["param_abs"] - r[1]["param_abs"]) < 1e-6 * r[0]["param_abs"]
    assert r[0]["lr"] == r[1]["lr"]
    # (3) rollout statistics: extras["log"] entries are reduced over ranks (rank r reports r + 1; both must log 1.5)
    assert r[0]["probe"] == pytest.approx(1.5) and r[1]["probe"] == pytest.approx(1.5)
    # (4) episode statistics: only rank 0 saw an episode end, both ranks must have taken part in the reduction and log its length
    assert r[0]["mean_len"] is not None and r[0]["mean_len"] == r[1]["mean_len"]
    # (1) the two shards are the two halves of one 32-env job
    from oracle.oracle import Oracle
    whole = Oracle(cfg, 32, seed=42, threads=2)
    s = whole.get_state(["root_pos", "command"])
    for rank in range(2):
        np.testing.assert_array_equal(np.asarray(r[rank]["root_pos"], np.float32), s["root_pos"][16 * rank:16 * (rank + 1)])
        np.testing.assert_array_equal(np.asarray(r[rank]["command"], np.float32), s["command"][16 * rank:16 * (rank + 1)])
    assert not np.array_equal(np.asarray(r[0]["root_pos"]), np.asarray(r[1]["root_pos"]))
