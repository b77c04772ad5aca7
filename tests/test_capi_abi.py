"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/h1v2_b200.h declares,
the ctypes mirrors have the C layout, and the product refuses to run without a GPU instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "h1v2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(h1v2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from h1v2_isaac_b200 import _capi
    lib = _capi.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/h1v2_b200.h but not exported"
        assert s in _capi._SYMBOLS, f"{s} has no ctypes binding"


def test_envs_per_warp_rule():
    """Host logic of the launch geometry (h1v2_envs_per_warp, no device needed): on a 148-SM B200 up to 2368 envs run 4 envs per warp (four mirror
    lanes per lane, at most one warp per scheduler), up to 5624 envs 8 per warp (two mirrors; BASELINE configs[1] = 4096), 16 above (the north
    star's 32768); the plain instantiations keep round 1's rule (smallest group within 3.5 warps per SM)."""
    from h1v2_isaac_b200 import _capi
    lib = _capi.load_library()
    f = lambda n, plain=0, sms=148: lib.h1v2_envs_per_warp(n, sms, plain)
    assert [f(n) for n in (1, 3, 256, 1024, 2048, 2368)] == [4] * 6
    assert [f(n) for n in (2369, 4096, 5120, 5624)] == [8] * 4
    assert [f(n) for n in (5625, 8192, 32768, 65536)] == [16] * 4
    assert [f(n, 1) for n in (256, 518, 519, 1024, 2048, 4096, 8192, 32768)] == [1, 1, 2, 2, 4, 8, 16, 16]
    assert f(4096, 0, 74) == 16 and f(1184, 0, 74) == 4 and f(0) == 4 and f(4096, 0, 0) == 8  # half the SMs; degenerate arguments


def test_struct_layout_matches_c(tmp_path):
    from h1v2_isaac_b200 import _capi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "h1v2_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(H1v2Config), sizeof(H1v2State), offsetof(H1v2Config, rew_weight), offsetof(H1v2Config, env_id_offset), offsetof(H1v2Config, history_length), offsetof(H1v2Config, command_class), offsetof(H1v2Config, root_link_com), offsetof(H1v2Config, reserved));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c, d, e, f, g, h = map(int, subprocess.check_output([str(exe)]).split())
    assert (_capi.H1v2Config.command_class.offset, _capi.H1v2Config.root_link_com.offset, _capi.H1v2Config.reserved.offset) == (f, g, h)
    assert C.sizeof(_capi.H1v2Config) == a and C.sizeof(_capi.H1v2State) == b
    assert _capi.H1v2Config.rew_weight.offset == c and _capi.H1v2Config.env_id_offset.offset == d
    assert _capi.H1v2Config.history_length.offset == e


def test_default_config_is_the_flat_task(cfg):
    # C12/flat_env_cfg.py:25-48, C12/rough_env_cfg.py:18-125, V/velocity_env_cfg.py:302-305, A/robots/h12.py:58-113
    assert cfg.decimation == 4 and abs(cfg.sim_dt - 0.005) < 1e-9 and cfg.history_length == 10
    assert list(cfg.joint_perm) == [0, 6, 1, 7, 2, 8, 3, 9, 4, 10, 5, 11]
    assert list(cfg.kp[:6]) == [200, 200, 200, 300, 40, 40] and list(cfg.effort_limit[:6]) == [220, 220, 220, 360, 45, 45]
    w = list(cfg.rew_weight)
    assert w[0] == -200 and w[1] == 1 and w[2] == 1 and abs(w[3] - 0.75) < 1e-7 and abs(w[8] + 2e-6) < 1e-12
    assert w[12] == 0 and w[13] == 0  # lin_vel_z_l2 and undesired_contacts are removed for H12
    assert 45 * cfg.history_length == 450


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from h1v2_isaac_b200 import _capi
    from h1v2_isaac_b200.backend import H1v2Sim
    with pytest.raises(RuntimeError):
        H1v2Sim(4, device="cuda:0")
    h = C.c_void_p()
    cfg = _capi.default_config()
    assert _capi.load_library().h1v2_create(C.byref(cfg), 4, 0, 1, C.byref(h)) != 0
    assert b"no CUDA device" in _capi.load_library().h1v2_last_error()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, include, link or call it."""
    pkg = os.path.join(ROOT, "h1v2_isaac_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|oracle/|h1v2o_|libh1v2_oracle|h1v2_oracle\.h")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), f"{f} references the oracle"


def test_model_tables_match_reference_when_present():
    ref = "/root/reference/packages/biped_assets/biped_assets/models/h12/scene/h12_12dof.xml"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted (GPU box)")
    subprocess.check_call(["python", os.path.join(ROOT, "tools", "compile_model.py"), "--check"])
