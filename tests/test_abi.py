"""The C-ABI shared library loads in the GPU-less container and exports every symbol include/mrp_b200.h declares;
without a CUDA device it must fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from gym_puzzles_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mrp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def product_lib():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return abi.load()


def test_header_and_binding_agree():
    assert _declared() == sorted(abi.EXPORTS)


def test_library_exports_every_declared_symbol(product_lib):
    for name in _declared():
        assert hasattr(product_lib.lib, name), name
    assert product_lib.backend == "cuda-sm_100a"


def test_is_built_for_sm_100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_layout_function():
    from oracle_lib import Layout  # same struct
    for variant, (n, obs, act, steps) in {0: (2, 28, 6, 2000), 1: (5, 40, 15, 3000), 2: (2, 39, 4, 2000), 3: (2, 39, 4, 2000)}.items():
        from emu_lib import emu_lib
        h = abi.Handle(variant, 1, lib=emu_lib())
        l = h.layout
        assert (l.n_agents, l.obs_dim, l.act_dim, l.max_episode_steps) == (n, obs, act, steps)
        assert l.n_fixtures == l.n_dyn_fixtures + 4 and l.max_contacts <= 32


def test_no_cpu_fallback(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(abi.MrpError, match="no CUDA device|no CPU path"):
        abi.Handle("MultiRobotPuzzle-v0", 4)
    import gym_puzzles_b200 as gp
    with pytest.raises(abi.MrpError):
        gp.VectorEnv("MultiRobotPuzzle-v0", 4)


def test_package_does_not_import_oracle_or_emu():
    pkg = os.path.join(ROOT, "gym_puzzles_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in txt.replace("tests/emu", "") or f == "mrp_b200.cu" or "oracle" not in txt, (f,)
                assert "liboracle" not in txt and "libmrp_emu" not in txt.replace("tests/emu/libmrp_emu.so", ""), f
