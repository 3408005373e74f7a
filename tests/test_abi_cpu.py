"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/cavoid.h declares; the ctypes mirror matches the header's struct sizes; config flattening and the
observation packing follow the reference's Config semantics.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from gym_collision_avoidance_b200 import abi
from gym_collision_avoidance_b200.envs import config as cfgmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cavoid.h")


@pytest.fixture(scope="module")
def libpath():
    from gym_collision_avoidance_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(libpath):
    text = open(HEADER).read()
    declared = set(re.findall(r"^(?:int|size_t|const char \*)\s*(cav_[a-z_0-9]+)\(", text, flags=re.M))
    assert declared == set(abi.EXPORTED_SYMBOLS), declared ^ set(abi.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(libpath)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.cav_abi_version.restype = ctypes.c_int
    assert lib.cav_abi_version() == abi.CAV_ABI_VERSION


def test_ctypes_mirror_matches_header_sizes(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "cavoid.h"\nint main(){printf("%zu %zu %zu\\n", '
                   'sizeof(cav_config_t), sizeof(cav_buffers_t), sizeof(cav_scenario_bank_t));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c = map(int, subprocess.check_output([str(exe)]).split())
    assert a == ctypes.sizeof(abi.CavConfig)
    assert b == ctypes.sizeof(abi.CavBuffers)
    assert c == ctypes.sizeof(abi.CavScenarioBank)


def test_sass_is_sm100a_only(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_create_rejects_bad_config_without_gpu(libpath):
    L = abi.load_library()
    h = ctypes.c_void_p()
    c = cfgmod.to_cav_config(cfgmod.make_config(4), 8, 4)
    bad = abi.CavConfig.from_buffer_copy(c)
    bad.struct_size = 12
    assert L.cav_create(ctypes.byref(bad), ctypes.byref(h)) == -1  # CAV_ERR_BAD_CONFIG
    assert b"struct_size" in L.cav_last_error(None)
    bad = abi.CavConfig.from_buffer_copy(c)
    bad.num_agents = 65
    assert L.cav_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    bad = abi.CavConfig.from_buffer_copy(c)
    bad.agent_sorting_method = 7  # not a cav_sorting_t
    assert L.cav_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    bad = abi.CavConfig.from_buffer_copy(c)
    bad.rvo_collab_coeff = -0.5  # anti-collaborative RVO needs its re-draw period
    bad.rvo_anti_collab_t = 0.0
    assert L.cav_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert L.cav_create(None, ctypes.byref(h)) == -2  # CAV_ERR_NULL_POINTER
    assert L.cav_step(None, None, 0.0, None) == -6  # CAV_ERR_BAD_HANDLE
    assert L.cav_destroy(None) == -6


def test_product_fails_loudly_without_library(tmp_path, monkeypatch):
    monkeypatch.setattr(abi, "_LIB", None)
    monkeypatch.setattr(abi, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(abi.CavoidError, match="no CPU fallback"):
        abi.load_library()


def test_config_defaults_match_reference_semantics():
    c = cfgmod.Config()
    assert (c.DT, c.MAX_TIME_RATIO, c.NEAR_GOAL_THRESHOLD) == (0.2, 2.0, 0.2)
    assert c.MAX_NUM_AGENTS_IN_ENVIRONMENT == 4 and c.MAX_NUM_OTHER_AGENTS_OBSERVED == 3
    assert c.STATES_IN_OBS == ["is_learning", "num_other_agents", "dist_to_goal", "heading_ego_frame", "pref_speed",
                               "radius", "other_agents_states"]
    e = cfgmod.EvaluateConfig()
    assert (e.DT, e.MAX_TIME_RATIO, e.EVALUATE_MODE, e.MAX_NUM_AGENTS_IN_ENVIRONMENT) == (0.1, 8.0, True, 19)
    m = cfgmod.make_config(10)
    off, total = cfgmod.obs_layout(m)
    assert total == 6 + 7 * 9 and off["other_agents_states"] == 6
    cc = cfgmod.to_cav_config(m, 32, precision="f32")
    assert (cc.precision, cc.num_envs, cc.num_agents, cc.obs_len) == (abi.CAV_F32, 32, 10, 69)
    assert cc.game_over_mode == abi.CAV_GAME_OVER_ALL_DONE and cc.rvo_max_neighbors == 10
    assert list(cc.obs_offset) == [0, 1, 2, 3, 4, 5, 6, -1, -1]
    with pytest.raises(abi.CavoidError):
        m.STATES_IN_OBS = ["dist_to_goal", "other_agents_states_encoded"]
        cfgmod.obs_layout(m)


def test_ga3c_entry_points_validate_arguments_without_gpu(libpath):
    """Argument errors of the GA3C-CADRL entry points are reported before any CUDA call: NULL blob / X, negative row count,
    misaligned blob, bad handles."""
    L = abi.load_library()
    buf = ctypes.create_string_buffer(64)
    a16 = (ctypes.addressof(buf) + 15) & ~15
    assert L.cav_ga3c_forward(None, None, 4, None, None, None) == -2      # CAV_ERR_NULL_POINTER
    assert b"NULL" in L.cav_last_error(None)
    assert L.cav_ga3c_forward(a16, None, 4, None, None, None) == -2
    assert L.cav_ga3c_forward(a16, a16, -1, None, None, None) == -1       # CAV_ERR_BAD_CONFIG
    assert L.cav_ga3c_forward(a16 + 4, a16, 1, None, None, None) == -3    # CAV_ERR_MISALIGNED
    assert L.cav_bind_ga3c(None, a16) == -6                               # CAV_ERR_BAD_HANDLE
    assert L.cav_k_ga3c(None, None, None) == -6
    assert L.cav_env_schedule(None, None, None, None) == -6
