"""CPU tests of the host-side mirror of the reference's plugin surface (no GPU): Agent initialisation arithmetic
against the golden reset states, registries, Map bitmap packing, observation dict <-> packed array."""
import numpy as np
import pytest

import common
from gym_collision_avoidance_b200 import abi
from gym_collision_avoidance_b200.envs import config as cfgmod
from gym_collision_avoidance_b200 import envs as envs_pkg


@pytest.fixture()
def eval_config():
    old = envs_pkg.Config
    cfg = envs_pkg.set_config(cfgmod.make_config(10, evaluate=True))
    yield cfg
    envs_pkg.set_config(old)


def test_agent_reset_matches_reference_golden(eval_config):
    from gym_collision_avoidance_b200.envs import test_cases as tc
    g = common.Golden("rvo10_c0")
    rows = np.zeros((10, 6))
    rows[:, 0:2] = g.z["st_pv"][0][:, 0:2]
    rows[:, 2:4] = g.z["st_goal"][0][:, 0:2]
    rows[:, 4] = g.z["st_goal"][0][:, 3]
    rows[:, 5] = g.z["st_goal"][0][:, 2]
    agents = tc.cadrl_test_case_to_agents(rows, policies="RVO")
    for i, a in enumerate(agents):
        assert a.policy.str == "RVO" and type(a.policy).kernel_id == abi.CAV_POLICY_RVO
        assert not a.policy.is_external and not a.policy.is_still_learning
        np.testing.assert_allclose(a.heading_global_frame, g.z["st_ht"][0][i, 0], rtol=0, atol=1e-15)
        np.testing.assert_allclose(a.time_remaining_to_reach_goal, g.z["st_ht"][0][i, 1], rtol=1e-14)
        np.testing.assert_allclose([a.dist_to_goal, a.heading_ego_frame], g.z["st_ego"][0][i, :2], rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(a.ref_prll, g.z["st_ego"][0][i, 2:], rtol=1e-14)
        assert a.global_state_history.shape == (int(1.2 * a.time_remaining_to_reach_goal / 0.1), 11)
        assert (a.t, a.step_num, a.is_done, a.is_at_goal) == (0.0, 0, False, False)


def test_registries_and_plugin_flags(eval_config):
    from gym_collision_avoidance_b200.envs import test_cases as tc
    assert set(tc.policy_dict) >= {"RVO", "noncoop", "static", "external", "learning"}
    assert set(tc.sensor_dict) == {"other_agents_states", "laserscan"}
    assert set(tc.dynamics_dict) >= {"unicycle", "external"}
    lp = tc.policy_dict["learning"]()
    assert lp.is_external and lp.is_still_learning and lp.str == "learning"

    class A(object):
        max_heading_change, pref_speed = np.pi / 3, 1.5
    np.testing.assert_allclose(lp.external_action_to_action(A(), [0.5, 0.75]), [0.75, np.pi / 6])
    ep = tc.policy_dict["external"]()
    assert ep.is_external and not ep.is_still_learning and ep.find_next_action(None, None, 0) is None
    a, b = tc.get_testcase_two_agents(["RVO", "noncoop"])
    assert (a.pos_global_frame == [-3, -3]).all() and (b.goal_global_frame == [-3, -3]).all()
    assert b.heading_global_frame == np.pi and isinstance(b.heading_global_frame, np.float64)
    s = tc.sensor_dict["other_agents_states"]()
    s.set_args({"max_num_other_agents_observed": 3})
    assert s.max_num_other_agents_observed == 3 and s.name == "other_agents_states"
    c = tc.gen_circle_test_case(8, 5.0)
    np.testing.assert_allclose(c[:, 0:2], -c[:, 2:4], atol=1e-12)


def test_map_bits_and_obs_packing():
    from gym_collision_avoidance_b200.envs.Map import Map
    from gym_collision_avoidance_b200.envs.wrappers import dict_to_packed, packed_to_dict
    sm = np.zeros((160, 160), bool)
    sm[3, 0] = sm[3, 31] = sm[3, 32] = sm[159, 159] = True
    m = Map(16, 16, 0.1, static_map=sm)
    bits = m.bits().reshape(160, 5)
    assert bits[3, 0] == (1 | (1 << 31)) and bits[3, 1] == 1 and bits[159, 4] == (1 << 31) and bits.sum() == bits[3].sum() + bits[159, 4]
    idx, inside = m.world_coordinates_to_map_indices([0.05, 0.05])
    assert inside and tuple(idx) == (79, 80)
    cfg = cfgmod.make_config(4)
    packed = np.arange(4 * 27, dtype=np.float64).reshape(4, 27)
    d = packed_to_dict(cfg, packed)
    assert d[2]["other_agents_states"].shape == (3, 7) and d[1]["dist_to_goal"] == packed[1, 2]
    np.testing.assert_array_equal(dict_to_packed(cfg, d), packed)


def test_config_singleton_env_vars(tmp_path, monkeypatch):
    import importlib
    p = tmp_path / "mycfg.py"
    p.write_text("from gym_collision_avoidance_b200.envs.config import EvaluateConfig\n"
                 "class Mine(EvaluateConfig):\n    def __init__(self):\n        self.MAX_NUM_AGENTS_IN_ENVIRONMENT = 7\n"
                 "        EvaluateConfig.__init__(self)\n        self.DT = 0.05\n")
    monkeypatch.setenv("GYM_CONFIG_PATH", str(p))
    monkeypatch.setenv("GYM_CONFIG_CLASS", "Mine")
    mod = importlib.reload(envs_pkg)
    try:
        assert mod.Config.MAX_NUM_AGENTS_IN_ENVIRONMENT == 7 and mod.Config.DT == 0.05
        assert mod.Config.MAX_NUM_OTHER_AGENTS_OBSERVED == 6 and mod.Config.EVALUATE_MODE
    finally:
        monkeypatch.delenv("GYM_CONFIG_PATH")
        monkeypatch.delenv("GYM_CONFIG_CLASS")
        importlib.reload(envs_pkg)


def test_image_map_loading_matches_the_reference_recipe(tmp_path):
    """Map(map_filename=png): nearest-neighbour resize to (rows, cols), occupied = not pure white (reference Map.py:17-21)."""
    from PIL import Image
    from gym_collision_avoidance_b200.envs.Map import Map
    # 80 x 80 source, one black block and one grey pixel -> 160 x 160: every source pixel becomes a 2 x 2 block
    src = np.full((80, 80), 255, np.uint8)
    src[10:20, 30:50] = 0
    src[70, 5] = 128
    f = tmp_path / "m.png"
    Image.fromarray(src).save(str(f))
    m = Map(16, 16, 0.1, map_filename=str(f))
    want = np.kron(src != 255, np.ones((2, 2), bool))
    assert m.static_map.shape == (160, 160) and m.static_map.dtype == bool and np.array_equal(m.static_map, want)
    # non-integer scale (500 -> 160): output pixel (r, c) samples source pixel floor((r + 0.5) * 500 / 160)
    big = np.full((500, 500), 255, np.uint8)
    rng = np.random.RandomState(0)
    big[rng.randint(0, 500, 4000), rng.randint(0, 500, 4000)] = 0
    big[100:180, 200:320] = 0
    f2 = tmp_path / "big.png"
    Image.fromarray(big).save(str(f2))
    m2 = Map(16, 16, 0.1, map_filename=str(f2))
    idx = np.floor((np.arange(160) + 0.5) * 500 / 160).astype(int)
    assert np.array_equal(m2.static_map, (big != 255)[np.ix_(idx, idx)])
    # same-size image: no resize; the bitmap the library reads has the same set bits
    f3 = tmp_path / "same.png"
    Image.fromarray(np.where(want, 0, 255).astype(np.uint8)).save(str(f3))
    m3 = Map(16, 16, 0.1, map_filename=str(f3))
    assert np.array_equal(m3.static_map, want)
    bits = m3.bits().reshape(160, 5)
    r, c = np.nonzero(want)
    assert all((bits[rr, cc >> 5] >> (cc & 31)) & 1 for rr, cc in zip(r[:200], c[:200])) and \
        int(sum(bin(int(w)).count("1") for w in bits.ravel())) == int(want.sum())


def test_reference_config_subclasses():
    """The reference's own Config subclasses (config.py:193-304) exist with the fields the step path reads."""
    for name, n_agents, k_obs in (("EvaluateConfig", 19, 18), ("Example", 19, 18), ("Formations", 19, 18),
                                   ("SmallTestSuite", 19, 18), ("LargeNumAgents", 19, 39), ("FullTestSuite", 19, 19),
                                   ("CollectRegressionDataset", 4, 3)):
        c = getattr(cfgmod, name)()
        assert c.EVALUATE_MODE and not c.TRAIN_MODE and c.DT == 0.1 and c.MAX_TIME_RATIO == 8.0
        assert (c.MAX_NUM_AGENTS_IN_ENVIRONMENT, c.MAX_NUM_OTHER_AGENTS_OBSERVED) == (n_agents, k_obs)
        assert c.STATE_INFO_DICT["other_agents_states"]["size"] == (k_obs, 7)
        cc = cfgmod.to_cav_config(c, 4)
        assert cc.obs_len == 6 + 7 * k_obs and cc.game_over_mode == abi.CAV_GAME_OVER_ALL_DONE
    assert cfgmod.Formations().LETTERS == ["C", "A", "D", "R", "L"]
    assert cfgmod.CollectRegressionDataset().TEST_CASE_ARGS["policies"] == "CADRL"


def test_preset_suites_formation_and_laserscanner_case(eval_config):
    from gym_collision_avoidance_b200.envs import test_cases as tc
    counts = {n: len(tc.preset_testCases(n)) for n in (1, 2, 3, 4, 5, 6, 10, 20)}
    assert counts == {1: 2, 2: 8, 3: 9, 4: 9, 5: 2, 6: 4, 10: 1, 20: 1}
    np.testing.assert_allclose(tc.preset_testCases(2)[0], [[-3, 0, 3, 0, 1, 0.3], [3, 0, -3, 0, 1, 0.3]])
    np.testing.assert_allclose(tc.preset_testCases(20)[0], tc.gen_circle_test_case(20, 10))  # test_cases.py:884-889
    with pytest.raises(AssertionError):
        tc.preset_testCases(7)
    # (the reference serves ONE list for 3 and 4 agents -- four 3-agent cases, then five 4-agent ones, test_cases.py:716)
    assert [c.shape[0] for c in tc.preset_testCases(3)] == [3] * 4 + [4] * 5
    agents = tc.small_test_suite(3, 2, policies="RVO")
    assert len(agents) == 3 and all(a.policy.str == "RVO" for a in agents)
    np.testing.assert_allclose([a.pos_global_frame for a in agents], tc.preset_testCases(3)[2][:, 0:2])
    a, b = tc.get_testcase_two_agents_laserscanners()
    assert [s.name for s in a.sensors] == ["laserscan"] and a.policy.is_still_learning and b.heading_global_frame == np.pi
    six = tc.cadrl_test_case_to_agents(tc.preset_testCases(6)[0], policies="RVO")
    np.random.seed(3)
    tc.formation(six, "R")
    goals = sorted(map(tuple, np.round([ag.goal_global_frame for ag in six], 6).tolist()))
    assert goals == sorted(map(tuple, (2.0 * np.array(tc._LETTERS["R"])).tolist()))
    assert all(ag.step_num == 0 and not ag.is_done for ag in six)


def test_reference_wrapper_classes_over_a_dict_observation_env(eval_config):
    """MultiagentDictToMultiagentArrayWrapper / MultiagentFlattenDictWrapper / FlattenDictWrapper (wrappers.py:11-173)
    on an object with the environment's surface (observation_space {agent: {state: box}}, reset / step)."""
    from gym_collision_avoidance_b200.envs import wrappers as W
    from gym_collision_avoidance_b200.envs.collision_avoidance_env import _Space
    cfg = eval_config
    N = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
    rng = np.random.RandomState(0)

    class FakeEnv(object):
        def __init__(self):
            self.observation_space = _Space()
            for a in range(N):
                self.observation_space.spaces[a] = _Space()
                for s in cfg.STATES_IN_OBS:
                    info = cfg.STATE_INFO_DICT[s]
                    self.observation_space.spaces[a][s] = _Space(np.zeros(info["size"]), np.ones(info["size"]))
            self.action_space, self.extra = "A", 7

        def _obs(self):
            return {a: {s: rng.rand(*np.atleast_1d(cfg.STATE_INFO_DICT[s]["size"])) for s in cfg.STATES_IN_OBS}
                    for a in range(N)}

        def reset(self):
            self.last = self._obs()
            return self.last, {}

        def step(self, actions):
            self.last = self._obs()
            return self.last, np.zeros(N), False, False, {"k": 1}

    env = W.MultiagentDictToMultiagentArrayWrapper(FakeEnv(), dict_keys=cfg.STATES_IN_OBS, max_num_agents=N)
    assert env.extra == 7 and env.observation_space.shape == (N, 6 + 7 * (N - 1))
    obs, info = env.reset()
    np.testing.assert_array_equal(obs, W.dict_to_packed(cfg, env.unwrapped.last))
    obs, rew, over, trunc, info = env.step({})
    assert obs.shape == (N, 69) and info == {"k": 1} and over is False
    np.testing.assert_array_equal(W.packed_to_dict(cfg, obs)[3]["other_agents_states"],
                                  env.unwrapped.last[3]["other_agents_states"])
    flat = W.MultiagentFlattenDictWrapper(FakeEnv(), cfg.STATES_IN_OBS, N)
    o, _ = flat.reset()
    assert o.shape == (N * 69,) and flat.singleAgentObservationInds(2) == [2 * 69, 3 * 69]
    np.testing.assert_array_equal(flat.observationArrayToDict(o)[1]["dist_to_goal"], flat.unwrapped.last[1]["dist_to_goal"])
    one = W.FlattenDictWrapper(FakeEnv(), cfg.STATES_IN_OBS)
    o, _ = one.reset()
    assert o.shape == (69,) and one.keyToArrayInds("radius") == [[5, 6]]
