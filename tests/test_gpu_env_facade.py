"""-m gpu: the single-environment drop-in (CollisionAvoidanceEnv / Agent / test_cases) used exactly like the
reference's experiment scripts use theirs (set_agents -> reset -> step loop), checked against the golden episodes."""
import numpy as np
import pytest

import common
from gym_collision_avoidance_b200 import abi
from gym_collision_avoidance_b200 import envs as envs_pkg
from gym_collision_avoidance_b200.envs import config as cfgmod

pytestmark = pytest.mark.gpu


def _rows(g):
    rows = np.zeros((g.n, 6))
    rows[:, 0:2] = g.z["st_pv"][0][:g.n, 0:2]
    rows[:, 2:4] = g.z["st_goal"][0][:g.n, 0:2]
    rows[:, 4] = g.z["st_goal"][0][:g.n, 3]
    rows[:, 5] = g.z["st_goal"][0][:g.n, 2]
    return rows


@pytest.mark.parametrize("name,policy", [("rvo10_c0", "RVO"), ("noncoop10_c1", "noncoop"), ("rvo3_in8slots", "RVO")])
def test_episode_through_the_reference_api(name, policy):
    g = common.Golden(name)
    old = envs_pkg.Config
    cfg = envs_pkg.set_config(g.py_config())
    try:
        from gym_collision_avoidance_b200.envs import test_cases as tc
        from gym_collision_avoidance_b200.envs.collision_avoidance_env import CollisionAvoidanceEnv
        env = CollisionAvoidanceEnv()
        agents = tc.cadrl_test_case_to_agents(_rows(g), policies=policy)
        env.set_agents(agents)
        obs, info = env.reset()
        assert info == {} and set(obs.keys()) == set(range(g.N))
        assert set(obs[0].keys()) == set(cfg.STATES_IN_OBS)
        want0 = common.Golden(name).z["obs"][0]
        from gym_collision_avoidance_b200.envs.wrappers import dict_to_packed
        common.compare_obs(g, dict_to_packed(cfg, obs)[None], want0[None], 1e-9, 1e-9)
        steps = 0
        while True:
            obs, rewards, game_over, truncated, info = env.step({})
            t = steps
            steps += 1
            assert truncated is False and rewards.shape == (g.n,)
            common.assert_close("rewards", rewards, g.z["rewards"][t][:g.n], 1e-6, 1e-6)
            assert [info["which_agents_done"][i] for i in range(g.n)] == list(g.z["dones"][t][:g.n].astype(bool))
            assert info["which_agents_learning"] == {i: False for i in range(g.n)}
            pos = np.array([a.pos_global_frame for a in env.agents])
            common.assert_close("pos", pos, g.z["st_pv"][t + 1][:g.n, 0:2], 1e-6, 1e-6)
            common.compare_obs(g, dict_to_packed(cfg, obs)[None], g.z["obs"][t + 1][None], 1e-6, 1e-6)
            assert game_over == bool(g.z["game_over"][t])
            if game_over or steps >= g.T:
                break
        assert steps == g.T
        for i, a in enumerate(env.agents):
            k = a.step_num
            assert k == int(g.z["st_step_num"][g.T][i]) and a.is_done == bool(g.z["st_flags"][g.T][i] & 0x20)
            common.assert_close("t", a.t, g.z["st_ht"][g.T][i, 2], 1e-9, 1e-9)
            common.assert_close("history", a.global_state_history[:k], g.z["history"][i, :k, :11], 1e-6, 1e-6)
        env.close()
    finally:
        envs_pkg.set_config(old)


def test_unknown_plugin_is_rejected_not_run_on_cpu():
    old = envs_pkg.Config
    envs_pkg.set_config(cfgmod.make_config(2, evaluate=True))
    try:
        from gym_collision_avoidance_b200.envs import test_cases as tc
        from gym_collision_avoidance_b200.envs.agent import Agent
        from gym_collision_avoidance_b200.envs.collision_avoidance_env import CollisionAvoidanceEnv
        from gym_collision_avoidance_b200.envs.policies.InternalPolicy import InternalPolicy

        class MyPolicy(InternalPolicy):
            def find_next_action(self, obs, agents, i):
                return np.array([1.0, 0.0])

        env = CollisionAvoidanceEnv()
        env.set_agents([Agent(0, 0, 3, 0, 0.3, 1.0, 0.0, MyPolicy, tc.dynamics_dict["unicycle"],
                              [tc.sensor_dict["other_agents_states"]], 0)])
        with pytest.raises(abi.CavoidError, match="no CPU fallback"):
            env.reset()
    finally:
        envs_pkg.set_config(old)


def test_external_actions_and_set_state():
    old = envs_pkg.Config
    cfg = envs_pkg.set_config(cfgmod.make_config(2, evaluate=False))
    try:
        from gym_collision_avoidance_b200.envs import test_cases as tc
        from gym_collision_avoidance_b200.envs.collision_avoidance_env import CollisionAvoidanceEnv
        env = CollisionAvoidanceEnv()
        env.set_agents(tc.get_testcase_two_agents(["external", "noncoop"]))
        env.reset()
        with pytest.raises(KeyError):
            env.step({})
        obs, rew, over, _, info = env.step({0: [0.5, 0.1]})
        a0 = env.agents[0]
        assert abs(a0.heading_global_frame - 0.1) < 1e-7 and abs(a0.speed_global_frame - 0.5) < 1e-7
        np.testing.assert_allclose(a0.pos_global_frame, [-3 + 0.5 * np.cos(np.float32(0.1)) * cfg.DT,
                                                         -3 + 0.5 * np.sin(np.float32(0.1)) * cfg.DT], rtol=1e-7)
        assert info["which_agents_learning"] == {0: False, 1: False}
        env.close()
    finally:
        envs_pkg.set_config(old)


def test_set_state_mid_episode_keeps_history_and_laser_rows():
    """Agent.set_state between steps (the reference's ExternalDynamics workflow, agent.py:155-190) must only move the
    agent: the trajectory rows written so far, the laserscan's 3-deep measurement history (and its 'first measurement'
    counter), and the other agents' state stay as they were; list / ndarray actions index like the reference's."""
    old = envs_pkg.Config
    cfg = envs_pkg.set_config(cfgmod.make_config(
        3, evaluate=True, use_static_map=True,
        states_in_obs=["is_learning", "num_other_agents", "dist_to_goal", "heading_ego_frame", "pref_speed", "radius",
                       "other_agents_states", "laserscan"]))
    try:
        from gym_collision_avoidance_b200.envs import test_cases as tc
        from gym_collision_avoidance_b200.envs.agent import Agent
        from gym_collision_avoidance_b200.envs.collision_avoidance_env import CollisionAvoidanceEnv
        sens = [tc.sensor_dict["other_agents_states"], tc.sensor_dict["laserscan"]]
        agents = [Agent(-2.0, 0.0, 3.0, 0.0, 0.4, 1.0, np.float64(0.0), tc.policy_dict["external"],
                        tc.dynamics_dict["external"], sens, 0),
                  Agent(2.0, 0.5, -3.0, 0.5, 0.4, 1.0, np.float64(np.pi), tc.policy_dict["noncoop"],
                        tc.dynamics_dict["unicycle"], sens, 1),
                  Agent(0.0, -2.0, 0.0, 3.0, 0.3, 0.8, np.float64(np.pi / 2), tc.policy_dict["RVO"],
                        tc.dynamics_dict["unicycle"], sens, 2)]
        env = CollisionAvoidanceEnv()
        env.set_agents(agents)
        env.reset()
        for k in range(3):
            env.step([np.array([0.0, 0.0]), None, None])  # sequence-typed actions
        hist1 = env.agents[1].global_state_history[:3].copy()
        laser1 = env.agents[1].sensor_data["laserscan"].copy()
        assert env.agents[1].step_num == 3 and np.abs(hist1[:, 1:3]).sum() > 0
        env.agents[0].set_state(-1.5, 0.2, 1.0, 0.0, 0.0)
        obs, rew, over, _, info = env.step(np.zeros((3, 2)))  # ndarray actions
        a1 = env.agents[1]
        np.testing.assert_array_equal(a1.global_state_history[:3], hist1)  # earlier rows intact
        assert a1.step_num == 4 and np.abs(a1.global_state_history[3, 1:3]).sum() > 0
        # the sensor's history shifted down by one row (LaserScanSensor.py:84-101): it was NOT re-initialised
        np.testing.assert_array_equal(a1.sensor_data["laserscan"][1:], laser1[:-1])
        np.testing.assert_allclose(env.agents[0].pos_global_frame, [-1.5, 0.2])
        env.close()
    finally:
        envs_pkg.set_config(old)


def test_user_subclass_overriding_a_device_policy_is_rejected():
    old = envs_pkg.Config
    envs_pkg.set_config(cfgmod.make_config(2, evaluate=True))
    try:
        from gym_collision_avoidance_b200.envs import test_cases as tc
        from gym_collision_avoidance_b200.envs.agent import Agent
        from gym_collision_avoidance_b200.envs.collision_avoidance_env import CollisionAvoidanceEnv

        class MyRVO(tc.policy_dict["RVO"]):
            def find_next_action(self, obs, agents, i):
                return np.array([0.1, 0.0])

        class RenamedRVO(tc.policy_dict["RVO"]):  # no behaviour change: accepted
            pass

        env = CollisionAvoidanceEnv()
        mk = lambda pol: [Agent(0, 0, 3, 0, 0.3, 1.0, np.float64(0.0), pol, tc.dynamics_dict["unicycle"],  # noqa: E731
                                [tc.sensor_dict["other_agents_states"]], 0)]
        env.set_agents(mk(MyRVO))
        with pytest.raises(abi.CavoidError, match="overrides"):
            env.reset()
        env.set_agents(mk(RenamedRVO))
        env.reset()
        env.close()
    finally:
        envs_pkg.set_config(old)
