"""Scenario builders returning ``list[Agent]`` with the reference's names and registries
(reference ``envs/test_cases.py:68-96,144-175,212-253,495-590,900-911``)."""
import os
import pickle

import numpy as np

from .agent import Agent
from .dynamics.ExternalDynamics import ExternalDynamics
from .dynamics.UnicycleDynamics import UnicycleDynamics
from .dynamics.UnicycleDynamicsMaxTurnRate import UnicycleDynamicsMaxTurnRate
from .policies.ExternalPolicy import ExternalPolicy
from .policies.GA3CCADRLPolicy import GA3CCADRLPolicy
from .policies.LearningPolicy import LearningPolicy
from .policies.NonCooperativePolicy import NonCooperativePolicy
from .policies.RVOPolicy import RVOPolicy
from .policies.StaticPolicy import StaticPolicy
from .scenarios import reference_test_case
from .sensors.LaserScanSensor import LaserScanSensor
from .sensors.OtherAgentsStatesSensor import OtherAgentsStatesSensor

policy_dict = {"RVO": RVOPolicy, "noncoop": NonCooperativePolicy, "external": ExternalPolicy,
               "learning": LearningPolicy, "static": StaticPolicy, "GA3C_CADRL": GA3CCADRLPolicy}
sensor_dict = {"other_agents_states": OtherAgentsStatesSensor, "laserscan": LaserScanSensor}
dynamics_dict = {"unicycle": UnicycleDynamics, "external": ExternalDynamics,
                 "unicycle_max_turn_rate": UnicycleDynamicsMaxTurnRate}


def _config():
    from . import Config
    return Config


def get_testcase_two_agents(policies=("learning", "RVO")):
    """Reference ``test_cases.py:144-175``: head-on swap (-3,-3)<->(3,3), r 0.5, pref speed 1."""
    gx = gy = 3
    return [Agent(-gx, -gy, gx, gy, 0.5, 1.0, 0.0, policy_dict[policies[0]], UnicycleDynamics,
                  [OtherAgentsStatesSensor], 0),
            Agent(gx, gy, -gx, -gy, 0.5, 1.0, np.pi, policy_dict[policies[1]], UnicycleDynamics,
                  [OtherAgentsStatesSensor], 1)]


def cadrl_test_case_to_agents(test_case, policies="RVO", policy_distr=None, agents_dynamics="unicycle",
                              agents_sensors=("other_agents_states",), policy_to_ensure=None, prev_agents=None):
    """Reference ``test_cases.py:495-590``: rows ``[px,py,gx,gy,pref_speed,radius]`` -> Agents."""
    test_case = np.asarray(test_case)
    n = test_case.shape[0]
    if isinstance(policies, str):
        plist = [policies] * n
    elif policy_distr is None:
        plist = list(policies)
    else:
        plist = list(np.random.choice(policies, n, p=policy_distr))
        if policy_to_ensure is not None and policy_to_ensure not in plist:
            plist[np.random.randint(n)] = policy_to_ensure
    C = _config()
    agents = []
    for i, row in enumerate(test_case):
        px, py, gx, gy, pref_speed, radius = row[:6]
        if C.EVALUATE_MODE:
            heading = np.arctan2(gy - py, gx - px)
        else:
            heading = np.random.uniform(-np.pi, np.pi)
        if prev_agents is not None and plist[i] == prev_agents[i].policy.str:
            prev_agents[i].reset(px=px, py=py, gx=gx, gy=gy, pref_speed=pref_speed, radius=radius, heading=heading)
            agents.append(prev_agents[i])
        else:
            agents.append(Agent(px, py, gx, gy, radius, pref_speed, heading, policy_dict[plist[i]],
                                dynamics_dict[agents_dynamics], [sensor_dict[s] for s in agents_sensors], i))
    return agents


def generate_rand_test_case_multi(num_agents, side_length, speed_bnds, radius_bnds):
    """Reference ``gen_rand_testcases.py:111-142`` (15 % swap / 15 % circle / 70 % random) on NumPy's global stream:
    after ``np.random.seed(s)`` this returns the reference's rows bit for bit (tests/test_scenarios_cpu.py)."""
    return reference_test_case(np.random, num_agents, side_length, speed_bnds, radius_bnds)


def get_testcase_random(num_agents=None, side_length=4, speed_bnds=(0.5, 2.0), radius_bnds=(0.2, 0.8),
                        policies="learning", policy_distr=None, agents_dynamics="unicycle",
                        agents_sensors=("other_agents_states",), policy_to_ensure=None, prev_agents=None):
    """Reference ``test_cases.py:212-253``."""
    C = _config()
    if num_agents is None:
        num_agents = np.random.randint(2, C.MAX_NUM_AGENTS_IN_ENVIRONMENT + 1)
    if isinstance(side_length, list):
        for comp in side_length:
            if comp["num_agents"][0] <= num_agents < comp["num_agents"][1]:
                side_length = np.random.uniform(comp["side_length"][0], comp["side_length"][1])
    tc = generate_rand_test_case_multi(num_agents, float(side_length), list(speed_bnds), list(radius_bnds))
    return cadrl_test_case_to_agents(tc, policies=policies, policy_distr=policy_distr,
                                     agents_dynamics=agents_dynamics, agents_sensors=agents_sensors,
                                     policy_to_ensure=policy_to_ensure, prev_agents=prev_agents)


def gen_circle_test_case(num_agents, radius):
    """Reference ``test_cases.py:900-911``."""
    tc = np.zeros((num_agents, 6))
    for i in range(num_agents):
        tc[i, 4], tc[i, 5] = 1.0, 0.5
        th0 = (2 * np.pi / num_agents) * i
        th1 = th0 + np.pi
        tc[i, 0], tc[i, 1] = radius * np.cos(th0), radius * np.sin(th0)
        tc[i, 2], tc[i, 3] = radius * np.cos(th1), radius * np.sin(th1)
    return tc


_PRESETS = None


def preset_testCases(num_agents, full_test_suite=False, test_case_dir=None, num_test_cases=500):
    """Reference ``test_cases.py:593-897``.  ``full_test_suite=False``: the hard-coded small suite (rows
    ``[px, py, gx, gy, pref_speed, radius]``) for 1, 2, 3, 4, 5, 6, 10 or 20 agents, shipped as data
    (``envs/data/preset_cases.npz``, minted from the reference by ``tools/make_preset_cases.py``).
    ``full_test_suite=True``: a reference fixture file ``{N}_agents_500_cases.p`` (list of (N,6) arrays, :601-624); those
    files are the reference's data and are not bundled: pass the directory that holds them
    (``.../gym_collision_avoidance/envs/test_cases``) or set ``CAV_TEST_CASE_DIR``."""
    global _PRESETS
    if not full_test_suite and test_case_dir is None:
        if _PRESETS is None:
            z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "preset_cases.npz"))
            _PRESETS = {k: z[k] for k in z.files}
        cases = []
        while "n%d_c%d" % (num_agents, len(cases)) in _PRESETS:
            cases.append(_PRESETS["n%d_c%d" % (num_agents, len(cases))].copy())
        assert cases, "[preset_testCases] invalid num_agents %r" % (num_agents,)
        return cases
    d = test_case_dir or os.environ.get("CAV_TEST_CASE_DIR")
    if d is None:
        raise FileNotFoundError("set CAV_TEST_CASE_DIR to the reference's envs/test_cases directory")
    with open(os.path.join(d, "%d_agents_%d_cases.p" % (num_agents, num_test_cases)), "rb") as f:
        return pickle.load(f, encoding="latin1")


def full_test_suite(num_agents, test_case_index, policies="learning", agents_dynamics="unicycle", agents_sensors=(),
                    prev_agents=None, test_case_dir=None):
    """Reference ``test_cases.py:332-355``."""
    tc = preset_testCases(num_agents, full_test_suite=True, test_case_dir=test_case_dir)[test_case_index]
    return cadrl_test_case_to_agents(tc, policies=policies, agents_dynamics=agents_dynamics,
                                     agents_sensors=agents_sensors, prev_agents=prev_agents)


def small_test_suite(num_agents, test_case_index, policies="learning", agents_dynamics="unicycle",
                     agents_sensors=("other_agents_states",), vpref_constraint=False, radius_bnds=None):
    """Reference ``test_cases.py:313-329``: case ``test_case_index`` of the hard-coded small suite."""
    return cadrl_test_case_to_agents(preset_testCases(num_agents)[test_case_index], policies=policies,
                                     agents_dynamics=agents_dynamics, agents_sensors=agents_sensors)


def get_testcase_two_agents_laserscanners():
    """Reference ``test_cases.py:178-209``: the head-on swap with laserscan-only learning agents (the reference names
    ``PPOPolicy``, a learning policy that is not part of its tree; ``LearningPolicy`` is its device equivalent)."""
    g = 3
    return [Agent(-g, -g, g, g, 0.5, 1.0, 0.0, LearningPolicy, UnicycleDynamics, [LaserScanSensor], 0),
            Agent(g, g, -g, -g, 0.5, 1.0, np.pi, LearningPolicy, UnicycleDynamics, [LaserScanSensor], 1)]


# goal layouts of the letters C A D R L, six agents each (reference test_cases.py:425-492), in units of 2 m
_LETTERS = {
    "A": [[-1.5, 0.0], [1.5, 0.0], [0.75, 1.5], [-0.75, 1.5], [0.0, 1.5], [0.0, 3.0]],
    "C": [[0.0, 0.0], [-0.5, 1.0], [-0.5, 2.0], [0.0, 3.0], [1.5, 0.0], [1.5, 3.0]],
    "L": [[0.0, 0.0], [0.0, 1.0], [0.0, 2.0], [0.0, 3.0], [0.75, 0.0], [1.5, 0.0]],
    "D": [[0.0, 0.0], [0.0, 1.5], [0.0, 3.0], [1.5, 1.5], [1.2, 2.5], [1.2, 0.5]],
    "R": [[0.0, 0.0], [0.0, 1.5], [0.0, 3.0], [1.3, 2.8], [1.2, 1.7], [1.7, 0.0]],
}


def formation(agents, letter, num_agents=6):
    """Reference ``test_cases.py:425-492``: send the agents (from where they stand) to a random assignment of the
    letter's goal positions."""
    goals = 2.0 * np.array(_LETTERS[letter])
    agent_inds = np.arange(num_agents)
    np.random.shuffle(agent_inds)
    for agent in agents:
        sx, sy = agent.pos_global_frame
        gx, gy = goals[agent_inds[agent.id]]
        agent.reset(px=sx, py=sy, gx=gx, gy=gy, heading=agent.heading_global_frame)
    return agents
