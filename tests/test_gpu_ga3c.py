"""-m gpu: the GA3C-CADRL network on the device (cav_ga3c.cu) against the golden vectors of the executed reference graph
(tests/golden/ga3c/ga3c_iros18.npz, tools/make_ga3c_golden.py) and against the oracle's restatement inside the step.

Tolerance: 1e-5 on the softmax (float32; the kernel uses FMAs and its own summation blocking), identical action wherever
the two best probabilities are more than 1e-4 apart; everything downstream of the action at the usual 1e-9."""
import os

import numpy as np
import pytest
import torch

import cav_oracle
import common
from cav_oracle import Oracle
from gym_collision_avoidance_b200 import abi
from gym_collision_avoidance_b200.envs.config import make_config, to_cav_config
from gym_collision_avoidance_b200.envs.policies.GA3C_CADRL import network
from gym_collision_avoidance_b200.envs.scenarios import random_scenario_bank
from gym_collision_avoidance_b200.envs.vec_env import VecCollisionAvoidanceEnv, kinds_array
from gym_collision_avoidance_b200.state import HostBank, HostState

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "ga3c", "ga3c_iros18.npz"))


def check_probs(got_p, got_a, want_p, tol=1e-5):
    assert np.abs(got_p - want_p).max() <= tol
    top2 = np.sort(want_p, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-4
    assert np.array_equal(got_a[clear], want_p.argmax(1)[clear])


@pytest.mark.parametrize("rows", [276, 1, 64, 65, 200])
def test_forward_against_the_executed_reference_graph(golden, rows):
    L = abi.load_library()
    blob = torch.from_numpy(golden["blob"]).cuda()
    X = torch.from_numpy(np.ascontiguousarray(golden["X"][:rows])).cuda()
    probs = torch.full((rows, 11), -1.0, dtype=torch.float32, device="cuda")
    act = torch.full((rows,), -1, dtype=torch.int32, device="cuda")
    rc = L.cav_ga3c_forward(blob.data_ptr(), X.data_ptr(), rows, probs.data_ptr(), act.data_ptr(), None)
    assert rc == 0, L.cav_last_error(None)
    torch.cuda.synchronize()
    check_probs(probs.cpu().numpy(), act.cpu().numpy(), golden["probs"][:rows])


def test_forward_large_batch_matches_oracle(golden):
    """20000 rows (313 CTAs, the last one ragged) of perturbed fixture rows against the C restatement."""
    rng = np.random.RandomState(5)
    idx = rng.randint(0, len(golden["X"]), 20000)
    X = golden["X"][idx].copy()
    X[:, 1:] += rng.uniform(-0.3, 0.3, (len(X), 137)).astype(np.float32)
    want_p, _ = cav_oracle.ga3c_forward(golden["blob"], X)
    net = network.NetworkVP_rnn("cuda:0", "network", 11)
    net.set_blob(golden["blob"])
    got = net.predict_p(X)
    check_probs(got, got.argmax(1), want_p)


def test_predict_p_crops_and_pads_like_the_reference(golden):
    net = network.NetworkVP_rnn("/cpu:0", "network", 11)  # the reference passes a TF device string
    net.set_blob(golden["blob"])
    X = golden["X"][:50]
    few = X[X[:, 0] <= 3][:, :5 + 7 * 3]          # a 3-neighbour observation: zero-padded to 138
    assert np.abs(net.predict_p(few) - golden["probs"][:50][X[:, 0] <= 3]).max() <= 1e-5
    wide = np.concatenate([X, np.ones((len(X), 30), np.float32)], 1)  # wider than the network input: cut
    assert np.abs(net.predict_p(wide) - golden["probs"][:50]).max() <= 1e-5


def _ga3c_case(N, prec, obs_f32, policies, E=48, seed=31):
    c = to_cav_config(make_config(N, evaluate=True), E, N, precision=prec, track_aux=True, obs_f32=obs_f32)
    rows = random_scenario_bank(E, N, seed=seed + N, side_length=2.5 + 0.45 * N ** 0.5)
    return c, HostBank(rows, kinds=kinds_array(E, N, policies))


@pytest.mark.parametrize("N,prec,obs_f32", [(4, "f64", False), (10, "f64", True), (10, "f32", False), (21, "f64", True)])
def test_k_ga3c_matches_the_oracle_policy(golden, N, prec, obs_f32):
    """Mid-episode states: softmax rows and the written actions of every live GA3C agent against the oracle (N = 21: the
    observation holds 20 neighbours, the network reads the first 19)."""
    from gpu_common import GpuRunner
    pol = ["GA3C_CADRL" if a % 3 != 1 else "RVO" for a in range(N)]
    c, bank = _ga3c_case(N, prec, obs_f32, pol)
    cav_oracle.set_ga3c(golden["blob"])
    try:
        hs = HostState(c)
        orc = Oracle(c)
        orc.reset(hs, bank)
        r = GpuRunner(c)
        r.cav.bind_ga3c(golden["blob"])
        for s in range(14):
            r.upload(hs)
            probs = torch.full((c.num_envs * N, 11), -1.0, dtype=torch.float32, device="cuda")
            r.cav.k_ga3c(probs)
            got = r.download()
            before = hs.actions.copy()
            orc.policy(hs)
            kind = hs.kinds.reshape(-1) & 0xff
            live = (kind == abi.CAV_POLICY_GA3C) & ((hs.flags.reshape(-1) & abi.CAV_FLAG_DONE) == 0)
            p = probs.cpu().numpy()
            assert (p[~live] == -1).all() and live.sum() > 0
            # oracle probabilities of the same rows
            src = cav_oracle.ga3c_gather(c)
            obs = hs.obs.reshape(c.num_envs * N, -1).astype(np.float32)
            X = np.where(src[None, :] >= 0, obs[:, np.maximum(src, 0)], 0.0).astype(np.float32)
            want_p, want_a = cav_oracle.ga3c_forward(golden["blob"], X[live])
            check_probs(p[live], p[live].argmax(1), want_p)
            top2 = np.sort(want_p, axis=1)[:, -2:]
            clear = np.zeros(len(live), bool)
            clear[np.nonzero(live)[0]] = (top2[:, 1] - top2[:, 0]) > 1e-4
            ga = got.actions.reshape(-1, 2)
            assert np.array_equal(ga[clear], hs.actions.reshape(-1, 2)[clear]), "float32 action pairs are exact"
            assert np.array_equal(ga[~live], before.reshape(-1, 2)[~live]), "other agents' actions untouched"
            orc.dynamics(hs)
            orc.pairwise_obs(hs)
    finally:
        cav_oracle.set_ga3c(None)


@pytest.mark.parametrize("E,N,prec,obs_f32", [(96, 6, "f64", False), (4200, 10, "f64", True), (64, 12, "f32", False)])
def test_step_with_ga3c_agents_matches_oracle(golden, E, N, prec, obs_f32):
    """Teacher-forced cav_step (the network runs inside it, ahead of policy + move; E = 4200 takes the split step) against
    the oracle step, GA3C agents mixed with RVO / non-cooperative ones.  Environments whose GA3C decision sits on a near-tie
    (two best probabilities within 1e-4) are excluded from the comparison of that step."""
    from gpu_common import GpuRunner
    pol = [("GA3C_CADRL", "RVO", "GA3C_CADRL", "noncoop")[a % 4] for a in range(N)]
    c, bank = _ga3c_case(N, prec, obs_f32, pol, E=E, seed=77)
    cav_oracle.set_ga3c(golden["blob"])
    try:
        hs = HostState(c)
        orc = Oracle(c, threads=0)
        orc.reset(hs, bank)
        r = GpuRunner(c)
        r.cav.bind_ga3c(golden["blob"])
        src = cav_oracle.ga3c_gather(c)
        steps = 10 if E < 1000 else 4
        tol = 1e-9 if prec == "f64" else 2e-4
        compared = 0
        for s in range(steps):
            r.upload(hs)
            obs = hs.obs.reshape(E * N, -1).astype(np.float32)
            X = np.where(src[None, :] >= 0, obs[:, np.maximum(src, 0)], 0.0).astype(np.float32)
            p, _ = cav_oracle.ga3c_forward(golden["blob"], X)
            top2 = np.sort(p, axis=1)[:, -2:]
            kind = hs.kinds.reshape(-1) & 0xff
            tie = ((top2[:, 1] - top2[:, 0]) <= 1e-4) & (kind == abi.CAV_POLICY_GA3C)
            ok_env = ~tie.reshape(E, N).any(1)
            orc.step(hs)
            r.cav.step()
            got = r.download()
            sel = np.nonzero(ok_env)[0]
            assert len(sel) > 0.9 * E
            for k in ("pv", "ht", "ego", "actions", "flags", "done", "game_over"):
                a, b = getattr(got, k)[sel], getattr(hs, k)[sel]
                if a.dtype.kind == "f":
                    assert np.allclose(a, b, rtol=tol, atol=tol), k
                else:
                    assert np.array_equal(a, b), k
            compared += len(sel)
        assert compared > 0
    finally:
        cav_oracle.set_ga3c(None)


def test_vec_env_with_ga3c_policy(golden):
    """VecCollisionAvoidanceEnv(policies='GA3C_CADRL', ga3c=blob): CUDA-graph step with the network inside; the agents get
    to their goals (the two-agent swap of the reference's example.py), and the policy without a network is refused."""
    with pytest.raises(abi.CavoidError):
        VecCollisionAvoidanceEnv(8, make_config(2, evaluate=True), 2, policies="GA3C_CADRL")
    rows = np.array([[[-3.0, 0.0, 3.0, 0.0, 1.0, 0.5, 0.0], [3.0, 0.0, -3.0, 0.0, 1.0, 0.5, np.pi]]])
    env = VecCollisionAvoidanceEnv(64, make_config(2, evaluate=True), 2, policies="GA3C_CADRL", ga3c=golden["blob"],
                                   scenario_bank=rows, auto_reset=False, obs_f32=True)
    env.reset()
    for _ in range(150):
        env.step()
    torch.cuda.synchronize()
    f = env.state.flags.cpu().numpy()
    assert ((f & abi.CAV_FLAG_AT_GOAL) != 0).all()
    assert ((f & abi.CAV_FLAG_IN_COLLISION) == 0).all()
    env.close()


def test_facade_with_ga3c_policy(golden):
    """The reference's example: CollisionAvoidanceEnv with GA3CCADRLPolicy agents (weights set from the fixture instead of
    initialize_network, whose checkpoint files are not on the GPU box), evaluation config (game over = all agents done)."""
    import gym_collision_avoidance_b200.envs as envs_pkg
    old = envs_pkg.Config
    envs_pkg.set_config(make_config(2, evaluate=True))
    try:
        from gym_collision_avoidance_b200.envs import test_cases as tc
        from gym_collision_avoidance_b200.envs.collision_avoidance_env import CollisionAvoidanceEnv
        env = CollisionAvoidanceEnv()
        agents = tc.get_testcase_two_agents(policies=("GA3C_CADRL", "GA3C_CADRL"))
        for a in agents:
            a.policy.nn.set_blob(golden["blob"])
        env.set_agents(agents)
        env.reset()
        done, steps = False, 0
        for steps in range(200):
            _, _, done, _, _ = env.step({})
            if done:
                break
        assert done and 40 < steps < 150
        assert all(a.is_at_goal for a in env.agents) and not any(a.in_collision for a in env.agents)
    finally:
        envs_pkg.set_config(old)


def test_bind_and_phase_errors(golden):
    """cav_k_ga3c without a network, a misaligned blob, an observation without pref_speed: status codes, no crash."""
    from gpu_common import GpuRunner
    N = 3
    c = to_cav_config(make_config(N, evaluate=True), 8, N, precision="f64")
    r = GpuRunner(c)
    L = r.cav.L
    assert L.cav_k_ga3c(r.cav._h, None, None) == -4 and b"cav_bind_ga3c" in L.cav_last_error(r.cav._h)   # NOT_BOUND
    blob = torch.zeros(abi.CAV_GA3C_BLOB_FLOATS + 4, dtype=torch.float32, device="cuda")
    assert L.cav_bind_ga3c(r.cav._h, blob.data_ptr() + 4) == -3                                          # MISALIGNED
    r.cav.bind_ga3c(golden["blob"])
    assert r.cav.L.cav_workspace_bytes(r.cav._h) >= (8 * N + 1) * 4
    r.cav.bind_ga3c(None)  # detach
    assert L.cav_k_ga3c(r.cav._h, None, None) == -4
    with pytest.raises(abi.CavoidError):
        r.cav.bind_ga3c(np.zeros(17, np.float32))
    r.cav.close()
    c2 = to_cav_config(make_config(N, evaluate=True, states_in_obs=["num_other_agents", "dist_to_goal", "other_agents_states"]),
                       8, N, precision="f64")
    r2 = GpuRunner(c2)
    rc = r2.cav.L.cav_bind_ga3c(r2.cav._h, torch.from_numpy(golden["blob"]).cuda().data_ptr())
    assert rc == -1 and b"pref_speed" in r2.cav.L.cav_last_error(r2.cav._h)                              # BAD_CONFIG
    r2.cav.close()
