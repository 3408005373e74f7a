"""Device-side scenario generation (cav_generate_scenarios) vs the host generator, which reproduces the reference's
generate_rand_test_case_multi bit for bit (tests/test_scenarios_cpu.py).  The device draws from Philox instead of
MT19937, so parity is statistical: hard invariants of the construction hold for every scenario, and the distribution
of every summary we can think of agrees with the host generator's within sampling error."""
import numpy as np
import pytest
import torch

from gym_collision_avoidance_b200 import abi
from gym_collision_avoidance_b200.device import Cavoid, DeviceBank
from gym_collision_avoidance_b200.envs import scenarios
from gym_collision_avoidance_b200.envs.config import make_config, to_cav_config
from gym_collision_avoidance_b200.envs.vec_env import VecCollisionAvoidanceEnv

pytestmark = pytest.mark.gpu


def _device_bank(N, S, seed):
    c = to_cav_config(make_config(N, evaluate=True), 8, N, precision="f64")
    cav = Cavoid(c, "cuda:0")
    bank = cav.generate_scenarios(DeviceBank.empty(S, N, device="cuda:0"), seed)
    torch.cuda.synchronize()
    rows, kind = bank.rows.cpu().numpy(), bank.case_kind.cpu().numpy()
    cav.close()
    return rows, kind


def _host_kind(rows):
    swap = (rows[:, 0, 1] == 0.0) & (rows[:, 1, 1] == 0.0) & (rows[:, 0, 0] == -rows[:, 1, 0])
    antipodal = np.all(np.abs(rows[:, :, 0:2] + rows[:, :, 2:4]) < 1e-9, axis=(1, 2))
    return np.where(swap, 0, np.where(antipodal, 1, 2))


def _summaries(rows, kind):
    d_sg = np.linalg.norm(rows[:, :, 2:4] - rows[:, :, 0:2], axis=-1)
    p = rows[:, :, 0:2]
    n = rows.shape[1]
    dpair = np.linalg.norm(p[:, :, None] - p[:, None], axis=-1) + np.eye(n)[None] * 1e9
    rnd = kind == 2
    return {
        "frac_swap": np.mean(kind == 0), "frac_circle": np.mean(kind == 1),
        "radius_mean": rows[..., 5].mean(), "radius_std": rows[..., 5].std(),
        "speed_mean": rows[..., 4].mean(), "speed_std": rows[..., 4].std(),
        "start_goal_mean_random": d_sg[rnd].mean(), "start_goal_std_random": d_sg[rnd].std(),
        "extent_mean_random": np.abs(rows[rnd][..., 0:4]).max(axis=(1, 2)).mean(),
        "nearest_start_mean_random": dpair[rnd].min(axis=-1).mean(),
        "ring_radius_mean_circle": np.linalg.norm(rows[kind == 1][..., 0:2], axis=-1).mean(),
        "start_goal_mean_all": d_sg.mean(),
    }


@pytest.mark.parametrize("N", [2, 4, 10])
def test_device_scenarios_hold_the_construction_invariants(N):
    rows, kind = _device_bank(N, 2048, seed=5)
    assert np.isfinite(rows).all()
    assert np.array_equal(kind, _host_kind(rows)) or N == 2  # (2 agents: a swap case is also antipodal)
    r = rows[..., 5]
    assert (r >= 0.2).all() and (r <= 0.8).all() and (rows[..., 4] >= 0.5).all() and (rows[..., 4] <= 2.0).all()
    for col in ((0, 2), (2, 4)):  # starts, then goals: r_i + r_j + GETTING_CLOSE_RANGE apart
        p = rows[:, :, col[0]:col[1]]
        d = np.linalg.norm(p[:, :, None] - p[:, None], axis=-1) + np.eye(N)[None] * 1e9
        assert (d >= r[:, :, None] + r[:, None] + scenarios.GETTING_CLOSE_RANGE - 1e-9).all()
    assert np.allclose(rows[..., 6], np.arctan2(rows[..., 3] - rows[..., 1], rows[..., 2] - rows[..., 0]), atol=1e-12)
    # random cases: every agent after the first interacts with an earlier one (no all-clear straight lines)
    checked = 0
    for s in np.nonzero(kind == 2)[0][:200]:
        tc = rows[s]
        for i in range(1, N):
            clear = all(scenarios._straight_lines_clear(tc[j, 0:2], tc[j, 2:4], tc[j, 4], tc[i, 0:2], tc[i, 2:4], tc[i, 4],
                                                        tc[j, 5] + tc[i, 5] + scenarios.GETTING_CLOSE_RANGE)
                        for j in range(i))
            assert not clear
            checked += 1
    assert checked > 100
    # different seeds / scenarios give different rows; same seed reproduces
    again, _ = _device_bank(N, 2048, seed=5)
    other, _ = _device_bank(N, 2048, seed=6)
    assert np.array_equal(again, rows) and not np.array_equal(other, rows)
    assert len({tuple(np.round(x, 9)) for x in rows[:, 0, 0:2]}) > 1500


@pytest.mark.parametrize("N", [4, 10])
def test_device_scenarios_match_host_generator_distribution(N):
    S = 6000
    drows, dkind = _device_bank(N, S, seed=11)
    hrows = scenarios.random_scenario_bank(S, N, seed=11)
    hs, ds = _summaries(hrows, _host_kind(hrows)), _summaries(drows, dkind)
    # tolerances: a few standard errors of each summary at S = 6000 (fractions: sqrt(.15*.85/6000) = 0.0046)
    tol = {"frac_swap": 0.02, "frac_circle": 0.02, "radius_mean": 0.005, "radius_std": 0.005, "speed_mean": 0.01,
           "speed_std": 0.01, "start_goal_mean_random": 0.15, "start_goal_std_random": 0.30,  # heavy tail: the side grows
           "extent_mean_random": 0.15, "nearest_start_mean_random": 0.08, "ring_radius_mean_circle": 0.08,
           "start_goal_mean_all": 0.15}
    bad = {k: (hs[k], ds[k]) for k in tol if abs(hs[k] - ds[k]) > tol[k]}
    assert not bad, bad
    # two-sample Kolmogorov-Smirnov distance on the start-goal distance of random cases
    a = np.sort(np.linalg.norm(hrows[..., 2:4] - hrows[..., 0:2], axis=-1)[_host_kind(hrows) == 2].ravel())
    b = np.sort(np.linalg.norm(drows[..., 2:4] - drows[..., 0:2], axis=-1)[dkind == 2].ravel())
    grid = np.linspace(0, max(a[-1], b[-1]), 400)
    ks = np.abs(np.searchsorted(a, grid) / len(a) - np.searchsorted(b, grid) / len(b)).max()
    assert ks < 0.03, ks


def test_vec_env_runs_on_device_generated_scenarios():
    N, E = 10, 256
    env = VecCollisionAvoidanceEnv(E, make_config(N, evaluate=True), N, precision="f64", policies="RVO",
                                   scenario_source="device", seed=3, num_scenarios=512)
    obs, _ = env.reset()
    first = env.bank.rows.clone()
    for _ in range(30):
        obs, rew, go, info = env.step()
    assert torch.isfinite(obs).all()
    env.regenerate_scenarios(4)
    torch.cuda.synchronize()
    assert not torch.equal(first, env.bank.rows)
    for _ in range(150):
        env.step()
    st = env.episode_stats()
    assert st["episodes"] > 0
    env.close()


def test_training_mode_banks_draw_random_initial_headings():
    """EVALUATE_MODE False: the reference draws heading ~ U(-pi, pi) per agent (test_cases.py:558-559) instead of pointing
    the agent at its goal; both bank sources of VecCollisionAvoidanceEnv follow (device: one more Philox draw per agent)."""
    N, E = 6, 64
    cfg = make_config(N, evaluate=False)
    for src in ("host", "device"):
        env = VecCollisionAvoidanceEnv(E, cfg, N, policies=["learning"] + ["RVO"] * (N - 1), scenario_source=src, seed=9,
                                       num_scenarios=2048)
        rows = env.bank.rows.cpu().numpy()
        hd = rows[..., 6].ravel()
        to_goal = np.arctan2(rows[..., 3] - rows[..., 1], rows[..., 2] - rows[..., 0]).ravel()
        assert (hd >= -np.pi).all() and (hd < np.pi).all()
        assert np.mean(np.abs(hd - to_goal) < 1e-9) < 0.01, src
        grid = np.linspace(-np.pi, np.pi, 200)
        ks = np.abs(np.searchsorted(np.sort(hd), grid) / hd.size - (grid + np.pi) / (2 * np.pi)).max()
        assert ks < 0.02, (src, ks)
        env.reset()
        got = env.state.ht[..., 0].cpu().numpy()
        sidx = np.arange(E) % 2048
        assert np.array_equal(got, rows[sidx, :, 6])
        env.close()
    ev = VecCollisionAvoidanceEnv(E, make_config(N, evaluate=True), N, scenario_source="device", seed=9, num_scenarios=256)
    r = ev.bank.rows.cpu().numpy()
    np.testing.assert_allclose(r[..., 6], np.arctan2(r[..., 3] - r[..., 1], r[..., 2] - r[..., 0]), atol=1e-12)
    ev.close()
