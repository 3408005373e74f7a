// Host build of the thread-per-agent kernels' Agent.take_action (csrc/cav_move.cuh::tpa_move + cav_math.cuh): what phase
// 2 of k_policy_move_tpa / k_policy_move_tpa_n executes per agent, here as a plain loop over a host SoA state.
// TEST INFRASTRUCTURE (tests/test_move_native_cpu.py compares the result with the oracle's dynamics step bit for bit).
#include "cav_move.cuh"

using namespace cav;

namespace {

template <typename T>
int run(const cav_config_t *cfg, const cav_buffers_t *bufs, double dt) {
    KParams<T> p{};
    p.E = cfg->num_envs;
    p.N = cfg->num_agents;
    p.dt = (T)cfg->dt;
    p.near_goal_thr = (T)cfg->near_goal_threshold;
    p.store_history = cfg->store_history;
    p.history_len = cfg->history_len;
    TpaBuffers b{};
    b.pv = bufs->pv;
    b.goal = bufs->goal;
    b.ht = bufs->ht;
    b.ego = bufs->ego;
    b.flags = bufs->flags;
    b.kinds = bufs->kinds;
    b.step_num = bufs->step_num;
    b.num_active = bufs->num_active;
    b.actions = bufs->actions;
    b.history = bufs->history;
    b.aux0 = bufs->aux0;
    b.past_actions = bufs->past_actions;
    b.past_vel = bufs->past_vel;
    const R4<T> *g_pv = reinterpret_cast<const R4<T> *>(b.pv);
    const R4<T> *g_goal = reinterpret_cast<const R4<T> *>(b.goal);
    const R4<T> *g_ht = reinterpret_cast<const R4<T> *>(b.ht);
    if (!(dt > 0.0)) dt = cfg->dt;
    for (int e = 0; e < p.E; ++e) {
        const int n = b.num_active ? b.num_active[e] : p.N;
        for (int a = 0; a < n; ++a) {
            const size_t gi = (size_t)e * p.N + a;
            tpa_move<T>(b, p, gi, ld4(g_pv + gi), ld4(g_goal + gi), ld4(g_ht + gi), b.flags[gi], b.kinds[gi],
                        make_float2(0.f, 0.f), false, (T)dt);
        }
    }
    return 0;
}

}  // namespace

extern "C" int host_move(const cav_config_t *cfg, const cav_buffers_t *bufs, double dt) {
    return cfg->precision == CAV_F64 ? run<double>(cfg, bufs, dt) : run<float>(cfg, bufs, dt);
}
