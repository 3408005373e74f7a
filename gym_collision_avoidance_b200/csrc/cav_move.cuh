// cav_move.cuh -- Agent.take_action of the thread-per-agent kernels (host/device source, see cav_math.cuh).
#pragma once
#include "cav_math.cuh"

namespace cav {

struct TpaBuffers {
    void *pv, *goal, *ht, *ego;
    uint32_t *flags;
    uint32_t *kinds;
    int32_t *step_num;
    const int32_t *num_active;
    float *actions;
    void *reward, *min_dist;
    uint8_t *done, *game_over;
    void *obs;
    void *history, *aux0, *past_actions, *past_vel;
    const int32_t *episode;
};

// Agent.take_action for one agent by one thread (agent.py:192-241: UnicycleDynamics.py:27-47 / MaxTurnRate :30-43 /
// ExternalDynamics, Dynamics.update_ego_frame :31-41, history row agent.py:257-289, goal / timeout checks).
// `act` is the action chosen this step when have_act, otherwise it is read from the actions column.
template <typename T>
CAV_HDI void tpa_move(const TpaBuffers &b, const KParams<T> &p, size_t gi, R4<T> pv, const R4<T> &gl,
                                         R4<T> ht, uint32_t f, uint32_t kind, float2 act, bool have_act, T dt) {
    R4<T> *g_pv = reinterpret_cast<R4<T> *>(b.pv);
    R4<T> *g_ht = reinterpret_cast<R4<T> *>(b.ht);
    R4<T> *g_ego = reinterpret_cast<R4<T> *>(b.ego);
    if (f & (CAV_FLAG_AT_GOAL | CAV_FLAG_RAN_OUT_OF_TIME | CAV_FLAG_IN_COLLISION)) {  // agent.py:202-209
        if (f & CAV_FLAG_AT_GOAL) f |= CAV_FLAG_WAS_AT_GOAL_ALREADY;
        if (f & CAV_FLAG_IN_COLLISION) f |= CAV_FLAG_WAS_IN_COLLISION_ALREADY;
        pv.z = (T)0;
        pv.w = (T)0;
        st4(g_pv + gi, pv);
        b.flags[gi] = f;
        if (b.past_vel) {
            R4<T> *pvel = reinterpret_cast<R4<T> *>(b.past_vel) + gi;
            R4<T> q = ld4(pvel);
            q.z = q.x;
            q.w = q.y;
            q.x = (T)0;
            q.y = (T)0;
            st4(pvel, q);
        }
    } else {
        if (!have_act) act = reinterpret_cast<const float2 *>(b.actions)[gi];
        const T a0 = (T)act.x;
        T a1 = (T)act.y;
        const uint32_t dyn = CAV_KIND_DYNAMICS(kind);
        if (b.past_actions) {  // agent.py:212-213
            R4<T> *pa = reinterpret_cast<R4<T> *>(b.past_actions) + gi;
            R4<T> q = ld4(pa);
            q.z = q.x;
            q.w = q.y;
            q.x = a0;
            q.y = a1;
            st4(pa, q);
        }
        T speed_out = (T)0, dhead_out = (T)0;
        if (b.aux0) {
            const R4<T> q = ld4(reinterpret_cast<R4<T> *>(b.aux0) + gi);
            speed_out = q.x;
            dhead_out = q.y;
        }
        if (dyn != CAV_DYN_EXTERNAL) {
            if (dyn == CAV_DYN_UNICYCLE_MAX_TURN_RATE) {  // UnicycleDynamicsMaxTurnRate.py:30-32
                // `action` is a float32 array (env.py:305-307) and dt a Python float, so under NumPy >= 2 (NEP 50) the
                // turning rate, its clip and `turning_rate*dt` are float32 operations; the sum with the heading is float64
                float tr = (float)a1 / (float)dt;
                tr = tr < -3.0f ? -3.0f : (tr > 3.0f ? 3.0f : tr);
                a1 = (T)(tr * (float)dt);
            }
            const T heading = ht.x;
            const T sel = wrap<T>(a1 + heading);  // UnicycleDynamics.py:28
            const T c = M<T>::cos_(sel), s = M<T>::sin_(sel);
            const T dx = a0 * c * dt, dy = a0 * s * dt;  // :30-31
            pv.x += dx;
            pv.y += dy;
            pv.z = a0 * c;  // :34-35
            pv.w = a0 * s;
            speed_out = a0;
            dhead_out = wrap<T>(sel - heading);  // :37-38
            ht.x = sel;
            if (dyn == CAV_DYN_UNICYCLE) {  // turning_dir, :42-47
                T td = ht.w;
                if (M<T>::abs_(td) < (T)1e-5) {
                    td = (T)0.11 * sign<T>(sel);
                } else if (td * sel < (T)0) {
                    T v = -td + sel;
                    v = v < pi<T>() ? v : pi<T>();
                    v = v > -pi<T>() ? v : -pi<T>();
                    td = v;
                } else {
                    T m = M<T>::abs_(td) - (T)0.1;
                    m = m > (T)0 ? m : (T)0.0;
                    td = sign<T>(td) * m;
                }
                ht.w = td;
            }
        }
        R4<T> ego;
        T vex, vey;
        update_ego<T>(pv, gl, ht.x, ego, vex, vey, b.aux0 != nullptr);  // agent.py:227
        if (b.aux0) st4(reinterpret_cast<R4<T> *>(b.aux0) + gi, R4<T>{speed_out, dhead_out, vex, vey});
        const int sn = b.step_num[gi];
        if (p.store_history && b.history && sn < p.history_len) {  // agent.py:257-260,275-289
            T *row = reinterpret_cast<T *>(b.history) + ((size_t)gi * p.history_len + sn) * CAV_HISTORY_ROW;
            const T v[CAV_HISTORY_ROW] = {ht.z, pv.x, pv.y, gl.x, gl.y, gl.z, gl.w, pv.z, pv.w,
                                          (dyn != CAV_DYN_EXTERNAL) ? a0 : speed_out, ht.x, ht.z, ego.x, ego.y};
            store_history_row(row, v);
        }
        // _check_if_at_goal, agent.py:150-153
        const T ex = pv.x - gl.x, ey = pv.y - gl.y;
        if (ex * ex + ey * ey <= p.near_goal_thr * p.near_goal_thr) f |= CAV_FLAG_AT_GOAL;
        else f &= ~CAV_FLAG_AT_GOAL;
        if (b.past_vel) {
            R4<T> *pvel = reinterpret_cast<R4<T> *>(b.past_vel) + gi;
            R4<T> q = ld4(pvel);
            q.z = q.x;
            q.w = q.y;
            q.x = pv.z;
            q.y = pv.w;
            st4(pvel, q);
        }
        ht.y -= dt;  // agent.py:235-239
        ht.z += dt;
        b.step_num[gi] = sn + 1;
        if (ht.y <= (T)0) f |= CAV_FLAG_RAN_OUT_OF_TIME;
        st4(g_pv + gi, pv);
        st4(g_ht + gi, ht);
        st4(g_ego + gi, ego);
        b.flags[gi] = f;
    }
}

}  // namespace cav
