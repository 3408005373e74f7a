/*
 * cav_oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Included twice by cav_oracle.c with
 * REAL = double (the reference's NumPy float64 arithmetic) and REAL = float (the same algorithm restated in
 * float32, the checker for the CAV_F32 layout).  See cav_oracle.c for the pinning statement.
 */

typedef struct {
    REAL x, y, z, w;
} SFX(r4);

#define R_PI ((REAL)3.14159265358979323846)
#define R_TWO_PI ((REAL)(2.0 * 3.14159265358979323846))

/* util.py:141-146 */
static inline REAL SFX(wrap)(REAL a) {
    while (a >= R_PI) a -= R_TWO_PI;
    while (a < -R_PI) a += R_TWO_PI;
    return a;
}
/* np.sign */
static inline REAL SFX(sign)(REAL a) { return (a > 0) ? (REAL)1 : ((a < 0) ? (REAL)-1 : (REAL)0); }
/* Python/NumPy float % (npy_divmod): result takes the sign of the divisor */
static inline REAL SFX(pymod)(REAL a, REAL b) {
    REAL m = R_FMOD(a, b);
    if (m != 0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = (REAL)copysign(0.0, (double)b);
    }
    return m;
}

/* Dynamics.update_ego_frame, dynamics/Dynamics.py:24-41 + Agent.get_ref, agent.py:329-349.
 * Returns vel_ego through vex/vey. */
static inline void SFX(update_ego)(SFX(r4) pv, SFX(r4) goal, REAL heading, SFX(r4) * ego, REAL *vex, REAL *vey) {
    const REAL gdx = goal.x - pv.x, gdy = goal.y - pv.y;
    const REAL dist = R_SQRT(gdx * gdx + gdy * gdy);
    REAL prx, pry;
    if (dist > (REAL)1e-8) {
        prx = gdx / dist;
        pry = gdy / dist;
    } else {
        prx = gdx;
        pry = gdy;
    }
    const REAL ang = R_ATAN2(pry, prx);
    const REAL he = SFX(wrap)(heading - ang);
    const REAL cur_speed = R_SQRT(pv.z * pv.z + pv.w * pv.w);
    *vex = cur_speed * R_COS(he);
    *vey = cur_speed * R_SIN(he);
    ego->x = dist;
    ego->y = he;
    ego->z = prx;
    ego->w = pry;
}

static inline int SFX(n_active)(const cav_config_t *cfg, const cav_buffers_t *b, int e) {
    return b->num_active ? b->num_active[e] : cfg->num_agents;
}

/* ---- action selection: collision_avoidance_env.py:304-323 ------------------------------------------------- */
static void SFX(policy_env)(const cav_config_t *cfg, const cav_buffers_t *b, const float *ext, int e) {
    const int N = cfg->num_agents, n = SFX(n_active)(cfg, b, e);
    SFX(r4) *pv = (SFX(r4) *)b->pv + (size_t)e * N;
    SFX(r4) *goal = (SFX(r4) *)b->goal + (size_t)e * N;
    SFX(r4) *ht = (SFX(r4) *)b->ht + (size_t)e * N;
    SFX(r4) *ego = (SFX(r4) *)b->ego + (size_t)e * N;
    uint32_t *flags = b->flags + (size_t)e * N;
    const uint32_t *kinds = b->kinds + (size_t)e * N;
    float *act = b->actions + (size_t)e * N * 2;

    /* All policies see the pre-move state of all agents.  StaticPolicy mutates its own goal while the agent loop
     * is running (StaticPolicy.py:21), which later agents in the same loop can see through RVOPolicy's goal
     * copy; that only feeds OTHER agents' pref velocities, which do not influence the ego solve, so the order
     * is immaterial. */
    float fpos[2 * CAVO_MAX_AGENTS], fvel[2 * CAVO_MAX_AGENTS], frad[CAVO_MAX_AGENTS];
    int have_rvo = 0;
    for (int i = 0; i < n; ++i)
        if (!(flags[i] & CAV_FLAG_DONE) && CAV_KIND_POLICY(kinds[i]) == CAV_POLICY_RVO) have_rvo = 1;
    if (have_rvo) {
        /* RVOPolicy.py:57-74: every agent (done ones too) is pushed as C floats (rvo2.pyx:212-228) */
        for (int a = 0; a < n; ++a) {
            fpos[2 * a] = (float)pv[a].x;
            fpos[2 * a + 1] = (float)pv[a].y;
            fvel[2 * a] = (float)pv[a].z;
            fvel[2 * a + 1] = (float)pv[a].w;
            frad[a] = (float)((REAL)(1 + 5e-2) * goal[a].z); /* (1+5e-2)*radius, RVOPolicy.py:71 */
        }
    }
    /* doStep -> kdTree_->buildAgentTree() (R/RVOSimulator.cpp:163-168): every live RVO agent's simulator holds the
     * same permutation (see the kd-tree notes in cav_oracle.c), so it is advanced once per environment step. */
    kd_tree_t kd;
    const kd_tree_t *tree = NULL;
    uint8_t local_perm[CAVO_MAX_AGENTS];
    if (have_rvo && n > CAVO_KD_MAX_LEAF) {
        uint8_t *perm = local_perm;
        if (b->kd_order) {
            perm = b->kd_order + (size_t)e * N;
        } else {
            for (int a = 0; a < n; ++a) local_perm[a] = (uint8_t)a;
        }
        cavo_kd_build(&kd, fpos, perm, n);
        tree = &kd;
#ifdef CAVO_TEST_NO_KD /* sensitivity experiments only: the pre-kd-tree restatement (index order) */
        tree = NULL;
#endif
    }
    for (int i = 0; i < n; ++i) {
        if (flags[i] & CAV_FLAG_DONE) continue; /* env.py:311-312 (act keeps the last live action) */
        const uint32_t pol = CAV_KIND_POLICY(kinds[i]);
        REAL a0 = 0, a1 = 0;
        if (pol == CAV_POLICY_EXTERNAL) { /* ExternalPolicy.py:14-16 */
            a0 = (REAL)ext[((size_t)e * N + i) * 2];
            a1 = (REAL)ext[((size_t)e * N + i) * 2 + 1];
        } else if (pol == CAV_POLICY_LEARNING) { /* LearningPolicy.py:30-32 */
            const REAL x0 = (REAL)ext[((size_t)e * N + i) * 2], x1 = (REAL)ext[((size_t)e * N + i) * 2 + 1];
            a1 = (REAL)cfg->max_heading_change * ((REAL)2. * x1 - (REAL)1.);
            a0 = goal[i].w * x0;
        } else if (pol == CAV_POLICY_NONCOOP) { /* NonCooperativePolicy.py:21 */
            a0 = goal[i].w;
            a1 = -ego[i].y;
        } else if (pol == CAV_POLICY_STATIC) { /* StaticPolicy.py:21-22 */
            goal[i].x = pv[i].x;
            goal[i].y = pv[i].y;
            a0 = 0;
            a1 = 0;
        } else if (pol == CAV_POLICY_GA3C) { /* GA3CCADRLPolicy.py:49-84 on the agent's observation of the last step */
            if (!g_ga3c_blob) continue;
            int16_t src[CAV_GA3C_X_LEN];
            cavo_ga3c_gather(cfg, src);
            const REAL *orow = (const REAL *)b->obs + ((size_t)e * N + i) * cfg->obs_len;
            float X[CAV_GA3C_X_LEN];
            for (int k = 0; k < CAV_GA3C_X_LEN; ++k) X[k] = src[k] >= 0 ? (float)orow[src[k]] : 0.0f; /* float32 feed */
            int32_t best = 0;
            cavo_ga3c_forward(g_ga3c_blob, X, 1, NULL, &best);
            double sp, dh;
            ga3c_action(best, &sp, &dh);
            /* pref_speed * raw_action[0] lands in the float32 all_actions array (env.py:305-307); raw_action[0] is 1, 0.5 or
             * 0, so rounding pref_speed to float32 first (the observation column may already be float32) gives the same value */
            a0 = (REAL)((double)(float)orow[cfg->obs_offset[CAV_OBS_PREF_SPEED]] * sp);
            a1 = (REAL)dh;
        } else if (pol == CAV_POLICY_RVO) { /* RVOPolicy.py:50-122 */
            const REAL gx = goal[i].x - pv[i].x, gy = goal[i].y - pv[i].y;      /* :66 */
            const REAL nrm = R_SQRT(gx * gx + gy * gy);                         /* np.linalg.norm */
            const REAL sc = goal[i].w / nrm;                                    /* :67 */
            const float prefx = (float)(sc * gx), prefy = (float)(sc * gy);
            const float maxSpeed = (float)goal[i].w; /* :70 */
            float nvx, nvy;
            /* RVOPolicy.py:76-90: the ego's collaboration coefficient.  Negative = anti-collaborative: every
             * RVO_ANTI_COLLAB_T seconds of agent time (within DT of a multiple of it) the agent re-draws whether it ignores
             * the others (use_non_coop_policy, probability 1 - |coeff| -> coefficient 0: no half-planes at all) or
             * steers into its closest neighbour (the negative coefficient, R/Agent.cpp:368-415). */
            float ego_collab = (float)cfg->rvo_collab_coeff;
            if (cfg->rvo_collab_coeff < 0) {
                const REAL T_ = (REAL)cfg->rvo_anti_collab_t;
                const REAL tm = SFX(pymod)(ht[i].z, T_);
                const REAL r1 = R_RINT(tm * (REAL)1000) / (REAL)1000, r2 = R_RINT((T_ - tm) * (REAL)1000) / (REAL)1000;
                if (r1 < (REAL)cfg->dt || r2 < (REAL)cfg->dt) {
                    const int ep = b->episode ? b->episode[e] : 0;
                    const double u = cavo_philox_u01(cfg->rvo_seed, (uint32_t)e, (uint32_t)i, (uint32_t)ep,
                                                     (uint32_t)b->step_num[(size_t)e * N + i]);
                    if (u < 1.0 - fabs(cfg->rvo_collab_coeff)) flags[i] |= CAV_FLAG_RVO_NON_COOP;
                    else flags[i] &= ~(uint32_t)CAV_FLAG_RVO_NON_COOP;
                }
                if (flags[i] & CAV_FLAG_RVO_NON_COOP) ego_collab = 0.0f;
            }
            const int reps = g_faithful_dostep ? n : 1;
            for (int r = 0; r < reps; ++r) { /* reps>1 only burns the reference's redundant solves */
                const int who = g_faithful_dostep ? r : i;
                float tx, ty;
                float px_ = prefx, py_ = prefy, ms = maxSpeed;
                if (who != i) {
                    const REAL ogx = goal[who].x - pv[who].x, ogy = goal[who].y - pv[who].y;
                    const REAL on = R_SQRT(ogx * ogx + ogy * ogy);
                    px_ = (float)(goal[who].w / on * ogx);
                    py_ = (float)(goal[who].w / on * ogy);
                    ms = (float)goal[who].w;
                }
                cavo_orca_solve_kd(n, who, fpos, fvel, frad, ms, px_, py_, (float)cfg->rvo_time_horizon, (float)cfg->dt,
                                   who == i ? ego_collab : 0.5f, (float)cfg->rvo_neighbor_dist,
                                   cfg->rvo_max_neighbors, tree, &tx, &ty, NULL, NULL);
                if (who == i) {
                    nvx = tx;
                    nvy = ty;
                }
            }
            /* Agent::update, R/Agent.cpp:470-475: position_ += velocity_ * timeStep_ in float */
            const float tstep = (float)cfg->dt;
            const float npx = fpos[2 * i] + nvx * tstep, npy = fpos[2 * i + 1] + nvy * tstep;
            const REAL dpx = (REAL)npx - pv[i].x, dpy = (REAL)npy - pv[i].y; /* :96-97 */
            const REAL ang1 = R_ATAN2(dpy, dpx);
            const REAL new_heading = SFX(pymod)(ang1 - (REAL)0, R_TWO_PI);  /* :101, ang2 = arctan2(0,1) = 0 */
            REAL dh = SFX(wrap)(new_heading - ht[i].x);                     /* :102 */
            REAL spd = ((REAL)1 / (REAL)cfg->dt) * R_SQRT(dpx * dpx + dpy * dpy); /* :105 */
            if (R_FABS(dh) > (REAL)cfg->rvo_max_delta_heading) {              /* :108-111 */
                dh = SFX(sign)(dh) * (REAL)cfg->rvo_max_delta_heading;
                spd = 0;
            }
            a0 = spd;
            a1 = dh;
        }
        act[2 * i] = (float)a0; /* env.py:305-307: stored into a float32 array */
        act[2 * i + 1] = (float)a1;
    }
}

/* ---- move: collision_avoidance_env.py:326-327 -> Agent.take_action, agent.py:192-241 ----------------------- */
static void SFX(dynamics_env)(const cav_config_t *cfg, const cav_buffers_t *b, REAL dt, int e) {
    const int N = cfg->num_agents, n = SFX(n_active)(cfg, b, e);
    SFX(r4) *pv = (SFX(r4) *)b->pv + (size_t)e * N;
    SFX(r4) *goal = (SFX(r4) *)b->goal + (size_t)e * N;
    SFX(r4) *ht = (SFX(r4) *)b->ht + (size_t)e * N;
    SFX(r4) *ego = (SFX(r4) *)b->ego + (size_t)e * N;
    uint32_t *flags = b->flags + (size_t)e * N;
    const uint32_t *kinds = b->kinds + (size_t)e * N;
    int32_t *step_num = b->step_num + (size_t)e * N;
    const float *act = b->actions + (size_t)e * N * 2;
    SFX(r4) *aux0 = b->aux0 ? (SFX(r4) *)b->aux0 + (size_t)e * N : NULL;
    SFX(r4) *pact = b->past_actions ? (SFX(r4) *)b->past_actions + (size_t)e * N : NULL;
    SFX(r4) *pvel = b->past_vel ? (SFX(r4) *)b->past_vel + (size_t)e * N : NULL;
    const REAL thr = (REAL)cfg->near_goal_threshold;

    for (int i = 0; i < n; ++i) {
        uint32_t f = flags[i];
        if (f & (CAV_FLAG_AT_GOAL | CAV_FLAG_RAN_OUT_OF_TIME | CAV_FLAG_IN_COLLISION)) { /* agent.py:202-209 */
            if (f & CAV_FLAG_AT_GOAL) f |= CAV_FLAG_WAS_AT_GOAL_ALREADY;
            if (f & CAV_FLAG_IN_COLLISION) f |= CAV_FLAG_WAS_IN_COLLISION_ALREADY;
            pv[i].z = 0;
            pv[i].w = 0;
            flags[i] = f;
            if (pvel) { /* _store_past_velocities, agent.py:351-353 */
                pvel[i].z = pvel[i].x;
                pvel[i].w = pvel[i].y;
                pvel[i].x = 0;
                pvel[i].y = 0;
            }
            continue;
        }
        const REAL a0 = (REAL)act[2 * i];
        REAL a1 = (REAL)act[2 * i + 1];
        if (pact) { /* agent.py:212-213 */
            pact[i].z = pact[i].x;
            pact[i].w = pact[i].y;
            pact[i].x = a0;
            pact[i].y = a1;
        }
        const uint32_t dyn = CAV_KIND_DYNAMICS(kinds[i]);
        REAL speed_out = aux0 ? aux0[i].x : 0, dhead_out = aux0 ? aux0[i].y : 0;
        if (dyn != CAV_DYN_EXTERNAL) {
            if (dyn == CAV_DYN_UNICYCLE_MAX_TURN_RATE) { /* UnicycleDynamicsMaxTurnRate.py:30-32 */
                /* `action` is a float32 array (env.py:305-307) and dt a Python float: under NumPy >= 2 (NEP 50)
                 * action[1]/dt, np.clip(.., -3, 3) and turning_rate*dt are float32 operations, the sum with the
                 * float64 heading is float64 (pinned by golden maxturn6_ext, minted with NumPy 2.3) */
                float tr = (float)a1 / (float)dt;
                tr = tr < -3.0f ? -3.0f : (tr > 3.0f ? 3.0f : tr);
                a1 = (REAL)(tr * (float)dt);
            }
            const REAL heading = ht[i].x;
            const REAL sel = SFX(wrap)(a1 + heading); /* UnicycleDynamics.py:28 */
            const REAL c = R_COS(sel), s = R_SIN(sel);
            const REAL dx = a0 * c * dt, dy = a0 * s * dt; /* :30-31 */
            pv[i].x += dx;
            pv[i].y += dy;
            pv[i].z = a0 * c; /* :34-35 */
            pv[i].w = a0 * s;
            speed_out = a0;
            dhead_out = SFX(wrap)(sel - heading); /* :37-38 */
            ht[i].x = sel;
            if (dyn == CAV_DYN_UNICYCLE) { /* turning_dir, :42-47 */
                REAL td = ht[i].w;
                if (R_FABS(td) < (REAL)1e-5) {
                    td = (REAL)0.11 * SFX(sign)(sel);
                } else if (td * sel < 0) {
                    REAL v = -td + sel;
                    v = v < R_PI ? v : R_PI;   /* min(np.pi, .) */
                    v = v > -R_PI ? v : -R_PI; /* max(-np.pi, .) */
                    td = v;
                } else {
                    REAL m = R_FABS(td) - (REAL)0.1;
                    m = m > 0 ? m : (REAL)0.0;
                    td = SFX(sign)(td) * m;
                }
                ht[i].w = td;
            }
        }
        REAL vex, vey;
        SFX(update_ego)(pv[i], goal[i], ht[i].x, &ego[i], &vex, &vey); /* agent.py:227 */
        if (aux0) {
            aux0[i].x = speed_out;
            aux0[i].y = dhead_out;
            aux0[i].z = vex;
            aux0[i].w = vey;
        }
        if (cfg->store_history && b->history && step_num[i] < cfg->history_len) { /* agent.py:257-260,275-289 */
            REAL *row = (REAL *)b->history + (((size_t)e * N + i) * cfg->history_len + step_num[i]) * CAV_HISTORY_ROW;
            row[0] = ht[i].z;
            row[1] = pv[i].x;
            row[2] = pv[i].y;
            row[3] = goal[i].x;
            row[4] = goal[i].y;
            row[5] = goal[i].z;
            row[6] = goal[i].w;
            row[7] = pv[i].z;
            row[8] = pv[i].w;
            row[9] = (dyn != CAV_DYN_EXTERNAL) ? a0 : speed_out;
            row[10] = ht[i].x;
            row[11] = ht[i].z;
            row[12] = ego[i].x;
            row[13] = ego[i].y;
        }
        /* _check_if_at_goal, agent.py:150-153 */
        const REAL ex = pv[i].x - goal[i].x, ey = pv[i].y - goal[i].y;
        if (ex * ex + ey * ey <= thr * thr) f |= CAV_FLAG_AT_GOAL;
        else f &= ~CAV_FLAG_AT_GOAL;
        if (pvel) {
            pvel[i].z = pvel[i].x;
            pvel[i].w = pvel[i].y;
            pvel[i].x = pv[i].z;
            pvel[i].y = pv[i].w;
        }
        ht[i].y -= dt; /* agent.py:235-239 */
        ht[i].z += dt;
        step_num[i] += 1;
        if (ht[i].y <= 0) f |= CAV_FLAG_RAN_OUT_OF_TIME;
        flags[i] = f;
    }
}

/* ---- collisions + rewards + observation + done:
 *      collision_avoidance_env.py:394-512 (rewards/collisions), :555-575 + OtherAgentsStatesSensor.py:58-144 +
 *      agent.py:323-327 (observation), :514-553 (done) -------------------------------------------------------- */
typedef struct {
    int idx;
    REAL rdist;  /* round(dist_2_other, 2) */
    REAL porth;
    REAL tti;    /* time_to_impact (only for AGENT_SORTING_METHOD == 'time_to_impact'), else 0 */
} SFX(crit_t);

/* mode 0: (rdist, porth); 1: (-rdist, porth); 2: (-tti, -rdist, porth)  (OtherAgentsStatesSensor.py:33-50) */
static int SFX(crit_less)(const SFX(crit_t) * a, const SFX(crit_t) * b, int mode) {
    if (mode == 2) {
        const REAL ta = -a->tti, tb = -b->tti;
        if (ta < tb) return 1;
        if (ta > tb) return 0;
    }
    const REAL da = mode ? -a->rdist : a->rdist, db = mode ? -b->rdist : b->rdist;
    if (da < db) return 1;
    if (da > db) return 0;
    return a->porth < b->porth;
}

/* util.compute_time_to_impact (util.py:23-79) with tangent_vecs_from_external_pt (:81-112), operation for operation:
 * host at (px,py) with velocity hv, other at (a,b) with velocity ov, r = combined radius. */
static REAL SFX(time_to_impact)(REAL px, REAL py, REAL a, REAL b, REAL hvx, REAL hvy, REAL ovx, REAL ovy, REAL r) {
    const REAL v0 = hvx - ovx, v1 = hvy - ovy; /* v_rel */
    const REAL sq = (px - a) * (px - a) + (py - b) * (py - b) - r * r;
    if (sq < 0) return (REAL)0.0; /* already inside the collision zone */
    const REAL sqrt_term = R_SQRT((px - a) * (px - a) + (py - b) * (py - b) - r * r);
    const REAL xnum1 = (r * r) * (px - a), xnum2 = r * (py - b) * sqrt_term;
    const REAL ynum1 = (r * r) * (py - b), ynum2 = r * (px - a) * sqrt_term;
    const REAL den = (px - a) * (px - a) + (py - b) * (py - b);
    const REAL c1x = ((xnum1 + xnum2) / den + a) - px, c1y = ((ynum1 - ynum2) / den + b) - py; /* vec1 */
    const REAL c2x = ((xnum1 - xnum2) / den + a) - px, c2y = ((ynum1 + ynum2) / den + b) - py; /* vec2 */
    const REAL k1v = c1x * v1 - c1y * v0, k12 = c1x * c2y - c1y * c2x; /* np.cross of 2-vectors */
    const REAL k2v = c2x * v1 - c2y * v0, k21 = c2x * c1y - c2y * c1x;
    if (!(k1v * k12 >= 0 && k2v * k21 >= 0)) return (REAL)INFINITY; /* v_rel outside the collision cone */
    if (R_FABS(v0) < (REAL)1e-5 && R_FABS(v1) < (REAL)1e-5) return (REAL)INFINITY;
    REAL x1, x2, y1, y2;
    if (R_FABS(v0) < (REAL)1e-5) { /* vertical v_rel */
        const REAL A = 1, B = -2 * b, C = b * b + (px - a) * (px - a) - r * r;
        x1 = x2 = px;
        y1 = (-B + R_SQRT(B * B - 4 * A * C)) / (2 * A);
        y2 = (-B - R_SQRT(B * B - 4 * A * C)) / (2 * A);
    } else {
        const REAL m = v1 / v0;
        const REAL A = 1 + m * m;
        const REAL B = -2 * a + 2 * m * (py - b - m * px);
        const REAL C = a * a - r * r + (m * px - (py - b)) * (m * px - (py - b));
        x1 = (-B + R_SQRT(B * B - 4 * A * C)) / (2 * A);
        x2 = (-B - R_SQRT(B * B - 4 * A * C)) / (2 * A);
        y1 = m * (x1 - px) + py;
        y2 = m * (x2 - px) + py;
    }
    const REAL d1 = R_SQRT((x1 - px) * (x1 - px) + (y1 - py) * (y1 - py));
    const REAL d2 = R_SQRT((x2 - px) * (x2 - px) + (y2 - py) * (y2 - py));
    const REAL d = d2 < d1 ? d2 : d1; /* Python min(d1, d2) */
    return d / R_SQRT(v0 * v0 + v1 * v1);
}
/* Python's sorted() is stable: insertion sort keeps equal keys in input order */
static void SFX(stable_sort)(SFX(crit_t) * v, int n, int negate_dist) {
    for (int i = 1; i < n; ++i) {
        SFX(crit_t) x = v[i];
        int j = i - 1;
        while (j >= 0 && SFX(crit_less)(&x, &v[j], negate_dist)) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = x;
    }
}

static void SFX(pairwise_obs_env)(const cav_config_t *cfg, const cav_buffers_t *b, int e, int obs_only) {
    const int N = cfg->num_agents, n = SFX(n_active)(cfg, b, e);
    const SFX(r4) *pv = (const SFX(r4) *)b->pv + (size_t)e * N;
    const SFX(r4) *goal = (const SFX(r4) *)b->goal + (size_t)e * N;
    const SFX(r4) *ego = (const SFX(r4) *)b->ego + (size_t)e * N;
    uint32_t *flags = b->flags + (size_t)e * N;
    const uint32_t *kinds = b->kinds + (size_t)e * N;
    const float *act = b->actions + (size_t)e * N * 2;
    REAL *reward = (REAL *)b->reward + (size_t)e * N;
    REAL *mind = b->min_dist ? (REAL *)b->min_dist + (size_t)e * N : NULL;
    uint8_t *done = b->done + (size_t)e * N;
    const int L = cfg->obs_len;
    REAL *obs = (REAL *)b->obs + (size_t)e * N * L;

    if (!obs_only) {
        /* _check_for_collisions, env.py:458-512 */
        int coll[CAVO_MAX_AGENTS];
        REAL md[CAVO_MAX_AGENTS];
        for (int i = 0; i < n; ++i) {
            coll[i] = 0;
            md[i] = (REAL)INFINITY;
        }
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) {
                const REAL dx = pv[i].x - pv[j].x, dy = pv[i].y - pv[j].y;
                const REAL d = R_SQRT(dx * dx + dy * dy); /* util.l2norm, util.py:17-21 */
                const REAL cr = goal[i].z + goal[j].z;
                const REAL gap = d - cr;
                md[i] = md[i] < gap ? md[i] : gap; /* min(a, b): b if b < a else a */
                md[j] = md[j] < gap ? md[j] : gap;
                if (d <= cr) {
                    coll[i] = 1;
                    coll[j] = 1;
                }
            }
        /* wall collisions, env.py:494-506: any static-map cell under the agent's disc mask (Map.py:52-56), only when
         * the agent's centre cell is inside the map */
        int wall[CAVO_MAX_AGENTS];
        for (int i = 0; i < n; ++i) wall[i] = 0;
        if (cfg->use_static_map && b->static_map_bits) {
            const cavo_map_t m = cavo_map_dims(cfg);
            const int wpr = (m.cols + 31) / 32;
            for (int i = 0; i < n; ++i) {
                int gx, gy;
                if (!cavo_world_to_cell(&m, (double)pv[i].x, (double)pv[i].y, &gx, &gy)) continue;
                const double rc = (double)goal[i].z / m.cell;
                const double rr2 = rc * rc;
                for (int r = 0; r < m.rows && !wall[i]; ++r)
                    for (int c = 0; c < m.cols; ++c) {
                        const double dc = (double)(c - gy), dr = (double)(r - gx);
                        if (dc * dc + dr * dr < rr2 && ((b->static_map_bits[r * wpr + (c >> 5)] >> (c & 31)) & 1u)) {
                            wall[i] = 1;
                            break;
                        }
                    }
            }
        }
        /* rewards, env.py:394-456 */
        REAL rmin = (REAL)cfg->reward_at_goal, rmax = rmin;
        const REAL vals[5] = {(REAL)cfg->reward_at_goal, (REAL)cfg->reward_collision_with_agent,
                              (REAL)cfg->reward_time_step, (REAL)cfg->reward_collision_with_wall,
                              (REAL)cfg->reward_wiggly_behavior};
        for (int k = 0; k < 5; ++k) {
            rmin = vals[k] < rmin ? vals[k] : rmin;
            rmax = vals[k] > rmax ? vals[k] : rmax;
        }
        for (int i = 0; i < n; ++i) {
            uint32_t f = flags[i] & ~(CAV_FLAG_COLLIDED_AGENT | CAV_FLAG_COLLIDED_WALL);
            if (coll[i]) f |= CAV_FLAG_COLLIDED_AGENT;
            if (wall[i]) f |= CAV_FLAG_COLLIDED_WALL;
            REAL r = (REAL)cfg->reward_time_step;
            if (f & CAV_FLAG_AT_GOAL) {
                if (!(f & CAV_FLAG_WAS_AT_GOAL_ALREADY)) r = (REAL)cfg->reward_at_goal;
            } else if (!(f & CAV_FLAG_WAS_IN_COLLISION_ALREADY)) {
                if (coll[i]) {
                    r = (REAL)cfg->reward_collision_with_agent;
                    f |= CAV_FLAG_IN_COLLISION;
                } else if (wall[i]) {
                    r = (REAL)cfg->reward_collision_with_wall;
                    f |= CAV_FLAG_IN_COLLISION;
                } else {
                    if (md[i] <= (REAL)cfg->getting_close_range) r = (REAL)-0.1 - md[i] / (REAL)2.0;
                    if (R_FABS((REAL)act[2 * i + 1]) > (REAL)cfg->wiggly_behavior_threshold)
                        r += (REAL)cfg->reward_wiggly_behavior;
                }
            }
            r = r < rmin ? rmin : (r > rmax ? rmax : r); /* np.clip */
            reward[i] = r;
            if (mind) mind[i] = md[i];
            flags[i] = f;
        }
    }

    /* observation: OtherAgentsStatesSensor.sense + get_observation_dict */
    const int K = cfg->max_num_other_agents_observed;
    const int *off = cfg->obs_offset;
    for (int h = 0; h < n; ++h) {
        REAL *o = obs + (size_t)h * L;
        SFX(crit_t) crit[CAVO_MAX_AGENTS];
        int nc = 0;
        const REAL prx = ego[h].z, pry = ego[h].w;
        const REAL orx = -pry, ory = prx; /* ref_orth, agent.py:347 */
        for (int i = 0; i < n; ++i) {
            if (i == h) continue;
            const REAL rx = pv[i].x - pv[h].x, ry = pv[i].y - pv[h].y;
            const REAL p_orth = rx * orx + ry * ory;
            const REAL dc = R_SQRT(rx * rx + ry * ry); /* vec2_l2_norm, util.py:114-116 */
            const REAL d2o = dc - goal[h].z - goal[i].z;
            if (dc > (REAL)cfg->sensing_horizon) continue;
            crit[nc].idx = i;
            crit[nc].rdist = R_RINT(d2o * (REAL)100) / (REAL)100; /* round(np.float64, 2) */
            crit[nc].porth = p_orth;
            crit[nc].tti = (REAL)0;
            if (cfg->agent_sorting_method == CAV_SORT_TIME_TO_IMPACT) /* :92-101 */
                crit[nc].tti = SFX(time_to_impact)(pv[h].x, pv[h].y, pv[i].x, pv[i].y, pv[h].z, pv[h].w, pv[i].z,
                                                   pv[i].w, goal[h].z + goal[i].z);
            ++nc;
        }
        const int tti_mode = cfg->agent_sorting_method == CAV_SORT_TIME_TO_IMPACT;
        SFX(stable_sort)(crit, nc, tti_mode ? 2 : 0); /* OtherAgentsStatesSensor.py:33-38 */
        if (nc > K) nc = K;                            /* :39 */
        if (cfg->agent_sorting_method == CAV_SORT_CLOSEST_LAST) SFX(stable_sort)(crit, nc, 1); /* :42-44 */
        /* time_to_impact: the second sort (:48-50) uses the same key as the first, a no-op on the clipped list */
        REAL first_row[7] = {0, 0, 0, 0, 0, 0, 0};
        int have_first = 0;
        if (off[CAV_OBS_OTHER_AGENTS_STATES] >= 0)
            for (int k = 0; k < K * 7; ++k) o[off[CAV_OBS_OTHER_AGENTS_STATES] + k] = 0; /* :108 */
        for (int k = 0; k < nc; ++k) {
            const int i = crit[k].idx;
            const REAL rx = pv[i].x - pv[h].x, ry = pv[i].y - pv[h].y;
            REAL row[7];
            row[0] = rx * prx + ry * pry;
            row[1] = rx * orx + ry * ory;
            row[2] = pv[i].z * prx + pv[i].w * pry;
            row[3] = pv[i].z * orx + pv[i].w * ory;
            row[4] = goal[i].z;
            row[5] = goal[h].z + goal[i].z;
            row[6] = R_SQRT(rx * rx + ry * ry) - goal[h].z - goal[i].z;
            if (k == 0) {
                memcpy(first_row, row, sizeof(row));
                have_first = 1;
            }
            if (off[CAV_OBS_OTHER_AGENTS_STATES] >= 0)
                memcpy(o + off[CAV_OBS_OTHER_AGENTS_STATES] + 7 * k, row, sizeof(row));
        }
        if (off[CAV_OBS_OTHER_AGENT_STATES] >= 0 && have_first)
            memcpy(o + off[CAV_OBS_OTHER_AGENT_STATES], first_row, sizeof(first_row)); /* :136-137 (persists) */
        if (off[CAV_OBS_IS_LEARNING] >= 0)
            o[off[CAV_OBS_IS_LEARNING]] = (CAV_KIND_POLICY(kinds[h]) == CAV_POLICY_LEARNING) ? (REAL)1 : (REAL)0;
        if (off[CAV_OBS_NUM_OTHER_AGENTS] >= 0) o[off[CAV_OBS_NUM_OTHER_AGENTS]] = (REAL)nc;
        if (off[CAV_OBS_DIST_TO_GOAL] >= 0) o[off[CAV_OBS_DIST_TO_GOAL]] = ego[h].x;
        if (off[CAV_OBS_HEADING_EGO_FRAME] >= 0) o[off[CAV_OBS_HEADING_EGO_FRAME]] = ego[h].y;
        if (off[CAV_OBS_PREF_SPEED] >= 0) o[off[CAV_OBS_PREF_SPEED]] = goal[h].w;
        if (off[CAV_OBS_RADIUS] >= 0) o[off[CAV_OBS_RADIUS]] = goal[h].z;
    }

    if (!obs_only) {
        /* _check_which_agents_done, env.py:514-553 */
        int all_done = 1, all_learning_done = 1;
        for (int i = 0; i < n; ++i) {
            uint32_t f = flags[i];
            const int d = (f & (CAV_FLAG_AT_GOAL | CAV_FLAG_RAN_OUT_OF_TIME | CAV_FLAG_IN_COLLISION)) != 0;
            f = d ? (f | CAV_FLAG_DONE) : (f & ~CAV_FLAG_DONE);
            flags[i] = f;
            done[i] = (uint8_t)d;
            if (!d) {
                all_done = 0;
                if (CAV_KIND_POLICY(kinds[i]) == CAV_POLICY_LEARNING) all_learning_done = 0;
            }
        }
        int go;
        if (cfg->game_over_mode == CAV_GAME_OVER_ALL_DONE) go = all_done;
        else if (cfg->game_over_mode == CAV_GAME_OVER_AGENT0) go = done[0];
        else go = all_learning_done;
        b->game_over[e] = (uint8_t)go;
    }
}

/* ---- laserscan: Map.add_agents_to_map (Map.py:46-64) + LaserScanSensor.sense (LaserScanSensor.py:49-101) ---- */
static void SFX(laserscan_env)(const cav_config_t *cfg, const cav_buffers_t *b, int e, uint8_t *grid) {
    const int N = cfg->num_agents, n = SFX(n_active)(cfg, b, e);
    const SFX(r4) *pv = (const SFX(r4) *)b->pv + (size_t)e * N;
    const SFX(r4) *goal = (const SFX(r4) *)b->goal + (size_t)e * N;
    const SFX(r4) *ht = (const SFX(r4) *)b->ht + (size_t)e * N;
    const cavo_map_t m = cavo_map_dims(cfg);
    const int wpr = (m.cols + 31) / 32;
    const int B = cfg->laserscan_length, P = cfg->laserscan_num_past;
    const int L = cfg->obs_len, off = cfg->obs_offset[CAV_OBS_LASERSCAN];
    const int nr = (int)ceil((cfg->laser_max_range - cfg->laser_min_range) / cfg->laser_range_resolution); /* np.arange len */

    /* Map.add_agents_to_map */
    for (int r = 0; r < m.rows; ++r)
        for (int c = 0; c < m.cols; ++c)
            grid[r * m.cols + c] =
                b->static_map_bits ? (uint8_t)((b->static_map_bits[r * wpr + (c >> 5)] >> (c & 31)) & 1u) : 0;
    int cgx[CAVO_MAX_AGENTS], cgy[CAVO_MAX_AGENTS], cin[CAVO_MAX_AGENTS];
    double rr2[CAVO_MAX_AGENTS];
    for (int a = 0; a < n; ++a) {
        cin[a] = cavo_world_to_cell(&m, (double)pv[a].x, (double)pv[a].y, &cgx[a], &cgy[a]);
        const double rc = (double)goal[a].z / m.cell;
        rr2[a] = rc * rc;
        if (!cin[a]) continue; /* Map.py:58-64: off-map agents paint nothing */
        for (int r = 0; r < m.rows; ++r)
            for (int c = 0; c < m.cols; ++c) {
                const double dc = (double)(c - cgy[a]), dr = (double)(r - cgx[a]);
                if (dc * dc + dr * dr < rr2[a]) grid[r * m.cols + c] = 1; /* Map.py:55 */
            }
    }
    const double step = (cfg->laser_max_angle - cfg->laser_min_angle) / (double)(B - 1); /* np.linspace */
    for (int h = 0; h < n; ++h) {
        REAL *o = (REAL *)b->obs + ((size_t)e * N + h) * L + off;
        uint32_t *cnt = b->laser_count + (size_t)e * N + h;
        /* history roll first (LaserScanSensor.py:84-88), newest row 0 */
        if (*cnt != 0)
            for (int p = P - 1; p >= 1; --p) memcpy(o + (size_t)p * B, o + (size_t)(p - 1) * B, sizeof(REAL) * (size_t)B);
        for (int bi = 0; bi < B; ++bi) {
            double ang0 = (double)bi * step + cfg->laser_min_angle;
            if (bi == B - 1) ang0 = cfg->laser_max_angle; /* linspace endpoint */
            const double ang = ang0 + (double)ht[h].x; /* :66 */
            const double ca = cos(ang), sa = sin(ang);
            int cum = 0, kstar = -1;
            for (int k = 0; k < nr; ++k) {
                const double rk = cfg->laser_min_range + (double)k * cfg->laser_range_resolution; /* np.arange */
                const double x = (double)pv[h].x + rk * ca, y = (double)pv[h].y + rk * sa;       /* :70-72 */
                int gx, gy;
                int hit = 0;
                if (cavo_world_to_cell(&m, x, y, &gx, &gy)) {
                    int ego_mask = 0;
                    if (cin[h]) {
                        const double dc = (double)(gy - cgy[h]), dr = (double)(gx - cgx[h]);
                        ego_mask = dc * dc + dr * dr < rr2[h];
                    }
                    hit = grid[gx * m.cols + gy] && !ego_mask; /* :76-77 */
                }
                cum += hit;
                if (cum == 1) kstar = k; /* :78-82: last index whose cumulative hit count is 1 */
            }
            const double rng = kstar >= 0 ? cfg->laser_min_range + (double)kstar * cfg->laser_range_resolution
                                          : cfg->laser_max_range;
            if (*cnt == 0)
                for (int p = 0; p < P; ++p) o[(size_t)p * B + bi] = (REAL)rng;
            else
                o[bi] = (REAL)rng;
        }
        *cnt += 1;
    }
}

/* ---- reset: collision_avoidance_env.py:236-282, agent.py:59-138, test_cases.py:495-590 -------------------- */
static void SFX(reset_env)(const cav_config_t *cfg, const cav_buffers_t *b, const cav_scenario_bank_t *bank, int sidx,
                           int e) {
    const int N = cfg->num_agents;
    SFX(r4) *pv = (SFX(r4) *)b->pv + (size_t)e * N;
    SFX(r4) *goal = (SFX(r4) *)b->goal + (size_t)e * N;
    SFX(r4) *ht = (SFX(r4) *)b->ht + (size_t)e * N;
    SFX(r4) *ego = (SFX(r4) *)b->ego + (size_t)e * N;
    const int n = bank->num_active ? bank->num_active[sidx] : N;
    if (b->num_active) b->num_active[e] = n;
    for (int i = 0; i < N; ++i) {
        const size_t g = (size_t)e * N + i;
        const double *row = bank->rows + ((size_t)sidx * N + i) * CAV_BANK_ROW;
        if (bank->kinds) b->kinds[g] = bank->kinds[(size_t)sidx * N + i];
        if (i >= n) {
            memset(&pv[i], 0, sizeof(pv[i]));
            memset(&goal[i], 0, sizeof(goal[i]));
            memset(&ht[i], 0, sizeof(ht[i]));
            memset(&ego[i], 0, sizeof(ego[i]));
            b->flags[g] = 0;
            b->step_num[g] = 0;
            b->done[g] = 0;
            b->actions[2 * g] = b->actions[2 * g + 1] = 0;
            memset((REAL *)b->obs + g * cfg->obs_len, 0, sizeof(REAL) * (size_t)cfg->obs_len);
            continue;
        }
        pv[i].x = (REAL)row[0];
        pv[i].y = (REAL)row[1];
        pv[i].z = 0;
        pv[i].w = 0;
        goal[i].x = (REAL)row[2];
        goal[i].y = (REAL)row[3];
        goal[i].w = (REAL)row[4];
        goal[i].z = (REAL)row[5];
        ht[i].x = (REAL)row[6];
        const REAL dx = pv[i].x - goal[i].x, dy = pv[i].y - goal[i].y;
        const REAL slt = (R_SQRT(dx * dx + dy * dy) - (REAL)cfg->near_goal_threshold) / goal[i].w; /* agent.py:98 */
        REAL tr = (REAL)cfg->max_time_ratio * slt;
        tr = tr > (REAL)cfg->dt ? tr : (REAL)cfg->dt; /* max(tr, dt_nominal), agent.py:103 */
        ht[i].y = tr;
        ht[i].z = 0;
        ht[i].w = 0; /* turning_dir */
        REAL vex, vey;
        SFX(update_ego)(pv[i], goal[i], ht[i].x, &ego[i], &vex, &vey);
        b->flags[g] = CAV_FLAG_RVO_NON_COOP; /* RVOPolicy.__init__: use_non_coop_policy = True (RVOPolicy.py:32) */
        b->step_num[g] = 0;
        b->done[g] = 0;
        b->actions[2 * g] = b->actions[2 * g + 1] = 0;
        if (b->aux0) memset((SFX(r4) *)b->aux0 + g, 0, sizeof(SFX(r4)));
        if (b->past_actions) memset((SFX(r4) *)b->past_actions + g, 0, sizeof(SFX(r4)));
        if (b->past_vel) memset((SFX(r4) *)b->past_vel + g, 0, sizeof(SFX(r4)));
        if (b->laser_count) b->laser_count[g] = 0;
        if (b->kd_order) b->kd_order[g] = (uint8_t)i; /* new Agent -> new RVOPolicy -> new simulator: KdTree::agents_ = identity */
        if (b->min_dist) ((REAL *)b->min_dist)[g] = (REAL)INFINITY;
        /* env.reset zeroes the observation dict before sensing (env.py:276-281) */
        memset((REAL *)b->obs + g * cfg->obs_len, 0, sizeof(REAL) * (size_t)cfg->obs_len);
    }
    /* game_over[e] is left untouched: the auto-reset caller still needs the finished step's flag
     * (DummyVecEnv.step_wait returns done=True together with the reset observation, vec_env.py:120-128). */
}

static void SFX(accumulate_stats)(const cav_config_t *cfg, const cav_buffers_t *b, int e) {
    const int N = cfg->num_agents, n = SFX(n_active)(cfg, b, e);
    const SFX(r4) *pv = (const SFX(r4) *)b->pv + (size_t)e * N;
    const SFX(r4) *goal = (const SFX(r4) *)b->goal + (size_t)e * N;
    const SFX(r4) *ht = (const SFX(r4) *)b->ht + (size_t)e * N;
    (void)pv;
    uint64_t *st = b->stats;
    double *std_ = (double *)b->stats;
    int any_coll = 0, all_goal = 1, any_stuck = 0;
    for (int i = 0; i < n; ++i) {
        const uint32_t f = b->flags[(size_t)e * N + i];
        st[CAV_STAT_AGENTS] += 1;
        st[CAV_STAT_AGENT_STEPS] += (uint64_t)b->step_num[(size_t)e * N + i];
        if (f & CAV_FLAG_IN_COLLISION) {
            any_coll = 1;
            st[CAV_STAT_AGENTS_COLLISION] += 1;
        }
        if (f & CAV_FLAG_AT_GOAL) {
            st[CAV_STAT_AGENTS_AT_GOAL] += 1;
            std_[CAV_STAT_SUM_TIME_TO_GOAL] += (double)ht[i].z;
            /* straight_line_time_to_reach_goal is recomputed from the reset-time relation
             * time_remaining0 = max_time_ratio * slt (agent.py:98-103): slt = (time_remaining + t) / ratio */
            const double slt = ((double)ht[i].y + (double)ht[i].z) / cfg->max_time_ratio;
            std_[CAV_STAT_SUM_EXTRA_TIME] += (double)ht[i].z - slt;
        } else {
            all_goal = 0;
        }
        if (f & CAV_FLAG_RAN_OUT_OF_TIME) {
            any_stuck = 1;
            st[CAV_STAT_AGENTS_STUCK] += 1;
        }
        (void)goal;
    }
    st[CAV_STAT_EPISODES] += 1;
    st[CAV_STAT_EPISODES_COLLISION] += (uint64_t)any_coll;
    st[CAV_STAT_EPISODES_ALL_AT_GOAL] += (uint64_t)all_goal;
    st[CAV_STAT_EPISODES_ANY_STUCK] += (uint64_t)any_stuck;
}

/* ---- batched drivers ----------------------------------------------------------------------------------------- */
static int SFX(policy)(const cav_config_t *cfg, const cav_buffers_t *b, const float *ext) {
    const int E = cfg->num_envs;
#pragma omp parallel for schedule(dynamic, 16) num_threads(cavo_get_max_threads())
    for (int e = 0; e < E; ++e) SFX(policy_env)(cfg, b, ext, e);
    return 0;
}
static int SFX(dynamics)(const cav_config_t *cfg, const cav_buffers_t *b, double dt) {
    const int E = cfg->num_envs;
#pragma omp parallel for schedule(static) num_threads(cavo_get_max_threads())
    for (int e = 0; e < E; ++e) SFX(dynamics_env)(cfg, b, (REAL)dt, e);
    return 0;
}
static int SFX(pairwise_obs)(const cav_config_t *cfg, const cav_buffers_t *b) {
    const int E = cfg->num_envs;
#pragma omp parallel for schedule(static) num_threads(cavo_get_max_threads())
    for (int e = 0; e < E; ++e) SFX(pairwise_obs_env)(cfg, b, e, 0);
    return 0;
}
static int SFX(laserscan)(const cav_config_t *cfg, const cav_buffers_t *b) {
    const int E = cfg->num_envs;
    const cavo_map_t m = cavo_map_dims(cfg);
#pragma omp parallel num_threads(cavo_get_max_threads())
    {
        uint8_t *grid = (uint8_t *)malloc((size_t)m.rows * m.cols);
#pragma omp for schedule(dynamic, 1)
        for (int e = 0; e < E; ++e) SFX(laserscan_env)(cfg, b, e, grid);
        free(grid);
    }
    return 0;
}
static int SFX(reset)(const cav_config_t *cfg, const cav_buffers_t *b, const uint8_t *mask,
                      const cav_scenario_bank_t *bank, const int32_t *scenario_idx, int accumulate_stats) {
    const int E = cfg->num_envs;
    const cavo_map_t m = cavo_map_dims(cfg);
    uint8_t *grid = cfg->obs_offset[CAV_OBS_LASERSCAN] >= 0 ? (uint8_t *)malloc((size_t)m.rows * m.cols) : NULL;
    for (int e = 0; e < E; ++e) {
        if (mask && !mask[e]) continue;
        if (accumulate_stats && b->stats) SFX(accumulate_stats)(cfg, b, e);
        const int ep = b->episode ? b->episode[e] : 0;
        const int sidx = scenario_idx ? scenario_idx[e]
                                      : (int)(((int64_t)e + (int64_t)ep * E) % bank->num_scenarios);
        if (b->episode) b->episode[e] = ep + 1;
        SFX(reset_env)(cfg, b, bank, sidx, e);
        SFX(pairwise_obs_env)(cfg, b, e, 1); /* reset returns self._get_obs(), env.py:282 */
        if (grid) SFX(laserscan_env)(cfg, b, e, grid);
    }
    free(grid);
    return 0;
}
