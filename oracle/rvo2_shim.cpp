// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// extern "C" shim over the UNMODIFIED reference RVO2 library
// (/root/reference/gym_collision_avoidance/envs/policies/Python-RVO2/src/*.cpp,
// compiled where it lies by oracle/Makefile into oracle/_ref/librvo2_ref.so).
// It exposes the subset of RVO::RVOSimulator that the reference's Cython
// binding `rvo2.pyx:80-230` exposes and that `RVOPolicy.py:25-96` calls, so the
// reference Python package can run here without a Cython build, and so that the
// C restatement (oracle/cav_oracle.c) can be pinned bit-for-bit against the real
// solver.  No reference source is copied: only its public headers are included.
#include "RVO.h"
#include <cstddef>

using RVO::RVOSimulator;
using RVO::Vector2;

extern "C" {

void *rvo2ref_new(float timeStep, float neighborDist, size_t maxNeighbors, float timeHorizon,
                  float timeHorizonObst, float radius, float maxSpeed, float vx, float vy) {
    return new RVOSimulator(timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius,
                            maxSpeed, Vector2(vx, vy));
}
void rvo2ref_delete(void *s) { delete static_cast<RVOSimulator *>(s); }
size_t rvo2ref_add_agent(void *s, float x, float y) {
    return static_cast<RVOSimulator *>(s)->addAgent(Vector2(x, y));
}
void rvo2ref_do_step(void *s) { static_cast<RVOSimulator *>(s)->doStep(); }
void rvo2ref_set_position(void *s, size_t i, float x, float y) {
    static_cast<RVOSimulator *>(s)->setAgentPosition(i, Vector2(x, y));
}
void rvo2ref_set_velocity(void *s, size_t i, float x, float y) {
    static_cast<RVOSimulator *>(s)->setAgentVelocity(i, Vector2(x, y));
}
void rvo2ref_set_pref_velocity(void *s, size_t i, float x, float y) {
    static_cast<RVOSimulator *>(s)->setAgentPrefVelocity(i, Vector2(x, y));
}
void rvo2ref_set_radius(void *s, size_t i, float r) { static_cast<RVOSimulator *>(s)->setAgentRadius(i, r); }
void rvo2ref_set_max_speed(void *s, size_t i, float v) { static_cast<RVOSimulator *>(s)->setAgentMaxSpeed(i, v); }
void rvo2ref_set_collab_coeff(void *s, size_t i, float c) {
    static_cast<RVOSimulator *>(s)->setAgentCollabCoeff(i, c);
}
void rvo2ref_set_neighbor_dist(void *s, size_t i, float d) {
    static_cast<RVOSimulator *>(s)->setAgentNeighborDist(i, d);
}
void rvo2ref_set_max_neighbors(void *s, size_t i, size_t n) {
    static_cast<RVOSimulator *>(s)->setAgentMaxNeighbors(i, n);
}
void rvo2ref_set_time_horizon(void *s, size_t i, float t) {
    static_cast<RVOSimulator *>(s)->setAgentTimeHorizon(i, t);
}
void rvo2ref_set_time_step(void *s, float t) { static_cast<RVOSimulator *>(s)->setTimeStep(t); }
void rvo2ref_get_position(void *s, size_t i, float *out) {
    const Vector2 &p = static_cast<RVOSimulator *>(s)->getAgentPosition(i);
    out[0] = p.x();
    out[1] = p.y();
}
void rvo2ref_get_velocity(void *s, size_t i, float *out) {
    const Vector2 &p = static_cast<RVOSimulator *>(s)->getAgentVelocity(i);
    out[0] = p.x();
    out[1] = p.y();
}
void rvo2ref_get_pref_velocity(void *s, size_t i, float *out) {
    const Vector2 &p = static_cast<RVOSimulator *>(s)->getAgentPrefVelocity(i);
    out[0] = p.x();
    out[1] = p.y();
}
size_t rvo2ref_num_agents(void *s) { return static_cast<RVOSimulator *>(s)->getNumAgents(); }
size_t rvo2ref_num_orca_lines(void *s, size_t i) { return static_cast<RVOSimulator *>(s)->getAgentNumORCALines(i); }
void rvo2ref_get_orca_line(void *s, size_t i, size_t l, float *out) {
    const RVO::Line &ln = static_cast<RVOSimulator *>(s)->getAgentORCALine(i, l);
    out[0] = ln.point.x();
    out[1] = ln.point.y();
    out[2] = ln.direction.x();
    out[3] = ln.direction.y();
}
size_t rvo2ref_num_agent_neighbors(void *s, size_t i) {
    return static_cast<RVOSimulator *>(s)->getAgentNumAgentNeighbors(i);
}
size_t rvo2ref_agent_neighbor(void *s, size_t i, size_t k) {
    return static_cast<RVOSimulator *>(s)->getAgentAgentNeighbor(i, k);
}

/* Batched convenience used by the C-port pinning tests and the CPU baseline:
 * one ego solve == exactly what RVOPolicy.find_next_action does between
 * RVOPolicy.py:57 and :96 (push all N agents, set ego collab coeff, doStep,
 * read ego position).  pos/vel/prefvel are [n][2] float32, radius/maxspeed [n]. */
void rvo2ref_solve_ego(void *s, size_t n, size_t ego, const float *pos, const float *vel, const float *prefvel,
                       const float *radius, const float *maxspeed, float collab, float *new_pos, float *new_vel) {
    RVOSimulator *sim = static_cast<RVOSimulator *>(s);
    while (sim->getNumAgents() < n) sim->addAgent(Vector2(0.0f, 0.0f));
    for (size_t a = 0; a < n; ++a) {
        sim->setAgentMaxSpeed(a, maxspeed[a]);
        sim->setAgentRadius(a, radius[a]);
        sim->setAgentPosition(a, Vector2(pos[2 * a], pos[2 * a + 1]));
        sim->setAgentVelocity(a, Vector2(vel[2 * a], vel[2 * a + 1]));
        sim->setAgentPrefVelocity(a, Vector2(prefvel[2 * a], prefvel[2 * a + 1]));
    }
    sim->setAgentCollabCoeff(ego, collab);
    sim->doStep();
    const Vector2 &p = sim->getAgentPosition(ego);
    const Vector2 &v = sim->getAgentVelocity(ego);
    new_pos[0] = p.x();
    new_pos[1] = p.y();
    new_vel[0] = v.x();
    new_vel[1] = v.y();
}

}  // extern "C"
