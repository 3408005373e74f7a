// cav_kdtree.cuh -- RVO2's agent kd-tree as far as it decides RESULTS (host/device source).
//
// The reference's neighbour list is ordered by ascending float32 distSq with ties in INSERTION order (RVO2
// Agent.cpp:424-448), and agents are inserted in the order the kd-tree query visits them: KdTree.cpp:266-298, depth
// first, the child whose bounding box is nearer to the ego first, leaves front to back.  Leaves are ranges of
// KdTree::agents_, a permutation that buildAgentTree (KdTree.cpp:47-111) re-partitions IN PLACE on every doStep and
// never resets -- it persists for the life of the simulator (one per RVOPolicy, i.e. per agent and episode; every
// simulator of an environment is fed the same positions on the same steps, so one permutation per environment stands
// for all of them: cavoid.h `kd_order`).  With at most MAX_LEAF_SIZE = 10 agents (KdTree.h:48) the root is a leaf and
// the order is the agent index; above that, exact distance ties (the reference's symmetric circle / swap presets) are
// resolved by this structure, so it is reproduced here operation for operation:
//   kd_build     one thread: bounding boxes, split axis / value, the Hoare-style partition with its swaps, and the
//                "left == begin" special case, over the persistent permutation
//   kd_traverse  any thread: calls f(agent) in the reference's insertion order for an ego at (px, py)
// Pruning by rangeSq (KdTree.cpp:277-294) is omitted: a pruned subtree holds only agents whose distSq >= rangeSq, which
// insertAgentNeighbor rejects anyway, so visiting them cannot change the list.
// g++ compiles this file into tests/native/orca_host.cpp, where it is checked against the reference library itself.
#pragma once
#include "cav_hd.cuh"

namespace cav {

#define CAV_KD_MAX_LEAF 10  // RVO2 KdTree.h:48 MAX_LEAF_SIZE

struct KdMeta {
    uint8_t begin, end, left, right;  // agents_[begin, end); children node ids (internal nodes only)
};

// Views into caller-provided storage (shared memory in the kernels).  box / meta hold 2 * n - 1 slots: the reference
// numbers nodes in pre-order with gaps (right = node + 2 * (left - begin), KdTree.cpp:105), unused slots have end == 0.
struct KdTree {
    float4 *box;    // (minX, maxX, minY, maxY)
    KdMeta *meta;
    uint8_t *perm;  // KdTree::agents_ as agent indices
};

CAV_HDI int kd_slots(int n) { return 2 * n - 1; }

// KdTree::buildAgentTree / buildAgentTreeRecursive (KdTree.cpp:47-111).  pos[i].x / .y = position of agent i.
// A node's range is written by its parent, and children have larger ids than their parent, so one pass over the slots
// in id order visits every node after its parent: no recursion, no stack.
CAV_HDI void kd_build(const float4 *pos, int n, KdTree t) {
    const int slots = kd_slots(n);
    for (int s = 1; s < slots; ++s) t.meta[s] = KdMeta{0, 0, 0, 0};
    t.meta[0] = KdMeta{0, (uint8_t)n, 0, 0};
    for (int node = 0; node < slots; ++node) {
        KdMeta m = t.meta[node];
        if (m.end == 0) continue;
        const int begin = m.begin, end = m.end;
        float4 p0 = pos[t.perm[begin]];
        float minX = p0.x, maxX = p0.x, minY = p0.y, maxY = p0.y;
        for (int i = begin + 1; i < end; ++i) {
            const float4 pi = pos[t.perm[i]];
            maxX = (maxX < pi.x) ? pi.x : maxX;  // std::max(a, b) = (a < b) ? b : a
            minX = (pi.x < minX) ? pi.x : minX;  // std::min(a, b) = (b < a) ? b : a
            maxY = (maxY < pi.y) ? pi.y : maxY;
            minY = (pi.y < minY) ? pi.y : minY;
        }
        t.box[node] = make_float4(minX, maxX, minY, maxY);
        if (end - begin > CAV_KD_MAX_LEAF) {
            const bool isVertical = (maxX - minX > maxY - minY);
            const float splitValue = isVertical ? 0.5f * (maxX + minX) : 0.5f * (maxY + minY);
            int left = begin, right = end;
            while (left < right) {
                while (left < right) {
                    const float4 q = pos[t.perm[left]];
                    if (!((isVertical ? q.x : q.y) < splitValue)) break;
                    ++left;
                }
                while (right > left) {
                    const float4 q = pos[t.perm[right - 1]];
                    if (!((isVertical ? q.x : q.y) >= splitValue)) break;
                    --right;
                }
                if (left < right) {
                    const uint8_t tmp = t.perm[left];
                    t.perm[left] = t.perm[right - 1];
                    t.perm[right - 1] = tmp;
                    ++left;
                    --right;
                }
            }
            if (left == begin) {
                ++left;
                ++right;
            }
            m.left = (uint8_t)(node + 1);
            m.right = (uint8_t)(node + 2 * (left - begin));
            t.meta[node] = m;
            t.meta[m.left] = KdMeta{(uint8_t)begin, (uint8_t)left, 0, 0};
            t.meta[m.right] = KdMeta{(uint8_t)left, (uint8_t)end, 0, 0};
        }
    }
}

CAV_HDI float kd_sqr(float a) { return a * a; }
CAV_HDI float kd_max0(float a) { return (0.0f < a) ? a : 0.0f; }  // std::max(0.0f, a)
// squared distance from (px, py) to a bounding box, term order of KdTree.cpp:274 / :276
CAV_HDI float kd_box_dist_sq(const float4 &b, float px, float py) {
    return kd_sqr(kd_max0(b.x - px)) + kd_sqr(kd_max0(px - b.y)) + kd_sqr(kd_max0(b.z - py)) + kd_sqr(kd_max0(py - b.w));
}

// KdTree::queryAgentTreeRecursive (KdTree.cpp:266-298) without the range pruning (see the header note): f(agent index)
// for every agent, the ego included (insertAgentNeighbor skips it, Agent.cpp:426), in the reference's visiting order.
template <typename F>
CAV_HDI void kd_traverse(const KdTree &t, float px, float py, F &&f) {
    uint8_t stack[CAV_MAX_AGENTS];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const int node = stack[--sp];
        const KdMeta m = t.meta[node];
        if ((int)m.end - (int)m.begin <= CAV_KD_MAX_LEAF) {
            for (int i = m.begin; i < m.end; ++i) f((int)t.perm[i]);
        } else {
            const float distSqLeft = kd_box_dist_sq(t.box[m.left], px, py);
            const float distSqRight = kd_box_dist_sq(t.box[m.right], px, py);
            if (distSqLeft < distSqRight) {  // left first
                stack[sp++] = m.right;
                stack[sp++] = m.left;
            } else {
                stack[sp++] = m.left;
                stack[sp++] = m.right;
            }
        }
    }
}

}  // namespace cav
