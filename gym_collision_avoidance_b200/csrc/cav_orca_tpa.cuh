// cav_orca_tpa.cuh -- one-thread-per-agent ORCA / RVO2 velocity solve in float32 (sm_100a).
//
// Same arithmetic as cav_orca.cuh (the sub-warp-group solver used for large N) and as the reference
// (RVO2 Agent.cpp:295-366 lines, :424-448 neighbour order, :477-630 LP1/LP2/LP3), but one thread owns one ego
// agent and walks its N-1 lines sequentially.  Lines live in a shared-memory array laid out [line][thread]
// (float4 per entry: conflict-free, consecutive threads -> consecutive 16-byte words); LP3's projected lines,
// which are touched only on the rare infeasible solves, live in thread-local memory.
// For small N this mapping needs ~10x fewer issue slots per agent than the group mapping (measured, see
// profiles/ and DESIGN.md): no lane sits idle and nothing is replicated across lanes.
#pragma once
#include "cav_orca.cuh"
#include "cav_orca_core.cuh"  // the solver itself (host/device source, also compiled by the CPU tests)
