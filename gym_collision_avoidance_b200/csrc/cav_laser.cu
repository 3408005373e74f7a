// cav_laser.cu -- LaserScanSensor raycast for libcavoid_sm100a.so (sm_100a).
//
// Replaces, per environment and step: Map.add_agents_to_map (reference Map.py:46-64: copy the static map and paint
// every agent's disc) and, per agent, LaserScanSensor.sense (LaserScanSensor.py:49-101: 512 beams x 60 range
// samples looked up in the 160x160 occupancy grid, ego disc masked out, "cumsum == 1" range pick, 3-deep history).
//
// Mapping: a CTA owns an ENVIRONMENT.  It rasterises the environment's occupancy grid ONCE into a bit-packed
// shared-memory bitmap (rows x ceil(cols/32) words; 3.2 KB for 160x160) with shared-memory atomicOr, then works through
// the environment's agents in passes of APC agents (all of them in one pass up to 32 agents): per pass it builds, for
// every (ego, other) pair, the other's inflated disc AND the range of beam indices whose rays can touch it, and then
// spreads the pass's APC x 512 (agent, beam) pairs over the CTA.  A beam with no disc in range is 6.0 without any
// trigonometry; otherwise one thread marches the window of samples that can hit, with bit tests out of shared memory and
// an early exit at the second hit.  (ncu on the round-1 kernel, which owned (environment, 4 agents) per CTA:
// profiles/r01g_ncu_laser_f64_1024x20.txt -- 236 M warp instructions = 720 thread instructions per beam, the barrier
// stall on top (8.2 warps per issue) because the bitmap was re-rasterised five times per environment at N = 20 and
// every beam walked every other agent's disc in float64.)
// Geometry is float64 in both precisions (positions / heading are promoted): grid-cell decisions then match the
// reference's float64 cells except where a sample lies within 1e-9 cells of a cell border (exact division path).
// The observation's laserscan segment doubles as the sensor's measurement_history (rows shift down in place).
#include "cav_kernels.h"

namespace cav {

struct LaserPtrs {
    const void *pv, *goal, *ht;
    const int32_t *num_active;
    void *obs;
    uint32_t *laser_count;
    const uint32_t *static_map_bits;
};

// per-pass candidate record of one (ego, other) pair
struct LaserCand {
    double ox, oy;   // other's centre relative to the ego
    double dd, R2;   // |o|^2, inflated radius^2
};

template <typename T, typename OT>
__global__ void __launch_bounds__(256) k_laserscan(LaserPtrs b, LaserParams lp, const uint8_t *__restrict__ env_mask,
                                                     int APC) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_map = reinterpret_cast<uint32_t *>(smem_raw);
    const int map_words = lp.rows * lp.wpr;
    double *s_px = reinterpret_cast<double *>(smem_raw + (((size_t)map_words * 4 + 15) & ~(size_t)15));
    double *s_py = s_px + CAV_MAX_AGENTS;
    double *s_hd = s_py + CAV_MAX_AGENTS;
    double *s_rr2 = s_hd + CAV_MAX_AGENTS;
    int *s_gx = reinterpret_cast<int *>(s_rr2 + CAV_MAX_AGENTS);
    int *s_gy = s_gx + CAV_MAX_AGENTS;
    int *s_in = s_gy + CAV_MAX_AGENTS;
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_in + CAV_MAX_AGENTS);
    int *s_ncand = reinterpret_cast<int *>(s_cnt + CAV_MAX_AGENTS);  // [APC]
    // per agent of the pass: the other agents its beams can reach, compacted in agent order, [APC][N]
    LaserCand *s_cand = reinterpret_cast<LaserCand *>(s_ncand + CAV_MAX_AGENTS);
    short2 *s_brange = reinterpret_cast<short2 *>(s_cand + (size_t)APC * lp.N);  // beams [lo, hi] per candidate
    // per agent of the pass: which beams have ANY candidate in range (bit per beam) -- three quarters of the beams of a
    // 20-agent scene have none and skip the candidate walk altogether
    const int bwords = (lp.B + 31) >> 5;
    uint32_t *s_bany = reinterpret_cast<uint32_t *>(s_brange + (size_t)APC * lp.N);  // [APC][bwords]

    const int tid = threadIdx.x;
    const int N = lp.N;
    const R4<T> *g_pv = reinterpret_cast<const R4<T> *>(b.pv);
    const R4<T> *g_goal = reinterpret_cast<const R4<T> *>(b.goal);
    const R4<T> *g_ht = reinterpret_cast<const R4<T> *>(b.ht);
    OT *g_obs = reinterpret_cast<OT *>(b.obs);  // OT: precision of the packed observation in HBM
    const bool has_static = b.static_map_bits != nullptr;
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;

    for (int e = blockIdx.x; e < lp.E; e += gridDim.x) {
        if (env_mask && !env_mask[e]) continue;  // block-uniform
        const int n = b.num_active ? b.num_active[e] : N;
        const size_t base = (size_t)e * N;
        // ---- 1. static map + agent records ----
        for (int k = tid; k < map_words; k += blockDim.x) s_map[k] = b.static_map_bits ? b.static_map_bits[k] : 0u;
        if (tid < n) {
            const R4<T> pv = ld4(g_pv + base + tid);
            const double px = (double)pv.x, py = (double)pv.y;
            s_px[tid] = px;
            s_py[tid] = py;
            s_hd[tid] = (double)ld4(g_ht + base + tid).x;
            const double rc = (double)ld4(g_goal + base + tid).z / lp.cell;  // Map.py:55 (radius/grid_cell_size)**2
            s_rr2[tid] = rc * rc;
            const double fx = cell_coord(lp.origin0, py, lp.cell, lp.inv_cell, true);
            const double fy = cell_coord(lp.origin1, px, lp.cell, lp.inv_cell, false);
            const bool in = (fx >= 0.0 && fy >= 0.0 && fx < (double)lp.rows && fy < (double)lp.cols);
            s_in[tid] = in ? 1 : 0;
            s_gx[tid] = in ? (int)fx : -1;
            s_gy[tid] = in ? (int)fy : -1;
            s_cnt[tid] = b.laser_count[base + tid];
        }
        __syncthreads();
        // ---- 2. paint discs (Map.py:46-64): cells with (col-gy)^2 + (row-gx)^2 < (r/cell)^2 ----
        for (int a = warp; a < n; a += nwarps) {
            if (!s_in[a]) continue;  // off-map agents paint nothing (Map.py:58-64)
            const double rr2 = s_rr2[a];
            double rcd = sqrt(rr2);
            if (!(rcd < (double)(lp.rows + lp.cols))) rcd = (double)(lp.rows + lp.cols);
            const int R = (int)rcd + 1;
            const int gr = s_gx[a], gc = s_gy[a];
            const int r0 = max(0, gr - R), r1 = min(lp.rows - 1, gr + R);
            const int c0 = max(0, gc - R), c1 = min(lp.cols - 1, gc + R);
            const int wbox = c1 - c0 + 1, cells = wbox * (r1 - r0 + 1);
            for (int k = lane; k < cells; k += 32) {
                const int r = r0 + k / wbox, c = c0 + k % wbox;
                const double dc = (double)(c - gc), dr = (double)(r - gr);
                if (dc * dc + dr * dr < rr2) atomicOr(&s_map[r * lp.wpr + (c >> 5)], 1u << (c & 31));
            }
        }
        __syncthreads();
        for (int a0 = 0; a0 < n; a0 += APC) {  // passes over the environment's agents (one pass up to 32 agents)
            const int a_end = min(n, a0 + APC);
            // ---- 3a. which samples can hit at all?  A sample can only land in a cell painted for agent c if it lies
            // within R_c = r_c + sqrt(2)*cell (+eps) of c's centre (the sample and the centre are each anywhere inside
            // their cells), so with an empty static map only samples inside the union of the ray's intersections with
            // those inflated discs need the exact cell test.  Candidates (painted agents within max_range + R_c) are
            // compacted once per agent of the pass -- one warp per agent, ballot compaction keeps agent order -- together
            // with the range of BEAM INDICES whose rays can touch the disc: seen from the ego the disc subtends
            // +-asin(R/D) around its bearing, the beams are a regular fan relative to the heading, so the range is two
            // divisions (widened by a beam and 2e-3 rad on each side; float32 trigonometry is ample for that).  A beam
            // then looks at a candidate's float64 geometry only if its index is in range.
            if (!has_static) {
                const float inv_step = lp.ang_step > 0.0 ? (float)(1.0 / lp.ang_step) : 0.f;
                for (int k = tid; k < (a_end - a0) * bwords; k += blockDim.x) s_bany[k] = 0u;
                __syncthreads();
                for (int q = warp; q < a_end - a0; q += nwarps) {
                    const int h = a0 + q;
                    const double px = s_px[h], py = s_py[h];
                    const float hd = (float)s_hd[h];
                    int nc = 0;
                    for (int c0 = 0; c0 < n; c0 += 32) {
                        const int c = c0 + lane;
                        bool keep = false;
                        double ox = 0.0, oy = 0.0, dd = 0.0, R = 0.0;
                        int blo = 0, bhi = lp.B - 1;
                        if (c < n && c != h && s_in[c]) {
                            ox = s_px[c] - px;
                            oy = s_py[c] - py;
                            R = sqrt(s_rr2[c]) * lp.cell + 1.4143 * lp.cell + 1e-6;
                            dd = ox * ox + oy * oy;
                            const double reach = lp.max_range + R;
                            keep = !(dd > reach * reach);
                            if (keep && dd > R * R && lp.B > 1) {
                                const float D = sqrtf((float)dd), Rf = (float)R;
                                const float wdt = asinf(fminf(1.0f, Rf / D * 1.0001f)) + 2e-3f;
                                float phi = atan2f((float)oy, (float)ox) - hd;  // bearing relative to the heading
                                phi -= 6.2831853f * rintf(phi * 0.15915494f);   // into [-pi, pi]
                                const float amin = (float)lp.min_angle, amax = (float)lp.max_angle;
                                int pieces = 0;
                                float lo = 0.f, hi = 0.f;
#pragma unroll
                                for (int im = -1; im <= 1; ++im) {  // the arc and its +-2 pi images against the fan
                                    const float l = phi + 6.2831853f * (float)im - wdt, r = phi + 6.2831853f * (float)im + wdt;
                                    if (r >= amin - 1e-3f && l <= amax + 1e-3f) {
                                        ++pieces;
                                        lo = l;
                                        hi = r;
                                    }
                                }
                                if (pieces == 0) {
                                    keep = false;  // no beam of the fan points at this disc
                                } else if (pieces == 1) {
                                    blo = max(0, (int)floorf((lo - amin) * inv_step) - 1);
                                    bhi = min(lp.B - 1, (int)ceilf((hi - amin) * inv_step) + 1);
                                }  // two pieces (arc across both ends of the fan): every beam stays in range
                            }
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, keep);
                        if (keep) {
                            const int pos = q * N + nc + __popc(m & ((1u << lane) - 1u));
                            s_cand[pos] = LaserCand{ox, oy, dd, R * R};
                            s_brange[pos] = make_short2((short)blo, (short)bhi);
                            for (int w = blo >> 5; w <= (bhi >> 5); ++w) {  // mark beams [blo, bhi]
                                const int lo = max(blo, w << 5) & 31, hi = min(bhi, (w << 5) + 31) & 31;
                                atomicOr(&s_bany[q * bwords + w], (hi == 31 ? ~0u : ((2u << hi) - 1u)) & ~((1u << lo) - 1u));
                            }
                        }
                        nc += __popc(m);
                    }
                    if (lane == 0) s_ncand[q] = nc;
                }
                __syncthreads();
            }
            // ---- 3b. beams: the pass's (agent, beam) pairs flattened over the CTA, no barrier between agents ----
            const int ntask = (a_end - a0) * lp.B;
            const int bshift = (lp.B & (lp.B - 1)) == 0 ? 31 - __clz(lp.B) : -1;  // 512 beams: shift instead of divide
            for (int task = tid; task < ntask; task += blockDim.x) {
                const int q = bshift >= 0 ? (task >> bshift) : task / lp.B, bi = task - q * lp.B;
                const int h = a0 + q;
                const uint32_t cnt = s_cnt[h];
                OT *o = g_obs + (base + h) * (size_t)lp.obs_len + lp.obs_off;
                // measurement history rows that will shift down: fetch them now, their latency hides behind the march
                OT older[3] = {(OT)0, (OT)0, (OT)0};
                const bool prefetched = cnt != 0u && lp.P <= 4;
                if (prefetched) {
#pragma unroll
                    for (int pp = 0; pp < 3; ++pp)
                        if (pp < lp.P - 1) older[pp] = o[(size_t)pp * lp.B + bi];
                }
                const int ncand = has_static ? 0 : s_ncand[q];
                const LaserCand *cand = s_cand + q * N;
                const short2 *brange = s_brange + q * N;
                // candidates whose beam range holds this beam, as a bit mask over the compacted list (<= 63 entries)
                unsigned long long inr = 0ull;
                const bool any = has_static ? false : ((s_bany[q * bwords + (bi >> 5)] >> (bi & 31)) & 1u) != 0u;
#pragma unroll 1
                for (int c = 0; any && c < ncand; ++c) {
                    const short2 br = brange[c];
                    if (bi >= br.x && bi <= br.y) inr |= 1ull << c;
                }
                int kstar = -1;
                if (has_static || inr != 0ull) {
                    const double px = s_px[h], py = s_py[h], hd = s_hd[h], rr2 = s_rr2[h];
                    const int ego_in = s_in[h], cgx = s_gx[h], cgy = s_gy[h];
                    double ang0 = (double)bi * lp.ang_step + lp.min_angle;  // np.linspace (LaserScanSensor.py:38)
                    if (bi == lp.B - 1) ang0 = lp.max_angle;
                    const double ang = ang0 + hd;  // :66
                    double sa, ca;
                    sincos(ang, &sa, &ca);
                    // Samples worth testing, as a bit mask over the (<= 64) range samples: with a static map every sample;
                    // otherwise only those inside some candidate's inflated disc -- the ray / disc chord [t0, t1] widened
                    // by one sample on each side.  Everything between two discs is free space by construction, so the
                    // march jumps over it (round 1 marched ONE window from the first disc's entry to the last one's exit).
                    unsigned long long smask;
                    const bool masked = !has_static && lp.nr <= 64;
                    if (!masked) {
                        smask = ~0ull;
                    } else {
                        smask = 0ull;
                        const double inv_res = 1.0 / lp.range_res;
                        while (inr != 0ull) {
                            const int c = __ffsll((long long)inr) - 1;
                            inr &= inr - 1ull;
                            const LaserCand cd = cand[c];
                            const double tca = cd.ox * ca + cd.oy * sa;
                            const double d2 = cd.dd - tca * tca;
                            if (d2 > cd.R2) continue;
                            const double thc = sqrt(cd.R2 - d2);
                            const double t1 = tca + thc;
                            if (t1 < 0.0) continue;
                            const double t0 = tca - thc;
                            const double a = floor((t0 - lp.min_range) * inv_res) - 1.0;
                            const double bnd = ceil((t1 - lp.min_range) * inv_res) + 1.0;
                            const int k0 = a > 0.0 ? (int)a : 0;
                            const int k1 = bnd < (double)(lp.nr - 1) ? (int)bnd : lp.nr - 1;
                            if (k1 >= k0) smask |= ((k1 - k0 >= 63) ? ~0ull : ((1ull << (k1 - k0 + 1)) - 1ull)) << k0;
                        }
                    }
                    int first = -1;
                    bool second = false;
                    int k = 0;
#pragma unroll 1
                    while (masked ? (smask != 0ull) : (k < lp.nr)) {
                        if (masked) {
                            k = __ffsll((long long)smask) - 1;
                            smask &= smask - 1ull;
                        }
                        const double rk = lp.min_range + (double)k * lp.range_res;  // np.arange (:39)
                        const double x = px + rk * ca, y = py + rk * sa;            // :70-72
                        const double fx = cell_coord(lp.origin0, y, lp.cell, lp.inv_cell, true);
                        const double fy = cell_coord(lp.origin1, x, lp.cell, lp.inv_cell, false);
                        int hit = 0;
                        if (fx >= 0.0 && fy >= 0.0 && fx < (double)lp.rows && fy < (double)lp.cols) {
                            const int gx = (int)fx, gy = (int)fy;
                            hit = (s_map[gx * lp.wpr + (gy >> 5)] >> (gy & 31)) & 1u;
                            if (hit && ego_in) {  // ego mask (:76-77)
                                const double dc = (double)(gy - cgy), dr = (double)(gx - cgx);
                                if (dc * dc + dr * dr < rr2) hit = 0;
                            }
                        }
                        // :78-82: the range is the LAST index whose cumulative hit count is 1 -- the sample just before the
                        // second hit (skipped samples are misses and do not change the count), or the last sample of the beam
                        if (hit) {
                            if (first < 0) {
                                first = k;
                            } else {
                                kstar = k - 1;
                                second = true;
                                break;
                            }
                        }
                        if (!masked) ++k;
                    }
                    if (first >= 0 && !second) kstar = lp.nr - 1;
                }
                const double rng = kstar >= 0 ? lp.min_range + (double)kstar * lp.range_res : lp.max_range;
                if (cnt == 0u) {  // first measurement fills every row (:84-85)
                    for (int pp = 0; pp < lp.P; ++pp) o[(size_t)pp * lp.B + bi] = (OT)rng;
                } else if (prefetched) {  // np.roll(history, 1, axis=0); row 0 = newest (:86-88)
#pragma unroll
                    for (int pp = 3; pp >= 1; --pp)
                        if (pp < lp.P) o[(size_t)pp * lp.B + bi] = older[pp - 1];
                    o[bi] = (OT)rng;
                } else {
                    for (int pp = lp.P - 1; pp >= 1; --pp) o[(size_t)pp * lp.B + bi] = o[(size_t)(pp - 1) * lp.B + bi];
                    o[bi] = (OT)rng;
                }
            }
            __syncthreads();  // the candidate lists are rebuilt by the next pass
        }
        if (tid < n) b.laser_count[base + tid] = s_cnt[tid] + 1u;
        __syncthreads();
    }
}

template <typename T, typename OT>
cudaError_t run_laserscan(const cav_buffers_t &bufs, const LaserParams &lp, const uint8_t *env_mask, int sm_count,
                          cudaStream_t st) {
    LaserPtrs p;
    p.pv = bufs.pv;
    p.goal = bufs.goal;
    p.ht = bufs.ht;
    p.num_active = bufs.num_active;
    p.obs = bufs.obs;
    p.laser_count = bufs.laser_count;
    p.static_map_bits = bufs.static_map_bits;
    const int APC = lp.N < 32 ? lp.N : 32;  // agents per pass: candidate records [APC][N] x 36 B <= 72 KB
    const size_t smem = (((size_t)lp.rows * lp.wpr * 4 + 15) & ~(size_t)15) + CAV_MAX_AGENTS * (4 * 8 + 5 * 4) +
                        (size_t)APC * lp.N * (sizeof(LaserCand) + sizeof(short2)) + (size_t)APC * ((lp.B + 31) / 32) * 4 + 64;
    auto kern = k_laserscan<T, OT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int grid = lp.E < sm_count * 16 ? lp.E : sm_count * 16;
    kern<<<grid, 256, smem, st>>>(p, lp, env_mask, APC);
    return cudaGetLastError();
}

template cudaError_t run_laserscan<float, float>(const cav_buffers_t &, const LaserParams &, const uint8_t *, int,
                                                 cudaStream_t);
template cudaError_t run_laserscan<double, double>(const cav_buffers_t &, const LaserParams &, const uint8_t *, int,
                                                   cudaStream_t);
template cudaError_t run_laserscan<double, float>(const cav_buffers_t &, const LaserParams &, const uint8_t *, int,
                                                  cudaStream_t);

}  // namespace cav
