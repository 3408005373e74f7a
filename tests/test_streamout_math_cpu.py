"""Index arithmetic of the compile-time-N pairwise+obs kernel's observation stream-out (csrc/cav_tpa_n.cu: TpaNShape,
stream_out), mirrored in Python for every instantiated agent count and both precisions: each staged element is copied
exactly once, from the right staging slot to the right offset of the warp's contiguous HBM block, including the
division-free row index `q_t + [lane >= W - r_t]` and the 16-byte phased float32 copy."""
import pytest

NTS = (2, 3, 4, 5, 6, 8, 10, 12)


def shape(nt, f64):
    no = nt - 1
    epw = 32 // nt
    lw = epw * nt
    cl = 6 + 7 * no
    p = 2 if (f64 and no >= 2) else 1
    ra = no if p == 1 else no // 2
    w0, w1 = 6 + 7 * ra, 7 * (no - ra)
    ws = max(w0, w1) | 1
    flat4 = (not f64) and p == 1 and ws == cl
    return dict(NO=no, EPW=epw, LW=lw, CL=cl, P=p, RA=ra, W0=w0, W1=w1, WS=ws, FLAT4=flat4)


@pytest.mark.parametrize("f64", [False, True])
@pytest.mark.parametrize("nt", NTS)
def test_two_pass_stream_out_covers_every_element_once(nt, f64):
    s = shape(nt, f64)
    assert s["LW"] <= 32 and s["WS"] % 2 == 1 and s["W0"] + (s["W1"] if s["P"] == 2 else 0) == s["CL"]
    written = {}
    passes = [(0, s["W0"])] + ([(s["W0"], s["W1"])] if s["P"] == 2 else [])
    for c0, w in passes:
        tot = s["LW"] * w
        seen_src = set()
        for t in range((tot + 31) // 32):
            for lane in range(32):
                idx = 32 * t + lane
                if w > 32:
                    j = (32 * t) // w + (1 if lane >= w - (32 * t) % w else 0)
                else:
                    j = idx // w
                if idx >= tot:
                    continue
                assert j == idx // w
                k = idx - j * w
                dst = idx + c0 + j * (s["CL"] - w)
                src = idx + j * (s["WS"] - w)
                assert dst == j * s["CL"] + c0 + k and src == j * s["WS"] + k and src not in seen_src
                seen_src.add(src)
                assert dst not in written
                written[dst] = (j, c0 + k)
        assert len(seen_src) == tot
    # the union of the passes is the warp's whole packed block, agent by agent
    assert sorted(written) == list(range(s["LW"] * s["CL"]))
    assert all(written[d] == (d // s["CL"], d % s["CL"]) for d in written)


@pytest.mark.parametrize("nt", NTS)
def test_flat_float32_copy_is_phase_aligned_and_complete(nt):
    s = shape(nt, False)
    if not s["FLAT4"]:
        pytest.skip("float32 at N = %d stages with a padded row stride (generic index path)" % nt)
    tot = s["LW"] * s["CL"]
    for wbase_agents in range(0, 8 * s["LW"], s["LW"]):          # consecutive warps of the grid
        mis = (wbase_agents * s["CL"]) & 3                        # block's offset from 16-byte alignment, in floats
        head = (4 - mis) & 3
        covered = list(range(min(head, 32)))                      # scalar head (lane < head)
        nvec = (tot - head) >> 2
        for v0 in range(0, tot // 4, 32):
            for lane in range(32):
                v = v0 + lane
                if v < nvec:
                    # global float index of the vector start is a multiple of 4 -> 16-byte aligned; the staging slice was
                    # shifted by `mis` floats, so its shared-memory address has the same phase
                    assert (wbase_agents * s["CL"] + head + 4 * v) % 4 == 0 and (mis + head + 4 * v) % 4 == 0
                    covered += [head + 4 * v + q for q in range(4)]
        covered += [head + 4 * nvec + lane for lane in range(32) if head + 4 * nvec + lane < tot]  # scalar tail
        assert sorted(covered) == list(range(tot))
        assert (tot // 4 + 31) // 32 * 32 >= nvec                 # the unrolled vector loop reaches every vector
