"""Packed 64-bit neighbour sort keys (csrc/cav_keys.cuh, host/device source) compiled for the host: an unsigned compare
of two keys must agree with the reference's ordering `(round(dist_2_other, 2), p_orth)` then agent index
(OtherAgentsStatesSensor.py:33-35, Python's stable sort) -- exactly in float32, and in float64 up to the documented
cases the kernels re-compare exactly (equal kint fields)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "gym_collision_avoidance_b200", "csrc")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libkeys_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "native", "keys_host.cpp"), "-I", CSRC, "-o", out])
    return ctypes.CDLL(out)


def _keys(lib, kint, po, idx, dtype, descending=False):
    kint = np.ascontiguousarray(kint, dtype=dtype)
    po = np.ascontiguousarray(po, dtype=dtype)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.zeros(kint.size, np.uint64)
    fn = lib.host_make_keys_f64 if dtype == np.float64 else lib.host_make_keys_f32
    fn(ctypes.c_int(kint.size), kint.ctypes.data_as(ctypes.c_void_p), po.ctypes.data_as(ctypes.c_void_p),
       idx.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(int(descending)), out.ctypes.data_as(ctypes.c_void_p))
    return out


def _samples(rng, n, dtype):
    # rounded distances in centimetres (rint(d * 100)): small range so that ties are frequent; p_orth with exact ties,
    # sign changes, zeros of both signs, tiny and huge magnitudes
    kint = rng.randint(-300, 3000, n).astype(dtype)
    po = rng.uniform(-12, 12, n).astype(dtype)
    po[rng.rand(n) < 0.1] = dtype(0.0)
    po[rng.rand(n) < 0.05] = dtype(-0.0)
    po[rng.rand(n) < 0.05] *= dtype(1e-30)
    po[rng.rand(n) < 0.02] *= dtype(1e20)
    dup = rng.rand(n) < 0.15
    po[dup] = np.roll(po, 1)[dup]
    kint[rng.rand(n) < 0.2] = kint[0]  # a popular value: many equal rounded distances
    idx = rng.randint(0, 64, n)
    return kint, po, idx


@pytest.mark.parametrize("descending", [False, True])
def test_float32_keys_order_exactly_like_the_reference_sort(lib, descending):
    rng = np.random.RandomState(1)
    kint, po, idx = _samples(rng, 200000, np.float32)
    k = _keys(lib, kint, po, idx, np.float32, descending)
    a, b = np.arange(0, k.size, 2), np.arange(1, k.size, 2)
    ka = -kint if descending else kint
    less_ref = (ka[a] < ka[b]) | ((ka[a] == ka[b]) & ((po[a] < po[b]) | ((po[a] == po[b]) & (idx[a] < idx[b]))))
    same = (ka[a] == ka[b]) & (po[a] == po[b]) & (idx[a] == idx[b])
    # +0.0 / -0.0 compare equal in Python but have distinct bit patterns: exclude only those pairs from the exact claim
    zero_pair = (po[a] == 0) & (po[b] == 0) & (np.signbit(po[a]) != np.signbit(po[b]))
    ok = same | zero_pair | ((k[a] < k[b]) == less_ref)
    assert ok.all(), np.nonzero(~ok)[0][:5]
    assert (k[a][same & ~zero_pair] == k[b][same & ~zero_pair]).all()


def test_float64_keys_never_contradict_the_reference_sort(lib):
    rng = np.random.RandomState(2)
    kint, po, idx = _samples(rng, 200000, np.float64)
    k = _keys(lib, kint, po, idx, np.float64)
    kb, pb = ctypes.c_int(), ctypes.c_int()
    lib.host_key_layout(1, ctypes.byref(kb), ctypes.byref(pb))
    shift = np.uint64(pb.value + 6)
    a, b = np.arange(0, k.size, 2), np.arange(1, k.size, 2)
    # different rounded distances: the key order IS the reference order
    d = kint[a] != kint[b]
    assert ((k[a][d] < k[b][d]) == (kint[a][d] < kint[b][d])).all()
    # equal rounded distances: the kint fields are equal (that is what triggers the kernels' exact re-compare) ...
    e = ~d
    assert ((k[a][e] >> shift) == (k[b][e] >> shift)).all()
    # ... and whenever the truncated p_orth fields differ, they already order like the full p_orth values
    pa, pbits = (k[a][e] >> np.uint64(6)) & np.uint64((1 << pb.value) - 1), (k[b][e] >> np.uint64(6)) & np.uint64((1 << pb.value) - 1)
    zero_pair = (po[a][e] == 0) & (po[b][e] == 0)  # +0.0 / -0.0: equal for Python, distinct bit patterns (re-compared exactly)
    dif = (pa != pbits) & ~zero_pair
    assert ((pa[dif] < pbits[dif]) == (po[a][e][dif] < po[b][e][dif])).all()
    # the index field breaks remaining ties in agent order
    tie = (pa == pbits)
    assert (((k[a][e][tie] & np.uint64(63)) < (k[b][e][tie] & np.uint64(63))) == (idx[a][e][tie] < idx[b][e][tie])).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_clamping_and_nan(lib, dtype):
    kb, pb = ctypes.c_int(), ctypes.c_int()
    lib.host_key_layout(int(dtype == np.float64), ctypes.byref(kb), ctypes.byref(pb))
    big = float(1 << (kb.value - 1))
    kint = np.array([-np.inf, -big - 5, -big, -1, 0, 1, big - 1, big + 7, np.inf, np.nan], dtype)
    k = _keys(lib, kint, np.zeros(kint.size, dtype), np.zeros(kint.size, np.int32), dtype)
    fld = k >> np.uint64(pb.value + 6)
    assert (np.diff(fld.astype(np.int64)) >= 0).all()          # monotone, saturating at both ends
    assert fld[0] == fld[1] == fld[2] == 0 and fld[6] == fld[7] == fld[8] == fld[9] == (1 << kb.value) - 1  # NaN sorts last
    assert (k < np.uint64(0xFFFFFFFFFFFFFFFF)).all()           # ~0 is the kernels' "no neighbour" key
