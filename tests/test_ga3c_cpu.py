"""GA3C-CADRL policy network, CPU side (reference policies/GA3CCADRLPolicy.py:49-84,
policies/GA3C_CADRL/network.py): the C restatement of the forward pass against the golden vectors that oracle/tf1_graph.py
minted by EXECUTING the reference's meta-graph with the IROS18 checkpoint (tools/make_ga3c_golden.py), the TensorFlow-free
readers of the reference's files, the observation -> network-input gather, and the policy inside the oracle step.

Tolerance: 1e-5 on the softmax (float32 network; the restatement sums in ascending k, numpy's matmul does not), the
chosen action identical wherever the two best probabilities are more than 1e-4 apart."""
import os

import numpy as np
import pytest

import cav_oracle
from cav_oracle import Oracle
from gym_collision_avoidance_b200 import abi
from gym_collision_avoidance_b200.envs.config import make_config, to_cav_config
from gym_collision_avoidance_b200.envs.policies.GA3C_CADRL import network, tf1_files
from gym_collision_avoidance_b200.envs.vec_env import kinds_array
from gym_collision_avoidance_b200.state import HostBank, HostState

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CKPT = "/root/reference/gym_collision_avoidance/envs/policies/GA3C_CADRL/checkpoints/"
needs_reference = pytest.mark.skipif(not os.path.isdir(REF_CKPT), reason="the reference's checkpoints are not on this machine")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "ga3c", "ga3c_iros18.npz"))


def check_probs(got_p, got_a, want_p):
    assert np.abs(got_p - want_p).max() <= 1e-5
    top2 = np.sort(want_p, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-4
    assert clear.sum() > 0.9 * len(clear)
    assert np.array_equal(got_a[clear], want_p.argmax(1)[clear])


def test_forward_matches_the_executed_reference_graph(golden):
    p, a = cav_oracle.ga3c_forward(golden["blob"], golden["X"])
    check_probs(p, a, golden["probs"])
    assert np.allclose(p.sum(1), 1.0, atol=1e-6)
    assert len(np.unique(golden["probs"].argmax(1))) >= 7, "the fixture must exercise most of the action table"


def test_sequence_length_semantics(golden):
    """dynamic_rnn: only the first int(X[:, 0]) rows are consumed (stale rows behind them do not matter), at most 19."""
    blob, X = golden["blob"], golden["X"].copy()
    rng = np.random.RandomState(3)
    Y = X.copy()
    for r in range(len(Y)):
        m = int(min(Y[r, 0], 19))
        Y[r, 5 + 7 * m:] = rng.uniform(-9, 9, 138 - 5 - 7 * m)
    p0, _ = cav_oracle.ga3c_forward(blob, X)
    p1, _ = cav_oracle.ga3c_forward(blob, Y)
    assert np.array_equal(p0, p1)
    Z = X[X[:, 0] >= 19].copy()
    Z19 = Z.copy()
    Z19[:, 0] = 19
    pa, _ = cav_oracle.ga3c_forward(blob, Z)
    pb, _ = cav_oracle.ga3c_forward(blob, Z19)
    assert np.abs(pa - pb).max() <= 1e-6  # (the length itself is also input 0 of the normalised vector, unused downstream)
    frac = X.copy()
    frac[:, 0] += 0.75  # ToInt32 truncates
    pf, _ = cav_oracle.ga3c_forward(blob, frac)
    assert np.array_equal(pf, p0)


def test_action_table_is_the_references(golden):
    a = network.Actions()
    assert a.num_actions == 11
    assert np.array_equal(a.actions, golden["actions"]), "bit-identical to np.mgrid's table (network.py:12-16)"


def test_crc32c_known_answers():
    assert tf1_files.crc32c(b"123456789") == 0xE3069283
    assert tf1_files.crc32c(b"") == 0
    assert tf1_files.crc32c(bytes(32)) == 0x8A9136AA


def test_checkpoint_reader_rejects_garbage(tmp_path):
    p = tmp_path / "bad"
    (tmp_path / "bad.index").write_bytes(b"\x00" * 64)
    (tmp_path / "bad.data-00000-of-00001").write_bytes(b"")
    with pytest.raises(ValueError):
        tf1_files.read_checkpoint(str(p))


def test_pack_blob_layout(golden):
    """Sections of the blob against a hand-built variable set: the LSTM kernel's [71, 4*64] columns regroup to [k][unit][gate]."""
    rng = np.random.RandomState(0)
    randn = rng.randn
    rng = type("R", (), {"randn": staticmethod(lambda *s: 0.08 * randn(*s)), "rand": rng.rand})  # trained-weight magnitudes
    v = {"rnn/lstm_cell/kernel:0": rng.randn(71, 256).astype(np.float32), "rnn/lstm_cell/bias:0": rng.randn(256).astype(np.float32),
         "layer1/kernel:0": rng.randn(68, 256).astype(np.float32), "layer1/bias:0": rng.randn(256).astype(np.float32),
         "layer2/kernel:0": rng.randn(256, 256).astype(np.float32), "layer2/bias:0": rng.randn(256).astype(np.float32),
         "fullyconnected1/kernel:0": rng.randn(256, 256).astype(np.float32), "fullyconnected1/bias:0": rng.randn(256).astype(np.float32),
         "logits_p/kernel:0": rng.randn(256, 11).astype(np.float32), "logits_p/bias:0": rng.randn(11).astype(np.float32)}
    avg, std = rng.randn(138).astype(np.float32), (1 + rng.rand(138)).astype(np.float32)
    blob = network.pack_blob(v, avg, std)
    assert blob.size == abi.CAV_GA3C_BLOB_FLOATS == golden["blob"].size
    wl = blob[276:276 + 71 * 256].reshape(71, 64, 4)
    for gate in range(4):
        assert np.array_equal(wl[:, :, gate], v["rnn/lstm_cell/kernel:0"][:, gate * 64:(gate + 1) * 64])
    assert np.array_equal(blob[276 + 71 * 256:276 + 71 * 256 + 256].reshape(64, 4)[:, 2], v["rnn/lstm_cell/bias:0"][128:192])
    assert np.array_equal(blob[-12:-1], v["logits_p/bias:0"]) and blob[-1] == 0
    # a plain numpy forward pass from the SAME variables agrees with the C restatement reading the blob
    X = golden["X"][:40]
    xn = (X - avg) / std
    h = np.zeros((len(X), 64), np.float32)
    c = np.zeros_like(h)
    sig = lambda z: 1 / (1 + np.exp(-z))  # noqa: E731
    for t in range(19):
        g = np.concatenate([xn[:, 5 + 7 * t:12 + 7 * t], h], 1) @ v["rnn/lstm_cell/kernel:0"] + v["rnn/lstm_cell/bias:0"]
        i, j, f, o = np.split(g, 4, axis=1)
        cn = sig(f + 1) * c + sig(i) * np.tanh(j)
        hn = sig(o) * np.tanh(cn)
        on = (t < X[:, 0].astype(np.int32))[:, None]
        c, h = np.where(on, cn, c), np.where(on, hn, h)
    a = np.maximum(np.concatenate([xn[:, 1:5], h], 1) @ v["layer1/kernel:0"] + v["layer1/bias:0"], 0)
    a = np.maximum(a @ v["layer2/kernel:0"] + v["layer2/bias:0"], 0)
    a = np.maximum(a @ v["fullyconnected1/kernel:0"] + v["fullyconnected1/bias:0"], 0)
    lg = a @ v["logits_p/kernel:0"] + v["logits_p/bias:0"]
    e = np.exp(lg - lg.max(1, keepdims=True))
    p, _ = cav_oracle.ga3c_forward(blob, X)
    assert np.abs(p - e / e.sum(1, keepdims=True)).max() <= 1e-5


@needs_reference
def test_loader_reads_the_references_checkpoint(golden):
    blob = network.load_blob(REF_CKPT + "IROS18/network_01900000")
    assert np.array_equal(blob, golden["blob"])


@needs_reference
@pytest.mark.parametrize("ckpt", ["run-20190727_192048-qedrf08y/network_01900000", "run-20190727_015942-jzuhlntn/network_01490000"])
def test_other_shipped_checkpoints_against_their_own_graph(ckpt, golden):
    """The reference ships two more checkpoints with the same graph: each one's blob, through the C restatement, against
    its OWN meta-graph executed by the interpreter."""
    import tf1_graph
    it = tf1_graph.load(REF_CKPT + ckpt)
    X = golden["X"][:92]
    want = it.run("Softmax", {"X": X})
    p, a = cav_oracle.ga3c_forward(network.load_blob(REF_CKPT + ckpt), X)
    check_probs(p, a, want)


def test_rejects_a_different_graph():
    nodes = {"X": {"op": "Placeholder", "inputs": [], "attrs": {"shape": ("shape", [-1, 26])}}}
    with pytest.raises(abi.CavoidError):
        network._check_graph(nodes)


@pytest.mark.parametrize("N", [2, 10, 24])
def test_observation_gather(N):
    """The network input is the flattened observation without is_learning, cut / zero-padded to 138 (network.py:24-35)."""
    c = to_cav_config(make_config(N, evaluate=True), 1, N, precision="f64")
    src = cav_oracle.ga3c_gather(c)
    K = N - 1
    want = [c.obs_offset[abi.CAV_OBS_NUM_OTHER_AGENTS], c.obs_offset[abi.CAV_OBS_DIST_TO_GOAL],
            c.obs_offset[abi.CAV_OBS_HEADING_EGO_FRAME], c.obs_offset[abi.CAV_OBS_PREF_SPEED], c.obs_offset[abi.CAV_OBS_RADIUS]]
    o = c.obs_offset[abi.CAV_OBS_OTHER_AGENTS_STATES]
    want += list(range(o, o + 7 * min(K, 19)))
    want += [-1] * (138 - len(want))
    assert list(src) == want


@pytest.mark.parametrize("prec,obs_f32", [("f64", False), ("f64", True), ("f32", False)])
def test_oracle_step_with_ga3c_agents(golden, prec, obs_f32):
    """Two GA3C-CADRL agents swapping places (the reference's example pairs this policy with itself): both reach their goals,
    nobody collides; and in a mixed environment the RVO agents are untouched by the network being present."""
    cav_oracle.set_ga3c(golden["blob"])
    try:
        N = 2
        c = to_cav_config(make_config(N, evaluate=True), 1, N, precision=prec, track_aux=True, obs_f32=obs_f32)
        rows = np.array([[[-3.0, 0.0, 3.0, 0.0, 1.0, 0.5, 0.0], [3.0, 0.0, -3.0, 0.0, 1.0, 0.5, np.pi]]])
        bank = HostBank(rows, kinds=kinds_array(1, N, "GA3C_CADRL"))
        hs = HostState(c)
        orc = Oracle(c)
        orc.reset(hs, bank)
        seen = set()
        for _ in range(150):
            orc.step(hs)
            seen.add(tuple(np.round(hs.actions[0, 0], 4)))
            if hs.game_over[0]:
                break
        f = hs.flags[0]
        assert all(int(x) & abi.CAV_FLAG_AT_GOAL for x in f), f
        assert not any(int(x) & abi.CAV_FLAG_IN_COLLISION for x in f)
        assert len(seen) >= 2, "the agents must have turned to pass each other"
    finally:
        cav_oracle.set_ga3c(None)


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _pb(field, wire, payload):
    key = _varint((field << 3) | wire)
    if wire == 0:
        return key + _varint(payload)
    if wire == 2:
        return key + _varint(len(payload)) + payload
    return key + payload  # fixed32 / fixed64


def _write_bundle(prefix, tensors, corrupt=None):
    """A minimal single-shard TF tensor bundle (uncompressed one-block SSTable index + raw data), written by hand so that
    the reader is tested wherever the reference's checkpoints are absent."""
    import struct
    data = b""
    entries = [(b"", b"")]  # BundleHeaderProto under the empty key (contents irrelevant to the reader)
    for name in sorted(tensors):
        arr = np.asarray(tensors[name])
        raw = arr.tobytes()
        crc = tf1_files._masked_crc(raw)
        if corrupt == name:
            raw = raw[:-1] + bytes([raw[-1] ^ 0xFF])
        shape = b"".join(_pb(2, 2, _pb(1, 0, int(d))) for d in arr.shape)
        dt = {np.dtype("float32"): 1, np.dtype("int32"): 3}[arr.dtype]
        entry = _pb(1, 0, dt) + _pb(2, 2, shape) + _pb(4, 0, len(data)) + _pb(5, 0, len(raw)) + _pb(6, 5, struct.pack("<I", crc))
        entries.append((name.encode(), entry))
        data += raw

    def block(kvs):
        body, prev = b"", b""
        for k, v in kvs:
            shared = 0
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
            body += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
            prev = k
        return body + struct.pack("<II", 0, 1)  # one restart point at offset 0, num_restarts = 1
    blocks = b""
    dblk = block(entries)
    d_off, blocks = len(blocks), blocks + dblk + b"\x00" + b"\x00" * 4  # trailer: no compression + (unchecked) block crc
    mblk = block([])
    m_off, blocks = len(blocks), blocks + mblk + b"\x00" + b"\x00" * 4
    iblk = block([(b"\xff", _varint(d_off) + _varint(len(dblk)))])
    i_off, blocks = len(blocks), blocks + iblk + b"\x00" + b"\x00" * 4
    foot = _varint(m_off) + _varint(len(mblk)) + _varint(i_off) + _varint(len(iblk))
    foot += b"\x00" * (40 - len(foot)) + struct.pack("<Q", 0xdb4775248b80fb57)
    with open(prefix + ".index", "wb") as fh:
        fh.write(blocks + foot)
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(data)


def test_bundle_reader_on_a_hand_written_bundle(tmp_path):
    """Prefix-compressed keys, the footer's block handles, dtype / shape / offset / size fields and the CRC check of the
    tensor-bundle reader, on a bundle built by this test."""
    rng = np.random.RandomState(4)
    tensors = {"layer1/bias:0": rng.randn(256).astype(np.float32), "layer1/kernel:0": rng.randn(68, 256).astype(np.float32),
               "layer2/kernel:0": rng.randn(4, 3, 2).astype(np.float32), "step:0": np.array(123, np.int32)}
    p = str(tmp_path / "net")
    _write_bundle(p, tensors)
    got = tf1_files.read_checkpoint(p)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)
    _write_bundle(p, tensors, corrupt="layer1/kernel:0")
    with pytest.raises(ValueError, match="CRC"):
        tf1_files.read_checkpoint(p)
    assert np.array_equal(tf1_files.read_checkpoint(p, verify=False)["layer1/bias:0"], tensors["layer1/bias:0"])


def test_meta_graph_reader_on_a_hand_written_graph(tmp_path):
    """NodeDef name / op / inputs / attrs (shape, tensor with packed floats, tensor_content, scalar int) from protobuf bytes."""
    import struct
    shape = _pb(2, 2, _pb(1, 0, (1 << 64) - 1)) + _pb(2, 2, _pb(1, 0, 138))               # [-1, 138]
    ph = _pb(1, 2, b"X") + _pb(2, 2, b"Placeholder") + _pb(5, 2, _pb(1, 2, b"shape") + _pb(2, 2, _pb(7, 2, shape)))
    tconst = _pb(1, 0, 1) + _pb(2, 2, _pb(2, 2, _pb(1, 0, 3))) + _pb(5, 2, struct.pack("<3f", 1.0, 5.0, 3.14))
    const = _pb(1, 2, b"Const") + _pb(2, 2, b"Const") + _pb(5, 2, _pb(1, 2, b"value") + _pb(2, 2, _pb(8, 2, tconst)))
    tcont = _pb(1, 0, 3) + _pb(2, 2, _pb(2, 2, _pb(1, 0, 2))) + _pb(4, 2, struct.pack("<2i", 0, 5))
    stack = _pb(1, 2, b"s/stack") + _pb(2, 2, b"Const") + _pb(5, 2, _pb(1, 2, b"value") + _pb(2, 2, _pb(8, 2, tcont)))
    sub = _pb(1, 2, b"sub") + _pb(2, 2, b"Sub") + _pb(3, 2, b"X") + _pb(3, 2, b"Const") + _pb(3, 2, b"^ctl") + \
        _pb(5, 2, _pb(1, 2, b"N") + _pb(2, 2, _pb(3, 0, 4)))
    graph = b"".join(_pb(1, 2, n) for n in (ph, const, stack, sub))
    meta = _pb(1, 2, b"ignored meta_info_def") + _pb(2, 2, graph)
    path = tmp_path / "g.meta"
    path.write_bytes(meta)
    nodes, order = tf1_files.read_meta_graph(str(path))
    assert order == ["X", "Const", "s/stack", "sub"]
    assert nodes["X"]["op"] == "Placeholder" and nodes["X"]["attrs"]["shape"] == ("shape", [-1, 138])
    assert np.allclose(nodes["Const"]["attrs"]["value"], [1.0, 5.0, 3.14]) and nodes["Const"]["attrs"]["value"].dtype == np.float32
    assert list(nodes["s/stack"]["attrs"]["value"]) == [0, 5]
    assert nodes["sub"]["inputs"] == ["X", "Const", "^ctl"] and nodes["sub"]["attrs"]["N"] == 4
