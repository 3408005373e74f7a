"""GA3C-CADRL policy network: the reference's TF1 checkpoint, loaded without TensorFlow and evaluated on the GPU.

Reference ``envs/policies/GA3C_CADRL/network.py``: ``Actions`` (:7-16, the 11 discrete actions) and ``NetworkVP_rnn``
(:18-77: ``simple_load`` restores ``<filename>.meta`` + checkpoint into a TF1 session, ``predict_p`` runs
``Softmax:0`` on a ``[rows, 138]`` batch after ``crop_x``).  Here ``simple_load`` reads the same files
(``tf1_files.py``), checks that the meta-graph is the architecture ``include/cavoid.h`` documents, packs the variables
into the float32 blob the CUDA kernels take (``CAV_GA3C_BLOB_FLOATS``), and ``predict_p`` calls ``cav_ga3c_forward``.
"""
import ctypes

import numpy as np

from .... import abi
from . import tf1_files

X_LEN, MAX_OTHERS, HIDDEN, FC, NUM_ACTIONS = 138, 19, 64, 256, 11
BLOB_FLOATS = 170784


class Actions(object):
    """11 actions ``[speed fraction of pref_speed, heading change]`` (reference network.py:7-16):
    full speed x {-pi/6, -pi/12, 0, pi/12, pi/6}; half speed and zero speed x {-pi/6, 0, pi/6}."""

    def __init__(self):
        def grid(speed, step, count):  # what np.mgrid[speed:..., -pi/6:pi/6+0.01:step] enumerates: start + k * step
            return [[speed, -np.pi / 6 + k * step] for k in range(count)]
        self.actions = np.array(grid(1.0, np.pi / 12, 5) + grid(0.5, np.pi / 6, 3) + grid(0.0, np.pi / 6, 3))
        self.num_actions = len(self.actions)


def _check_graph(nodes):
    """The meta-graph must be the network cavoid.h documents; anything else is rejected instead of mis-evaluated.  Node
    names differ between TF versions (the IROS18 graph says ``rnn/while/rnn/lstm_cell/lstm_cell/...``, the 2019 ones
    ``rnn/while/lstm_cell/...``), so nodes are found by op and wiring."""
    def need(cond, what):
        if not cond:
            raise abi.CavoidError("GA3C-CADRL meta-graph: unexpected %s" % what)

    def ins(name):
        return [i for i in nodes[name]["inputs"] if not i.startswith("^")]

    def const(name):
        return nodes[name]["attrs"]["value"]
    need(nodes.get("X", {}).get("op") == "Placeholder" and nodes["X"]["attrs"]["shape"] == ("shape", [-1, X_LEN]), "X shape")
    need("Softmax" in nodes and ins("Softmax") == ["logits_p/BiasAdd"], "output head")
    need(ins("sub") == ["X", "Const"] and nodes["sub"]["op"] == "Sub" and ins("div") == ["sub", "Const_1"] and
         nodes["div"]["op"] == "RealDiv", "input normalisation")
    need(list(const("Reshape/shape")) == [-1, MAX_OTHERS, 7] and ins("Reshape")[0] == "strided_slice_2", "sequence reshape")
    need(ins("strided_slice")[0] == "X" and list(const("strided_slice/stack")) == [0, 0] and
         list(const("strided_slice/stack_1")) == [0, 1], "sequence-length slice")
    need(ins("strided_slice_1")[0] == "div" and list(const("strided_slice_1/stack")) == [0, 1] and
         list(const("strided_slice_1/stack_1")) == [0, 5], "host-state slice")
    need(ins("strided_slice_2")[0] == "div" and list(const("strided_slice_2/stack")) == [0, 5], "other-agent slice")
    casts = [n for n, nd in nodes.items() if nd["op"] == "Cast" and ins(n) == ["strided_slice"]]
    need(len(casts) == 1 and nodes[casts[0]]["attrs"]["DstT"] == ("type", 3), "sequence length cast")
    splits = [n for n, nd in nodes.items() if nd["op"] == "Split" and "gradients" not in n and "lstm_cell" in n]
    need(len(splits) == 1 and nodes[splits[0]]["attrs"]["num_split"] == 4, "LSTM gate split")
    cell = splits[0][:-len("split")]
    need(float(const(cell + "add/y")) == 1.0, "LSTM forget bias")
    need(ins(cell + "Sigmoid_1") == [cell + "split"] and ins(cell + "Tanh") == [cell + "split:1"] and
         ins(cell + "add")[0] == cell + "split:2" and ins(cell + "Sigmoid_2") == [cell + "split:3"],
         "LSTM gate order (i, j, f, o)")
    need(ins(cell + "add_1") == [cell + "mul", cell + "mul_1"] and ins(cell + "mul")[0] == cell + "Sigmoid" and
         ins(cell + "mul_1") == [cell + "Sigmoid_1", cell + "Tanh"] and
         ins(cell + "mul_2") == [cell + "Sigmoid_2", cell + "Tanh_1"] and ins(cell + "Tanh_1") == [cell + "add_1"], "LSTM cell update")
    l1 = ins("layer1_input")
    need(nodes["layer1_input"]["op"] == "ConcatV2" and l1[0] == "strided_slice_1" and nodes[l1[1]]["op"] == "Exit", "layer1 input")
    # the Exit must carry h (the state concatenated with the input inside the cell), not c
    sw = ins(l1[1])[0]
    merge = ins(sw)[0]
    ident = [n for n, nd in nodes.items() if nd["op"] == "Identity" and ins(n) == [sw + ":1"]]
    concat = [n for n, nd in nodes.items() if nd["op"] == "ConcatV2" and n.startswith(cell) and len(ins(n)) == 3]
    need(len(ident) == 1 and any(ident[0] == ins(c)[1] for c in concat) and nodes[merge]["op"] == "Merge", "recurrent state wiring")
    need(ins("layer1/MatMul")[0] == "layer1_input" and ins("layer2/MatMul")[0] == "layer1/Relu" and
         ins("logits_p/MatMul")[0] == "fullyconnected1/Relu", "dense stack")
    flat = ins("fullyconnected1/MatMul")[0]
    need(nodes[flat]["op"] == "Reshape" and ins(flat)[0] == "layer2/Relu", "flatten input")
    for name in ("layer1", "layer2", "fullyconnected1"):
        need(nodes[name + "/Relu"]["op"] == "Relu" and ins(name + "/Relu") == [name + "/BiasAdd"] and
             ins(name + "/BiasAdd")[0] == name + "/MatMul", name + " activation")


def pack_blob(variables, avg, std):
    """float32 [BLOB_FLOATS] in the section order of include/cavoid.h from the checkpoint's variables."""
    def var(name, shape):
        v = variables.get(name + ":0", variables.get(name))
        if v is None or tuple(v.shape) != shape or v.dtype != np.float32:
            raise abi.CavoidError("GA3C-CADRL checkpoint: variable %s missing or not float32 %s" % (name, shape))
        return v
    wl = var("rnn/lstm_cell/kernel", (7 + HIDDEN, 4 * HIDDEN)).reshape(7 + HIDDEN, 4, HIDDEN).transpose(0, 2, 1)  # [k][unit][gate]
    bl = var("rnn/lstm_cell/bias", (4 * HIDDEN,)).reshape(4, HIDDEN).T
    parts = [np.asarray(avg, np.float32).reshape(X_LEN), np.asarray(std, np.float32).reshape(X_LEN), wl, bl,
             var("layer1/kernel", (4 + HIDDEN, FC)), var("layer1/bias", (FC,)),
             var("layer2/kernel", (FC, FC)), var("layer2/bias", (FC,)),
             var("fullyconnected1/kernel", (FC, FC)), var("fullyconnected1/bias", (FC,)),
             var("logits_p/kernel", (FC, NUM_ACTIONS)), var("logits_p/bias", (NUM_ACTIONS,)), np.zeros(1, np.float32)]
    blob = np.concatenate([np.ascontiguousarray(p, dtype=np.float32).ravel() for p in parts])
    assert blob.size == BLOB_FLOATS
    return blob


def load_blob(filename):
    """The reference's checkpoint ``filename`` (+ ``.meta`` / ``.index`` / ``.data-00000-of-00001``) as the weight blob."""
    nodes, _ = tf1_files.read_meta_graph(filename + ".meta")
    _check_graph(nodes)
    return pack_blob(tf1_files.read_checkpoint(filename), nodes["Const"]["attrs"]["value"], nodes["Const_1"]["attrs"]["value"])


class NetworkVPCore(object):
    def __init__(self, device, model_name, num_actions):
        self.device = device
        self.model_name = model_name
        self.num_actions = num_actions
        self.blob = None     # numpy float32 [BLOB_FLOATS]
        self.d_blob = None   # the same on the CUDA device

    def crop_x(self, x):
        """Reference network.py:24-35: cut or zero-pad the flattened observation to the network's input width."""
        x = np.asarray(x)
        if x.shape[-1] > X_LEN:
            return x[:, :X_LEN]
        if x.shape[-1] < X_LEN:
            x_ = np.zeros((x.shape[0], X_LEN))
            x_[:, :x.shape[1]] = x
            return x_
        return x

    def _torch_device(self):
        import torch
        d = self.device
        if isinstance(d, str) and d.startswith("/"):  # the reference passes TF device strings ('/cpu:0', '/gpu:0')
            d = "cuda:0"
        return torch.device(d)

    def simple_load(self, filename=None):
        if filename is None:
            raise NotImplementedError("[network.py] Didn't define simple_load filename")
        self.set_blob(load_blob(filename))

    def set_blob(self, blob):
        import torch
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        if blob.size != BLOB_FLOATS:
            raise abi.CavoidError("GA3C-CADRL weight blob must hold %d floats" % BLOB_FLOATS)
        self.blob = blob
        self.d_blob = torch.from_numpy(blob).to(self._torch_device())

    def predict_p(self, x):
        """Softmax over the 11 actions for a ``[rows, <=138+]`` batch (reference network.py:37-41), on the GPU."""
        import torch
        if self.d_blob is None:
            raise abi.CavoidError("GA3C-CADRL network: simple_load / set_blob first")
        x = np.ascontiguousarray(self.crop_x(x), dtype=np.float32)
        L = abi.load_library()
        dev = self.d_blob.device
        with torch.cuda.device(dev):
            dx = torch.from_numpy(x).to(dev)
            probs = torch.empty((x.shape[0], NUM_ACTIONS), dtype=torch.float32, device=dev)
            rc = L.cav_ga3c_forward(self.d_blob.data_ptr(), dx.data_ptr(), x.shape[0], probs.data_ptr(), None,
                                    ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if rc != 0:
                raise abi.CavoidError("cav_ga3c_forward failed (%d): %s" % (rc, L.cav_last_error(None).decode()))
            return probs.cpu().numpy()


class NetworkVP_rnn(NetworkVPCore):
    def __init__(self, device, model_name, num_actions):
        super(NetworkVP_rnn, self).__init__(device, model_name, num_actions)
