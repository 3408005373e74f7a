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
    """The meta-graph must be the network cavoid.h documents; anything else is rejected instead of mis-evaluated."""
    def need(cond, what):
        if not cond:
            raise abi.CavoidError("GA3C-CADRL meta-graph: unexpected %s" % what)
    need(nodes.get("X", {}).get("op") == "Placeholder" and nodes["X"]["attrs"]["shape"] == ("shape", [-1, X_LEN]), "X shape")
    need("Softmax" in nodes and nodes["Softmax"]["inputs"] == ["logits_p/BiasAdd"], "output head")
    need(nodes["sub"]["inputs"] == ["X", "Const"] and nodes["div"]["inputs"] == ["sub", "Const_1"], "input normalisation")
    need(list(nodes["Reshape/shape"]["attrs"]["value"]) == [-1, MAX_OTHERS, 7], "sequence reshape")
    need(list(nodes["strided_slice_1/stack"]["attrs"]["value"]) == [0, 1] and
         list(nodes["strided_slice_1/stack_1"]["attrs"]["value"]) == [0, 5], "host-state slice")
    need(list(nodes["strided_slice_2/stack"]["attrs"]["value"]) == [0, 5], "other-agent slice")
    need(nodes["ToInt32"]["inputs"] == ["strided_slice"] and nodes["sequence_length"]["inputs"] == ["ToInt32"], "sequence length")
    cell = "rnn/while/rnn/lstm_cell/lstm_cell/"
    need(float(nodes[cell + "add/y"]["attrs"]["value"]) == 1.0, "LSTM forget bias")
    need(nodes[cell + "split"]["attrs"]["num_split"] == 4, "LSTM gate split")
    need(nodes[cell + "Sigmoid_1"]["inputs"] == [cell + "split"] and nodes[cell + "Tanh"]["inputs"] == [cell + "split:1"] and
         nodes[cell + "add"]["inputs"][0] == cell + "split:2" and nodes[cell + "Sigmoid_2"]["inputs"] == [cell + "split:3"],
         "LSTM gate order (i, j, f, o)")
    need(nodes["layer1_input"]["inputs"][:2] == ["strided_slice_1", "rnn/while/Exit_3"], "layer1 input")
    for a, b in (("layer1", "layer1_input"), ("layer2", "layer1/Relu"), ("fullyconnected1", "Flatten/Reshape"),
                 ("logits_p", "fullyconnected1/Relu")):
        need(nodes[a + "/MatMul"]["inputs"][0] == b, a + " input")
    need(nodes["Flatten/Reshape"]["inputs"][0] == "layer2/Relu", "flatten input")


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
