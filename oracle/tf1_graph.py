"""tf1_graph.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy interpreter for the forward cone of a TF1 GraphDef: it EXECUTES the reference's own GA3C-CADRL graph
(``policies/GA3C_CADRL/checkpoints/*/network_*.meta``, the graph ``network.py:47-74`` restores and runs) node by node --
Placeholder / Const / VariableV2 (from the checkpoint) / Sub / RealDiv / StridedSlice / Reshape / Transpose / Cast /
TensorArray scatter + read / the dynamic_rnn while frame (Enter, Merge, Switch, LoopCond, NextIteration, Exit) / ConcatV2 /
MatMul / BiasAdd / Split / Sigmoid / Tanh / Mul / Add / Select / GreaterEqual / Relu / Softmax ... -- with each op's
documented semantics in float32.  TensorFlow itself is not installed in this image, so this interpreter is the pin for the
C restatement (oracle/cav_oracle.c: cavo_ga3c_forward) and the CUDA kernel: it takes nothing from either, only the graph
file and the checkpoint.  What it cannot pin is TensorFlow's own kernels' rounding (Eigen's matmul summation order, its
sigmoid / tanh approximations): parity is therefore stated at 1e-5 on the softmax output, not bit-exact.

The file parsers (protobuf wire format, tensor-bundle SSTable, CRC-32C checked) are
gym_collision_avoidance_b200/envs/policies/GA3C_CADRL/tf1_files.py.
"""
import numpy as np


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _sigmoid(x):
    x = _f32(x)
    return (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32)


class GraphInterpreter(object):
    def __init__(self, nodes, variables):
        self.nodes = nodes
        self.vars = variables
        self.iterations = 0

    def run(self, fetch, feeds):
        self.memo, self.feeds, self.tarrays = {}, feeds, {}
        need, stack = set(), [fetch]
        while stack:
            n = stack.pop().lstrip("^").split(":")[0]
            if n not in need:
                need.add(n)
                stack += self.nodes[n]["inputs"]
        self.need = need
        return self.ev(fetch, None)

    def _in_frame(self, name):
        return name.startswith("rnn/while/")

    def ev(self, ref, frame):
        name, _, idx = ref.lstrip("^").partition(":")
        idx = int(idx) if idx else 0
        if frame is not None and name in frame:
            v = frame[name]
        elif name in self.memo:
            v = self.memo[name]
        else:
            v = self.node(name, frame)
            if frame is not None and self._in_frame(name):
                frame[name] = v  # values inside the while frame live for one iteration
            else:
                self.memo[name] = v
        return v[idx] if isinstance(v, tuple) else v

    def node(self, name, frame):
        nd = self.nodes[name]
        op, A = nd["op"], nd["attrs"]
        ins = [i for i in nd["inputs"] if not i.startswith("^")]

        def e(k):
            return self.ev(ins[k], frame)
        if op == "Placeholder":
            return self.feeds[name]
        if op == "Const":
            return A["value"]
        if op == "VariableV2":
            return self.vars[name + ":0"] if name + ":0" in self.vars else self.vars[name]
        if op in ("Identity", "Enter", "LoopCond", "NextIteration"):
            return e(0)
        if op == "Sub":
            return e(0) - e(1)
        if op == "RealDiv":
            return e(0) / e(1)
        if op == "Add":
            return e(0) + e(1)
        if op == "Mul":
            return e(0) * e(1)
        if op == "Cast":
            return e(0).astype({3: np.int32, 1: np.float32}[A["DstT"][1]])  # float -> int32 truncates toward zero
        if op == "StridedSlice":
            x, b, en, st = e(0), e(1), e(2), e(3)
            sl = []
            for d in range(len(b)):
                if (A["shrink_axis_mask"] >> d) & 1:
                    sl.append(int(b[d]))
                else:
                    sl.append(slice(None if (A["begin_mask"] >> d) & 1 else int(b[d]),
                                    None if (A["end_mask"] >> d) & 1 else int(en[d]), int(st[d])))
            return np.asarray(x)[tuple(sl)]
        if op == "Reshape":
            return np.reshape(e(0), [int(v) for v in e(1)])
        if op == "Range":
            return np.arange(int(e(0)), int(e(1)), int(e(2)), dtype=np.int32)
        if op == "ConcatV2":
            n = A["N"]
            return np.concatenate([np.atleast_1d(self.ev(ins[k], frame)) for k in range(n)], axis=int(self.ev(ins[n], frame)))
        if op == "Transpose":
            return np.transpose(e(0), [int(v) for v in e(1)])
        if op == "Shape":
            return np.array(np.shape(e(0)), dtype=np.int32)
        if op == "ExpandDims":
            return np.expand_dims(e(0), int(e(1)))
        if op == "Fill":
            return np.full([int(v) for v in e(0)], e(1), dtype=np.asarray(e(1)).dtype)
        if op == "Pack":
            return np.stack([self.ev(i, frame) for i in ins], axis=A.get("axis", 0))
        if op == "Equal":
            return e(0) == e(1)
        if op == "All":
            return np.all(e(0))
        if op == "Assert":
            assert bool(e(0)), "graph Assert %s failed" % name
            return None
        if op == "Less":
            return e(0) < e(1)
        if op == "GreaterEqual":
            return e(0) >= e(1)
        if op == "LogicalAnd":
            return np.logical_and(e(0), e(1))
        if op == "Minimum":
            return np.minimum(e(0), e(1))
        if op == "Maximum":
            return np.maximum(e(0), e(1))
        if op == "Max":
            return np.max(e(0), axis=tuple(int(v) for v in np.atleast_1d(e(1))))
        if op == "Min":
            return np.min(e(0), axis=tuple(int(v) for v in np.atleast_1d(e(1))))
        if op == "Select":
            c, a, b = e(0), e(1), e(2)
            if np.ndim(c) == 1 and np.ndim(a) == 2:  # a vector condition selects rows
                c = c[:, None]
            return np.where(c, a, b)
        if op == "MatMul":
            assert not A["transpose_a"] and not A["transpose_b"]
            return (_f32(e(0)) @ _f32(e(1))).astype(np.float32)
        if op == "BiasAdd":
            return (e(0) + e(1)).astype(np.float32)
        if op == "Split":
            return tuple(np.split(e(1), A["num_split"], axis=int(e(0))))
        if op == "Sigmoid":
            return _sigmoid(e(0))
        if op == "Tanh":
            return np.tanh(_f32(e(0))).astype(np.float32)
        if op == "Relu":
            return np.maximum(e(0), np.float32(0))
        if op == "Softmax":
            x = _f32(e(0))
            ex = np.exp(x - x.max(axis=-1, keepdims=True))
            return (ex / ex.sum(axis=-1, keepdims=True)).astype(np.float32)
        if op == "Slice":
            x, b, s = e(0), e(1), e(2)
            return x[tuple(slice(int(bi), None if int(si) < 0 else int(bi) + int(si)) for bi, si in zip(b, s))]
        if op == "Prod":
            return np.prod(e(0), axis=tuple(int(v) for v in np.atleast_1d(e(1)))).astype(np.int32)
        if op == "Squeeze":
            return np.squeeze(e(0), axis=-1)  # the value head: squeeze_dims = [1] of a [rows, 1] tensor
        if op == "TensorArrayV3":
            return (name, np.float32(0))  # (handle, flow)
        if op == "TensorArrayScatterV3":
            h, idx, val = e(0), e(1), e(2)
            self.tarrays[h] = {int(i): val[k] for k, i in enumerate(idx)}
            return np.float32(0)
        if op == "TensorArrayReadV3":
            h = e(0)
            e(2)  # the flow edge orders the read after the scatter
            return self.tarrays[h][int(e(1))]
        if op == "Exit":
            return self._while(name)
        raise NotImplementedError("%s (%s)" % (op, name))

    def _while(self, exit_name):
        """The dynamic_rnn while frame: Merge_k <- (Enter_k, NextIteration_k), Switch_k(Merge_k, LoopCond), Exit_k."""
        if "__while__" not in self.memo:
            merges = [n for n, nd in self.nodes.items() if nd["op"] == "Merge" and self._in_frame(n) and n in self.need]
            state = {m: self.ev(self.nodes[m]["inputs"][0], None) for m in merges}  # the Enter side
            cond = [n for n, nd in self.nodes.items() if nd["op"] == "LoopCond" and n in self.need][0]
            switches = {n: nd for n, nd in self.nodes.items()
                        if nd["op"] == "Switch" and self._in_frame(n) and nd["inputs"][0] in merges}
            self.iterations = 0
            while True:
                frame = dict(state)
                if not bool(self.ev(cond, frame)):
                    break
                for s, nd in switches.items():
                    frame[s] = (None, state[nd["inputs"][0]])  # output 1 = the "true" branch
                state = {m: self.ev(self.nodes[m]["inputs"][1], frame) for m in merges}
                self.iterations += 1
            self.memo["__while__"] = {n: state[switches[nd["inputs"][0]]["inputs"][0]] for n, nd in self.nodes.items()
                                      if nd["op"] == "Exit" and n in self.need and nd["inputs"][0] in switches}
        return self.memo["__while__"][exit_name]


def load(prefix):
    """GraphInterpreter over ``<prefix>.meta`` + the checkpoint ``<prefix>.index`` / ``.data-*``."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gym_collision_avoidance_b200.envs.policies.GA3C_CADRL import tf1_files
    nodes, _ = tf1_files.read_meta_graph(prefix + ".meta")
    return GraphInterpreter(nodes, tf1_files.read_checkpoint(prefix))
