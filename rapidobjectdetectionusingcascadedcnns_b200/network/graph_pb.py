"""Reader (and test writer) for the reference's frozen stage graphs ``graph_<session>_<i>.pb``.

The reference exports each stage with ``convert_variables_to_constants`` + ``optimize_for_inference``
(`app/train_app.py:177-227`, `network/net.py:561-566`), which turns every variable into a ``Const`` node
that keeps the variable's name.  The names come from the graph builders:

    conv_layer_<j>/W, conv_layer_<j>/b          `network/net.py:192`, `network/net_builder.py:33-52`
    bottleneck_out/W|b   (stage 0 / single)      `network/net.py:229-231`, `network/net_builder.py:78-92`
    bottleneck_out_part1/W|b   (stages >= 1)
    fc2/W, fc2/b                                 `network/net.py:156-158`

No TensorFlow is needed: ``tensorboard.compat.proto`` ships the GraphDef / TensorProto messages.
Verified here only against files written by ``write_stage_pb`` (no TF-exported file is available
offline); a real export may carry an import prefix, which ``read_stage_pb`` strips by suffix match.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def _protos():
    from tensorboard.compat.proto import graph_pb2, tensor_pb2, tensor_shape_pb2, types_pb2
    return graph_pb2, tensor_pb2, tensor_shape_pb2, types_pb2


def _tensor_to_numpy(t) -> np.ndarray:
    shape = [d.size for d in t.tensor_shape.dim]
    if t.tensor_content:
        return np.frombuffer(t.tensor_content, dtype=np.float32).reshape(shape).copy()
    vals = np.asarray(t.float_val, np.float32)
    n = int(np.prod(shape)) if shape else 1
    if vals.size == 1 and n > 1:
        vals = np.full(n, vals[0], np.float32)
    return vals.reshape(shape)


def read_stage_pb(path: str) -> Dict[str, np.ndarray]:
    graph_pb2, _, _, types_pb2 = _protos()
    gd = graph_pb2.GraphDef()
    with open(path, "rb") as f:
        gd.ParseFromString(f.read())
    consts = {}
    for node in gd.node:
        if node.op == "Const" and "value" in node.attr and node.attr["value"].tensor.dtype == types_pb2.DT_FLOAT:
            consts[node.name] = _tensor_to_numpy(node.attr["value"].tensor)

    def find(*suffixes):
        for name, arr in consts.items():
            for s in suffixes:
                if name == s or name.endswith("/" + s):
                    return arr
        raise KeyError("no Const node named %s in %s" % (" / ".join(suffixes), path))

    out = {}
    j = 0
    while True:
        try:
            out["conv%d_W" % j] = find("conv_layer_%d/W" % j)
            out["conv%d_b" % j] = find("conv_layer_%d/b" % j)
        except KeyError:
            break
        j += 1
    if j == 0:
        raise KeyError("no conv_layer_0/W in %s (Inception-based graphs are out of scope)" % path)
    out["fc1_W"] = find("bottleneck_out/W", "bottleneck_out_part1/W")
    out["fc1_b"] = find("bottleneck_out/b", "bottleneck_out_part1/b")
    out["fc2_W"] = find("fc2/W")
    out["fc2_b"] = find("fc2/b")
    return out


def write_stage_pb(path: str, weights: Dict[str, np.ndarray], stage_has_bottleneck_in: bool, in_size: int):
    """Write a minimal GraphDef with the reference's node names (placeholder ``X`` with its static
    shape + the weight constants).  Test helper / converter; it does not reproduce the op nodes."""
    graph_pb2, tensor_pb2, tensor_shape_pb2, types_pb2 = _protos()
    gd = graph_pb2.GraphDef()
    x = gd.node.add()
    x.name, x.op = "X", "Placeholder"
    x.attr["dtype"].type = types_pb2.DT_FLOAT
    for d in (-1, in_size, in_size, 3):
        x.attr["shape"].shape.dim.add().size = d
    fc1 = "bottleneck_out_part1" if stage_has_bottleneck_in else "bottleneck_out"
    names = {"fc1_W": fc1 + "/W", "fc1_b": fc1 + "/b", "fc2_W": "fc2/W", "fc2_b": "fc2/b"}
    j = 0
    while "conv%d_W" % j in weights:
        names["conv%d_W" % j] = "conv_layer_%d/W" % j
        names["conv%d_b" % j] = "conv_layer_%d/b" % j
        j += 1
    for key, name in names.items():
        arr = np.ascontiguousarray(weights[key], np.float32)
        n = gd.node.add()
        n.name, n.op = name, "Const"
        n.attr["dtype"].type = types_pb2.DT_FLOAT
        t = n.attr["value"].tensor
        t.dtype = types_pb2.DT_FLOAT
        for d in arr.shape:
            t.tensor_shape.dim.add().size = d
        t.tensor_content = arr.tobytes()
    with open(path, "wb") as f:
        f.write(gd.SerializeToString())
