"""Reader (and writer) for the reference's frozen stage graphs ``graph_<session>_<i>.pb``.

The reference exports each stage with ``convert_variables_to_constants`` + ``optimize_for_inference`` and then
re-imports the result with ``name=""`` behind a new ``X`` placeholder that carries the static input shape
(`app/train_app.py:177-227`, `network/net.py:561-566`).  Every variable becomes a ``Const`` node that keeps the
variable's name; the op nodes keep the names the graph builders gave them:

    X                         Placeholder, shape [-1, H, W, 3]                 `app/train_app.py:203-205`
    conv_layer_<j>/W, /b      Const          `network/net.py:192`, `network/net_builder.py:33-52`
    conv_layer_<j>/conv       Conv2D  (strides [1,s,s,1], SAME)                `network/net_builder.py:40-45`
    conv_layer_<j>/h          BiasAdd;  Relu[_<j>]                             `network/net.py:199-201`
    pool_layer_<j>/MaxPool    MaxPool (ksize [1,k,k,1], strides [1,s,s,1], SAME)  `network/net_builder.py:6-17`
    flatten/Reshape           Reshape                                          `network/net_builder.py:107-123`
    bottleneck_out/W|b|MatMul|h|Relu        (stage 0 / single)                 `network/net.py:229-236`
    bottleneck_out_part1/...  + bottleneck_in Placeholder + bottleneck_out/Relu = ConcatV2  (stages >= 1) `:125-146`
    fc2/W|b|MatMul|h, final_softmax                                            `network/net.py:156-165`

``read_stage_pb`` takes the weights from the ``Const`` nodes (matched by name SUFFIX, so an ``import/`` or
``graph_<i>/`` prefix does not matter; ``<name>/read`` Identity nodes that a non-optimised freeze keeps are
ignored) and the GEOMETRY from the graph itself: input size from the ``X`` placeholder's shape, conv stride from
the ``Conv2D`` strides, pool size / stride from the ``MaxPool`` attributes.  A file whose geometry is missing or
contradicts its weights raises instead of running with assumed defaults.

No TensorFlow is needed: ``tensorboard.compat.proto`` ships the GraphDef / TensorProto messages.  No
TF-exported file is available offline; ``write_stage_pb`` emits the node set listed above (optionally with a
name prefix and the Identity ``/read`` nodes of an un-optimised freeze) and the tests read that back.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

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


def _strip(name: str, suffix: str) -> bool:
    return name == suffix or name.endswith("/" + suffix)


def read_stage_graph(path: str) -> Tuple[Dict[str, np.ndarray], Dict[str, int]]:
    """Weights (``conv<j>_W|b``, ``fc1_W|b``, ``fc2_W|b``) and geometry (``in_size``, ``conv_stride``,
    ``pool_size``, ``pool_stride``) of one frozen stage graph."""
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
                if _strip(name, s):
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

    # ---- geometry from the graph itself
    geo: Dict[str, int] = {}
    for node in gd.node:
        if node.op == "Placeholder" and _strip(node.name, "X") and "shape" in node.attr:
            dims = [d.size for d in node.attr["shape"].shape.dim]
            if len(dims) != 4 or dims[3] != 3:
                raise ValueError("%s: placeholder X has shape %r, expected [None, H, W, 3]" % (path, dims))
            if dims[1] != dims[2]:
                raise ValueError("%s: non-square input %dx%d is not supported" % (path, dims[1], dims[2]))
            geo["in_size"] = int(dims[1])
    conv_strides = {tuple(n.attr["strides"].list.i) for n in gd.node if n.op == "Conv2D"}
    pool_attrs = {(tuple(n.attr["ksize"].list.i), tuple(n.attr["strides"].list.i)) for n in gd.node if n.op == "MaxPool"}
    for n in gd.node:
        if n.op in ("Conv2D", "MaxPool") and n.attr["padding"].s not in (b"SAME", b""):
            raise ValueError("%s: node %s uses padding %r; the reference builds SAME only" % (path, n.name, n.attr["padding"].s))
    if len(conv_strides) > 1 or len(pool_attrs) > 1:
        raise ValueError("%s: conv / pool layers with different strides or sizes are not what network/net.py builds" % path)
    if conv_strides:
        st = conv_strides.pop()
        if len(st) != 4 or st[0] != 1 or st[3] != 1 or st[1] != st[2]:
            raise ValueError("%s: unsupported Conv2D strides %r" % (path, st))
        geo["conv_stride"] = int(st[1])
    if pool_attrs:
        ks, st = pool_attrs.pop()
        if len(ks) != 4 or len(st) != 4 or ks[1] != ks[2] or st[1] != st[2] or ks[0] != 1 or ks[3] != 1:
            raise ValueError("%s: unsupported MaxPool ksize %r / strides %r" % (path, ks, st))
        geo["pool_size"], geo["pool_stride"] = int(ks[1]), int(st[1])
    missing = [k for k in ("in_size", "conv_stride", "pool_size", "pool_stride") if k not in geo]
    if missing:
        raise ValueError("%s carries no %s (X placeholder shape / Conv2D / MaxPool nodes): refusing to guess the "
                         "geometry of a frozen graph" % (path, ", ".join(missing)))
    return out, geo


def read_stage_pb(path: str) -> Dict[str, np.ndarray]:
    """Weights only (kept for callers that bring their own geometry)."""
    return read_stage_graph(path)[0]


def write_stage_pb(path: str, weights: Dict[str, np.ndarray], stage_has_bottleneck_in: bool, in_size: int,
                   conv_stride: int = 1, pool_size: int = 3, pool_stride: int = 1, prefix: str = "",
                   identity_reads: bool = False, ops: bool = True):
    """Write a stage graph with the reference's node names: the ``X`` placeholder with its static shape, the weight
    constants and (``ops=True``) the op nodes of `network/net.py:101-240` with the attributes the reader takes
    the geometry from.  ``prefix`` prepends a scope (``import/``); ``identity_reads=True`` adds the
    ``<var>/read`` Identity nodes an un-optimised ``convert_variables_to_constants`` output keeps and routes
    the ops through them."""
    graph_pb2, tensor_pb2, tensor_shape_pb2, types_pb2 = _protos()
    gd = graph_pb2.GraphDef()
    P = prefix

    def node(name, op, inputs=()):
        n = gd.node.add()
        n.name, n.op = P + name, op
        for i in inputs:
            n.input.append(P + i)
        return n

    def const_f32(name, arr):
        arr = np.ascontiguousarray(arr, np.float32)
        n = node(name, "Const")
        n.attr["dtype"].type = types_pb2.DT_FLOAT
        t = n.attr["value"].tensor
        t.dtype = types_pb2.DT_FLOAT
        for d in arr.shape:
            t.tensor_shape.dim.add().size = d
        t.tensor_content = arr.tobytes()
        if identity_reads:
            r = node(name + "/read", "Identity", [name])
            r.attr["T"].type = types_pb2.DT_FLOAT
            return name + "/read"
        return name

    def const_i32(name, vals):
        arr = np.asarray(vals, np.int32)
        n = node(name, "Const")
        n.attr["dtype"].type = types_pb2.DT_INT32
        t = n.attr["value"].tensor
        t.dtype = types_pb2.DT_INT32
        for d in arr.shape:
            t.tensor_shape.dim.add().size = d
        t.tensor_content = arr.tobytes()
        return name

    def placeholder(name, dims):
        x = node(name, "Placeholder")
        x.attr["dtype"].type = types_pb2.DT_FLOAT
        for d in dims:
            x.attr["shape"].shape.dim.add().size = d

    placeholder("X", (-1, in_size, in_size, 3))
    cur = "X"
    j = 0
    side, chan = in_size, 3
    while "conv%d_W" % j in weights:
        scope = "conv_layer_%d" % j
        w = const_f32(scope + "/W", weights["conv%d_W" % j])
        b = const_f32(scope + "/b", weights["conv%d_b" % j])
        chan = int(np.asarray(weights["conv%d_W" % j]).shape[3])
        if ops:
            c = node(scope + "/conv", "Conv2D", [cur, w])
            c.attr["T"].type = types_pb2.DT_FLOAT
            c.attr["strides"].list.i.extend([1, conv_stride, conv_stride, 1])
            c.attr["padding"].s = b"SAME"
            c.attr["data_format"].s = b"NHWC"
            h = node(scope + "/h", "BiasAdd", [scope + "/conv", b])
            h.attr["T"].type = types_pb2.DT_FLOAT
            relu = "Relu" if j == 0 else "Relu_%d" % j
            node(relu, "Relu", [scope + "/h"]).attr["T"].type = types_pb2.DT_FLOAT
            mp = node("pool_layer_%d/MaxPool" % j, "MaxPool", [relu])
            mp.attr["T"].type = types_pb2.DT_FLOAT
            mp.attr["ksize"].list.i.extend([1, pool_size, pool_size, 1])
            mp.attr["strides"].list.i.extend([1, pool_stride, pool_stride, 1])
            mp.attr["padding"].s = b"SAME"
            cur = "pool_layer_%d/MaxPool" % j
        side = -(-side // conv_stride)
        side = -(-side // pool_stride)
        j += 1
    fc1 = "bottleneck_out_part1" if stage_has_bottleneck_in else "bottleneck_out"
    w1 = const_f32(fc1 + "/W", weights["fc1_W"])
    b1 = const_f32(fc1 + "/b", weights["fc1_b"])
    w2 = const_f32("fc2/W", weights["fc2_W"])
    b2 = const_f32("fc2/b", weights["fc2_b"])
    if ops:
        const_i32("flatten/Reshape/shape", [-1, side * side * chan])
        node("flatten/Reshape", "Reshape", [cur, "flatten/Reshape/shape"]).attr["T"].type = types_pb2.DT_FLOAT
        node(fc1 + "/MatMul", "MatMul", ["flatten/Reshape", w1]).attr["T"].type = types_pb2.DT_FLOAT
        node(fc1 + "/h", "BiasAdd", [fc1 + "/MatMul", b1]).attr["T"].type = types_pb2.DT_FLOAT
        node(fc1 + "/Relu", "Relu", [fc1 + "/h"]).attr["T"].type = types_pb2.DT_FLOAT
        feat = fc1 + "/Relu"
        if stage_has_bottleneck_in:
            n_prev = int(np.asarray(weights["fc2_W"]).shape[0]) - int(np.asarray(weights["fc1_W"]).shape[1])
            placeholder("bottleneck_in", (-1, n_prev))
            const_i32("bottleneck_out/Relu/axis", 1)
            cc = node("bottleneck_out/Relu", "ConcatV2", [feat, "bottleneck_in", "bottleneck_out/Relu/axis"])
            cc.attr["T"].type = types_pb2.DT_FLOAT
            cc.attr["N"].i = 2
            feat = "bottleneck_out/Relu"
        # dropout with keep probability 1.0 survives the export as an identity (`network/net.py:150-153`)
        dp = node("dropout_placeholder", "PlaceholderWithDefault", [])
        dp.attr["dtype"].type = types_pb2.DT_FLOAT
        node("dropout/mul", "Mul", [feat, "dropout_placeholder"]).attr["T"].type = types_pb2.DT_FLOAT
        node("fc2/MatMul", "MatMul", ["dropout/mul", w2]).attr["T"].type = types_pb2.DT_FLOAT
        node("fc2/h", "BiasAdd", ["fc2/MatMul", b2]).attr["T"].type = types_pb2.DT_FLOAT
        node("final_softmax", "Softmax", ["fc2/h"]).attr["T"].type = types_pb2.DT_FLOAT
    with open(path, "wb") as f:
        f.write(gd.SerializeToString())
