"""Stage-model interface: the weights of one cascade stage in the reference's layouts.

The reference defines the per-stage graph in `network/net.py:101-240` / `network/net_builder.py`
and ships it as a frozen GraphDef ``graph_<session>_<i>.pb`` (`app/train_app.py:177-227`).  A
``StageModel`` holds exactly the constants of that graph:

    conv0_W [3,3,3,C] (HWIO)  conv0_b [C]         "conv_layer_0/W|b"
    fc1_W   [S*S*C, 512]      fc1_b  [512]        "bottleneck_out[_part1]/W|b"  (NHWC flatten order)
    fc2_W   [512*(i+1), 2]    fc2_b  [2]          "fc2/W|b"

Files: ``graph_<key>_<i>.npz`` (this package's own container) or the reference's ``.pb``
(read with ``tensorboard.compat.proto``; no TensorFlow needed).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np

from .. import synth



class StageModel:
    """Weights + geometry of one stage.  ``in_size`` must be given for nets whose spatial size changes
    (strides > 1); for the default stride-1 nets it follows from the fc1 shape."""

    def __init__(self, weights: Dict[str, np.ndarray], in_size: Optional[int] = None, conv_stride: int = 1,
                 pool_size: int = 3, pool_stride: int = 1):
        n_conv = 0
        while "conv%d_W" % n_conv in weights:
            n_conv += 1
        if n_conv < 1:
            raise ValueError("stage model lacks conv0_W")
        for k in ("fc1_W", "fc1_b", "fc2_W", "fc2_b"):
            if k not in weights:
                raise ValueError("stage model lacks %s" % k)
        keys = ["conv%d_%s" % (j, t) for j in range(n_conv) for t in ("W", "b")] + ["fc1_W", "fc1_b", "fc2_W", "fc2_b"]
        for k in keys:
            if k not in weights:
                raise ValueError("stage model lacks %s" % k)
        self.weights = {k: np.ascontiguousarray(weights[k], np.float32) for k in keys}
        self.n_conv = n_conv
        self.conv_stride, self.pool_size, self.pool_stride = int(conv_stride), int(pool_size), int(pool_stride)
        cw = self.weights["conv0_W"]
        if cw.ndim != 4 or cw.shape[0] != cw.shape[1] or cw.shape[2] != 3:
            raise ValueError("conv0_W must be HWIO [k,k,3,C] (got %r)" % (cw.shape,))
        self.ksize = int(cw.shape[0])
        self.conv_channels = [int(self.weights["conv%d_W" % j].shape[3]) for j in range(n_conv)]
        self.channels = self.conv_channels[-1]
        k, self.fc1_size = self.weights["fc1_W"].shape

        def flat_size(side):
            for _ in range(n_conv):
                side = -(-side // self.conv_stride)
                side = -(-side // self.pool_stride)
            return side * side * self.channels
        if in_size is None:
            cands = [s_ for s_ in range(1, 1025) if flat_size(s_) == k]
            if not cands:
                raise ValueError("fc1 rows %d do not match any input size" % k)
            in_size = cands[0] if len(cands) == 1 or self.conv_stride * self.pool_stride == 1 else None
            if in_size is None:
                raise ValueError("ambiguous input size for a strided net: pass in_size")
        self.in_size = int(in_size)
        if flat_size(self.in_size) != k:
            raise ValueError("fc1 rows %d are not the flattened size of a %dx%d input" % (k, self.in_size, self.in_size))
        self.n_prev = int(self.weights["fc2_W"].shape[0]) - int(self.fc1_size)
        if self.n_prev < 0 or self.weights["fc2_W"].shape[1] != 2:
            raise ValueError("bad fc2 shape %r" % (self.weights["fc2_W"].shape,))

    @property
    def is_default_architecture(self) -> bool:
        return (self.n_conv == 1 and self.conv_channels == [32] and self.ksize == 3 and self.conv_stride == 1
                and self.pool_size == 3 and self.pool_stride == 1)

    # X placeholder shape of the exported graph: [None, H, W, 3] (`app/train_app.py:203-205`)
    @property
    def supported_img_width(self) -> int:
        return self.in_size

    @property
    def supported_img_height(self) -> int:
        return self.in_size

    @property
    def bottleneck_out_width(self) -> int:
        return self.fc1_size + self.n_prev

    def save(self, path: str):
        """``.npz``: this package's container; ``.pb``: a GraphDef with the reference's node names and op
        attributes (`network/graph_pb.py`)."""
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        if path.endswith(".pb"):
            from .graph_pb import write_stage_pb
            write_stage_pb(path, self.weights, self.n_prev > 0, self.in_size, self.conv_stride, self.pool_size,
                           self.pool_stride)
            return
        with open(path, "wb") as f:
            np.savez(f, in_size=self.in_size, conv_stride=self.conv_stride, pool_size=self.pool_size,
                     pool_stride=self.pool_stride, **self.weights)

    @staticmethod
    def load(path: str) -> "StageModel":
        if path.endswith(".pb"):
            # weights from the Const nodes, geometry from the X placeholder / Conv2D / MaxPool attributes
            from .graph_pb import read_stage_graph
            weights, geo = read_stage_graph(path)
            return StageModel(weights, **geo)
        with np.load(path) as z:
            d = {k: z[k] for k in z.files}
        geo = {k: int(d.pop(k)) for k in ("in_size", "conv_stride", "pool_size", "pool_stride") if k in d}
        return StageModel(d, **geo)


def xavier_cascade_models(n_stages: int = 3, img_max: int = 48,
                          fg_bias_offsets: Optional[List[float]] = None) -> List[StageModel]:
    """Random-init cascade of the reference architecture (SURVEY.md §8d weight recipe)."""
    return [StageModel(w) for w in synth.xavier_cascade(n_stages, img_max, fg_bias_offsets=fg_bias_offsets)]


def model_paths(graph_dir: str, key: str, index: Optional[int]) -> List[str]:
    """Candidate file names for a stage: the reference's ``.pb`` first, then ``.npz``
    (`app/inference_app.py:38-40`, `app/inference_cascade_app.py:89-117`)."""
    base = "graph_%s" % key if index is None else "graph_%s_%d" % (key, index)
    return [os.path.join(graph_dir, base + ext) for ext in (".pb", ".npz")]
