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

KEYS = ("conv0_W", "conv0_b", "fc1_W", "fc1_b", "fc2_W", "fc2_b")


class StageModel:
    def __init__(self, weights: Dict[str, np.ndarray]):
        for k in KEYS:
            if k not in weights:
                raise ValueError("stage model lacks %s" % k)
        if "conv1_W" in weights:
            raise NotImplementedError("multi-conv-layer stage nets are not built yet (SURVEY.md §8f #4)")
        self.weights = {k: np.ascontiguousarray(weights[k], np.float32) for k in KEYS}
        cw = self.weights["conv0_W"]
        if cw.shape[:3] != (3, 3, 3):
            raise NotImplementedError("only the default 3x3 RGB conv is built (got %r)" % (cw.shape,))
        self.channels = int(cw.shape[3])
        k, self.fc1_size = self.weights["fc1_W"].shape
        self.in_size = int(round((k // self.channels) ** 0.5))
        if self.in_size * self.in_size * self.channels != k:
            raise ValueError("fc1 rows %d are not S*S*%d" % (k, self.channels))
        self.n_prev = int(self.weights["fc2_W"].shape[0]) - int(self.fc1_size)
        if self.n_prev < 0 or self.weights["fc2_W"].shape[1] != 2:
            raise ValueError("bad fc2 shape %r" % (self.weights["fc2_W"].shape,))

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
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "wb") as f:
            np.savez(f, **self.weights)

    @staticmethod
    def load(path: str) -> "StageModel":
        if path.endswith(".pb"):
            from .graph_pb import read_stage_pb
            return StageModel(read_stage_pb(path))
        with np.load(path) as z:
            return StageModel({k: z[k] for k in z.files})


def xavier_cascade_models(n_stages: int = 3, img_max: int = 48,
                          fg_bias_offsets: Optional[List[float]] = None) -> List[StageModel]:
    """Random-init cascade of the reference architecture (SURVEY.md §8d weight recipe)."""
    return [StageModel(w) for w in synth.xavier_cascade(n_stages, img_max, fg_bias_offsets=fg_bias_offsets)]


def model_paths(graph_dir: str, key: str, index: Optional[int]) -> List[str]:
    """Candidate file names for a stage: the reference's ``.pb`` first, then ``.npz``
    (`app/inference_app.py:38-40`, `app/inference_cascade_app.py:89-117`)."""
    base = "graph_%s" % key if index is None else "graph_%s_%d" % (key, index)
    return [os.path.join(graph_dir, base + ext) for ext in (".pb", ".npz")]
