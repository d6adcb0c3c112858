"""Seeded synthetic frames and stage weights (SURVEY.md §8(d) recipes).

There is no network and no dataset on the build/bench machines, so every
benchmark and parity case runs on synthetic frames and random-init weights of
the reference architecture.  Both recipes are pure NumPy and deterministic, so
the oracle, the golden-vector generator and the CUDA path all see the same
bytes.

* Frames: multi-octave value noise (white noise would average to flat grey
  under the 4-37x antialiased pyramid and make every window identical).
* Weights: Xavier-uniform like the reference's initialisers
  (`network/net_builder.py:33-52,78-92`: ``xavier_initializer[_conv2d]`` for W,
  ``constant_initializer(0.0)`` for b), stage ``i`` seeded ``1000 + i``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

N_OCTAVES = 6


def synth_frame(height: int, width: int, seed: int) -> np.ndarray:
    """``[H, W, 3]`` uint8 multi-octave value-noise frame."""
    rs = np.random.RandomState(seed)
    acc = np.zeros((height, width, 3), dtype=np.float64)
    for o in range(1, N_OCTAVES + 1):
        cell = 1 << o
        coarse = rs.uniform(0.0, 1.0, size=((height >> o) + 2, (width >> o) + 2, 3))
        up = np.kron(coarse, np.ones((cell, cell, 1)))[:height, :width]
        acc += float(cell) * up
    lo, hi = acc.min(), acc.max()
    img = (acc - lo) / (hi - lo) * 255.0
    return img.astype(np.uint8)


def frame_set(shapes: Sequence[Sequence[int]], base_seed: int) -> List[np.ndarray]:
    """Frame ``j`` of a set uses ``seed = base_seed + j``; ``shapes[j] = (H, W)``."""
    return [synth_frame(int(h), int(w), base_seed + j) for j, (h, w) in enumerate(shapes)]


def fddb_shape(j: int) -> (int, int):
    """Config-3 frame sizes: ``W = 450 - 10 (j mod 5)``, ``H = 400 - 8 (j mod 7)``."""
    return 400 - 8 * (j % 7), 450 - 10 * (j % 5)


def _xavier(rs: np.random.RandomState, shape, fan_in: int, fan_out: int) -> np.ndarray:
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return rs.uniform(-limit, limit, size=shape).astype(np.float32)


def stage_input_sizes(n_stages: int, img_max: int = 48) -> List[int]:
    """Input side of stage ``i``: halve ``img_max`` ``n-1-i`` times with ``int(w/2)``
    (`app/train_cascade_app.py:203-228`)."""
    out = []
    for i in range(n_stages):
        w = img_max
        for _ in range(n_stages - i - 1):
            w = int(w / 2)
        out.append(w)
    return out


def xavier_stage_weights(stage: int, in_size: int, *, n_prev_bottleneck: int = 0,
                         conv_filters: Sequence[int] = (32,), ksize: int = 3,
                         fc1_size: int = 512, n_classes: int = 2,
                         fg_bias_offset: float = 0.0,
                         seed: Optional[int] = None, conv_stride: int = 1, pool_stride: int = 1) -> Dict[str, np.ndarray]:
    """Random-init weights of one stage net in the reference's layouts.

    conv ``W[kh,kw,Cin,Cout]`` (HWIO), fc1 ``W[K,512]`` with NHWC flatten order,
    fc2 ``W[512 + n_prev_bottleneck, 2]``; biases zero.  ``fg_bias_offset`` is the
    documented calibration knob added to ``fc2.b[1]`` for configs that need a
    target pass rate (the same offset feeds the oracle and the GPU path).
    SAME conv/pool: every layer maps a side ``n`` to ``ceil(n / stride)`` (stride 1 = reference default keeps
    ``in_size`` up to the flatten).
    """
    rs = np.random.RandomState(1000 + stage if seed is None else seed)
    w: Dict[str, np.ndarray] = {}
    cin = 3
    side = in_size
    for j, cout in enumerate(conv_filters):
        w["conv%d_W" % j] = _xavier(rs, (ksize, ksize, cin, cout), ksize * ksize * cin, ksize * ksize * cout)
        w["conv%d_b" % j] = np.zeros((cout,), np.float32)
        cin = cout
        side = -(-side // conv_stride)
        side = -(-side // pool_stride)
    k = side * side * cin
    w["fc1_W"] = _xavier(rs, (k, fc1_size), k, fc1_size)
    w["fc1_b"] = np.zeros((fc1_size,), np.float32)
    k2 = fc1_size + n_prev_bottleneck
    w["fc2_W"] = _xavier(rs, (k2, n_classes), k2, n_classes)
    b2 = np.zeros((n_classes,), np.float32)
    b2[1] += np.float32(fg_bias_offset)
    w["fc2_b"] = b2
    return w


def xavier_cascade(n_stages: int = 3, img_max: int = 48, fc1_size: int = 512,
                   fg_bias_offsets: Optional[Sequence[float]] = None,
                   reuse_bottlenecks: bool = True) -> List[Dict[str, np.ndarray]]:
    """All stage weight dicts of a default-architecture cascade (12/24/48)."""
    sizes = stage_input_sizes(n_stages, img_max)
    out = []
    for i, s in enumerate(sizes):
        off = 0.0 if fg_bias_offsets is None else float(fg_bias_offsets[i])
        prev = fc1_size * i if reuse_bottlenecks else 0
        out.append(xavier_stage_weights(i, s, n_prev_bottleneck=prev, fc1_size=fc1_size, fg_bias_offset=off))
    return out
