"""Oracle (test infrastructure): the per-stage TensorFlow graph restated in torch-CPU.

Follows `network/net.py:101-240` (``_set_up_architecture`` /
``_set_up_architecture_middle_custom_net``) and `network/net_builder.py:6-123`:

    X[n,H,W,3] float (raw 0..255, no normalisation at inference,
                      `app/inference_cascade_app.py:200-202`)
    for each conv layer j: conv2d(HWIO, stride s, SAME) + b -> ReLU -> max_pool(k, s, SAME)
    flatten in NHWC order -> x W1 + b1 -> ReLU             ("bottleneck_out[_part1]")
    stage >= 1: concat([own, bottleneck_in], axis 1)        (`network/net.py:140-145`)
    dropout keep=1.0 (identity)                             (`network/net.py:150-153`)
    x W2 + b2 -> softmax                                    ("final_softmax")

**Parity unpinned**: TensorFlow 1.4 cannot run here and the reference holds no
golden vectors for the graph; this restates the published op semantics.  The
fp64 variant is the arbiter for score tolerances; the fp32 variant is what the
``cpu_baseline`` times.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def _same_pad(n: int, k: int, s: int) -> Tuple[int, int]:
    out = int(math.ceil(n / s))
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def n_conv_layers(w: Dict[str, np.ndarray]) -> int:
    j = 0
    while "conv%d_W" % j in w:
        j += 1
    return j


def stage_forward(x_u8: np.ndarray, w: Dict[str, np.ndarray],
                  bottleneck_in: Optional[np.ndarray] = None, *,
                  dtype=torch.float64, conv_stride: int = 1, pool_size: int = 3,
                  pool_stride: int = 1):
    """Returns ``(softmax [n,2], bottleneck_out [n, 512*(i+1)], logits [n,2], features [n,K])``
    as torch tensors of ``dtype``."""
    x = torch.from_numpy(np.ascontiguousarray(x_u8)).to(dtype).permute(0, 3, 1, 2)   # NCHW
    for j in range(n_conv_layers(w)):
        W = torch.from_numpy(w["conv%d_W" % j]).to(dtype).permute(3, 2, 0, 1)        # HWIO -> OIHW
        b = torch.from_numpy(w["conv%d_b" % j]).to(dtype)
        kh, kw = W.shape[2], W.shape[3]
        pt, pb = _same_pad(x.shape[2], kh, conv_stride)
        pl, pr = _same_pad(x.shape[3], kw, conv_stride)
        x = F.conv2d(F.pad(x, (pl, pr, pt, pb)), W, b, stride=conv_stride)
        x = torch.relu(x)
        pt, pb = _same_pad(x.shape[2], pool_size, pool_stride)
        pl, pr = _same_pad(x.shape[3], pool_size, pool_stride)
        x = F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float("-inf")), pool_size, pool_stride)
    feat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)                               # NHWC flatten
    fc1 = torch.relu(feat @ torch.from_numpy(w["fc1_W"]).to(dtype) + torch.from_numpy(w["fc1_b"]).to(dtype))
    if bottleneck_in is not None:
        bi = bottleneck_in if torch.is_tensor(bottleneck_in) else torch.from_numpy(np.ascontiguousarray(bottleneck_in))
        fc1 = torch.cat([fc1, bi.to(dtype)], dim=1)
    logits = fc1 @ torch.from_numpy(w["fc2_W"]).to(dtype) + torch.from_numpy(w["fc2_b"]).to(dtype)
    probs = torch.softmax(logits, dim=1)
    return probs, fc1, logits, feat


def stage_forward_batched(x_u8, w, bottleneck_in=None, *, dtype=torch.float64, batch: int = 1200, **kw):
    """Runs ``stage_forward`` in batches like `app/inference_cascade_app.py:196-215`."""
    n = x_u8.shape[0]
    ps, bs, ls = [], [], []
    for s in range(0, n, batch):
        e = min(s + batch, n)
        bi = None if bottleneck_in is None else bottleneck_in[s:e]
        p, b, l, _ = stage_forward(x_u8[s:e], w, bi, dtype=dtype, **kw)
        ps.append(p); bs.append(b); ls.append(l)
    return torch.cat(ps), torch.cat(bs), torch.cat(ls)
