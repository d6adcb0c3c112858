import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "meta" in d:
        d["meta"] = json.loads(str(d["meta"]))
    return d


CASCADE_CASES = ["cascade_small", "cascade_small_reject", "cascade_640x480", "cascade_factor12_enlarge",
                 "cascade_conf_avg", "cascade_conf_mult"]
# reference options off by default: resize-then-crop next-stage input, its ..._iff_not_cached variant
# (window_scale_factor 2 so that both branches are taken), one input size for all stages
OPTION_CASES = ["cascade_scaled_patches", "cascade_iff_not_cached", "cascade_equal_dims"]


def case_inputs(name):
    """Frames, weights (with the calibrated offsets) and the oracle Config of a golden case."""
    from oracle import cascade as oc
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    g = load_golden(name)
    meta = g["meta"]
    frames = [synth.synth_frame(h, w, s) for (h, w, s) in meta["frames"]]
    if meta["single"]:
        weights = [synth.xavier_stage_weights(2, 48)]
    elif meta.get("sizes"):
        weights = [synth.xavier_stage_weights(i, sz, n_prev_bottleneck=512 * i) for i, sz in enumerate(meta["sizes"])]
    else:
        weights = synth.xavier_cascade()
    for w, o in zip(weights, meta["offsets"]):
        w["fc2_b"][1] += np.float32(o)
    cfg = oc.Config(**meta["cf"])
    return g, frames, weights, cfg


@pytest.fixture(scope="session")
def golden_loader():
    return load_golden
