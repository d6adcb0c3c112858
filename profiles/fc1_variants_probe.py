"""fc1 (stage 2, K = 73728) timed ALONE through the library's event hook, for one m per process:

    python profiles/fc1_variants_probe.py <m> [burst|sustained]

burst: 1 launch, then 50 ms idle, repeated (no power capping); sustained: 40 launches back to back.
The kernel variant is chosen by the environment (RODC_FC1_2CTA, RODC_FC1_KC, RODC_FC1_BN), which the
library reads once per process.  Prints one JSON line.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from rapidobjectdetectionusingcascadedcnns_b200 import _lib, native, synth  # noqa: E402


def main():
    m = int(sys.argv[1])
    mode = sys.argv[2] if len(sys.argv) > 2 else "burst"
    dev = torch.device("cuda", 0)
    st = native.Stage(synth.xavier_stage_weights(2, 48, n_prev_bottleneck=1024))
    rs = np.random.RandomState(0)
    win = torch.from_numpy(rs.randint(0, 256, (m, 48, 48, 3)).astype(np.uint8)).to(dev)
    src = native.make_source(dense=win)
    bott = torch.empty((m, st.width), dtype=torch.float32, device=dev)
    wsb = st.workspace_bytes(m, native.IMPL_TENSOR)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    L = _lib.lib()
    for _ in range(3):
        native.stage_forward(st, src, 0, m, bott, ws, native.IMPL_TENSOR)
    torch.cuda.synchronize()
    L.rodc_profile_enable(1)
    reps = 10 if mode == "burst" else 40
    for _ in range(reps):
        native.stage_forward(st, src, 0, m, bott, ws, native.IMPL_TENSOR)
        if mode == "burst":
            torch.cuda.synchronize()
            time.sleep(0.05)
    torch.cuda.synchronize()
    a, b, d, e = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    n1, n2 = C.c_int64(), C.c_int64()
    L.rodc_profile_collect(C.byref(a), C.byref(b), C.byref(n1), C.byref(d), C.byref(e), C.byref(n2))
    L.rodc_profile_enable(0)
    ms = a.value / max(n1.value, 1)
    print(json.dumps({"m": m, "mode": mode, "env": {k: v for k, v in os.environ.items() if k.startswith("RODC_FC1")},
                      "fc1_ms_per_launch": ms, "launches": n1.value,
                      "tflops_algorithmic": b.value / max(a.value, 1e-9) / 1e9,
                      "us_per_kblock_per_wave": ms * 1e3 / 1152 / max(1, -(-(-(-m // 128) * 2) // 148))}))


if __name__ == "__main__":
    main()
