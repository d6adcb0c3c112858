"""Where does the host time of the config-3 end-to-end path go?  cProfile of app.run_inference_rows over N pinned
FDDB-shaped frames (main thread) + wall-clock of the resident equivalent."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rapidobjectdetectionusingcascadedcnns_b200 as pkg  # noqa: E402
from rapidobjectdetectionusingcascadedcnns_b200 import synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1300
pkg.cf.reset()
pkg.cf.set("max_windows_per_launch", 1 << 22)
w = synth.xavier_cascade(fg_bias_offsets=[-40.0, -85.0, -24.0])
app = pkg.InferenceCascadeApp("p", stage_models=[pkg.StageModel(x) for x in w])
frames = [synth.synth_frame(*synth.fddb_shape(j), 100 + j) for j in range(N)]
infos = [pkg.ArrayImageInfo(f, "f_%d.png" % j, pin=True) for j, f in enumerate(frames)]
for _ in range(2):
    rows = app.run_inference_rows(infos)
torch.cuda.synchronize()
t0 = time.perf_counter(); rows = app.run_inference_rows(infos); torch.cuda.synchronize(); t_e2e = time.perf_counter() - t0
eng = app._engine
chunks = [frames[i:i + 650] for i in range(0, N, 650)]
res = [eng.upload_mixed(c) for c in chunks]
torch.cuda.synchronize()
for mf in res:
    eng.run_mixed_device(mf)
torch.cuda.synchronize()
t0 = time.perf_counter()
for mf in res:
    eng.run_mixed_device(mf)
torch.cuda.synchronize(); t_res = time.perf_counter() - t0
t0 = time.perf_counter(); up = [eng.upload_mixed(c) for c in chunks]; torch.cuda.synchronize(); t_up = time.perf_counter() - t0
print("frames %d  e2e %.1f ms  resident %.1f ms  upload alone %.1f ms  boxes %d" % (N, t_e2e * 1e3, t_res * 1e3, t_up * 1e3, rows.shape[0]))
pr = cProfile.Profile()
pr.enable(); app.run_inference_rows(infos); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
