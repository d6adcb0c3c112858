import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import rapidobjectdetectionusingcascadedcnns_b200 as pkg
from rapidobjectdetectionusingcascadedcnns_b200 import synth
from rapidobjectdetectionusingcascadedcnns_b200.engine import CascadeEngine
eng = CascadeEngine(synth.xavier_cascade())
d = torch.from_numpy(synth.synth_frame(480, 640, 0)[None]).cuda()
for _ in range(5): eng.run_frames(d)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): eng.run_frames(d)
torch.cuda.synchronize(); print("B=1 ms/frame", (time.perf_counter() - t0) / 20 * 1e3)
