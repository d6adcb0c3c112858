"""FDDB caller regime: 450x400 frames, window_scale_factor 1.005 -> 496 levels,
116 065 windows per frame.  Times K1 alone and the whole path per frame (CUDA events)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rapidobjectdetectionusingcascadedcnns_b200 as pkg
from rapidobjectdetectionusingcascadedcnns_b200 import native, synth
from rapidobjectdetectionusingcascadedcnns_b200.engine import CascadeEngine

pkg.cf.reset(); pkg.cf.set("window_scale_factor", 1.005)
B = 4
frames = torch.from_numpy(np.stack([synth.synth_frame(400, 450, 100 + j) for j in range(B)])).cuda()
eng = CascadeEngine(synth.xavier_cascade(fg_bias_offsets=[-60.0, -90.0, -25.0]))
plan = eng.plan_for(450, 400)
print("levels", plan.n_levels, "windows", plan.n_windows, "pyramid MB/frame", plan.pyramid_bytes / 1e6, "scratch MB/frame", plan.scratch_bytes / 1e6)

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
pyr = torch.empty((B, plan.pyramid_bytes), dtype=torch.uint8, device="cuda")
scr = torch.empty((B, plan.scratch_bytes), dtype=torch.uint8, device="cuda")
print("K1 pyramid      %.3f ms / frame" % (t(lambda: native.pyramid(plan, frames, pyr, scr)) / B))
res = eng.run_frames(frames)
print("stages in/kept", res.stage_in, res.stage_kept, "boxes", res.boxes.shape[0])
print("whole path      %.3f ms / frame" % (t(lambda: eng.run_frames(frames)) / B))
