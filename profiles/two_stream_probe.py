"""Does driving two half-batches on two CUDA streams (two host threads) beat one full batch on one stream?
(features of one half could fill the SMs left idle by the tail waves of the other half's fc1.)"""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapidobjectdetectionusingcascadedcnns_b200 import synth
from rapidobjectdetectionusingcascadedcnns_b200.engine import CascadeEngine

B = 32
w = synth.xavier_cascade()
frames = torch.from_numpy(np.stack([synth.synth_frame(480, 640, j) for j in range(B)])).cuda()
e0, e1 = CascadeEngine(w), CascadeEngine(w)
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()

def one():
    e0.run_frames(frames)

def two():
    def work(e, s, d):
        with torch.cuda.stream(s):
            e.run_frames(d)
    t0 = threading.Thread(target=work, args=(e0, s0, frames[:B // 2]))
    t1 = threading.Thread(target=work, args=(e1, s1, frames[B // 2:]))
    t0.start(); t1.start(); t0.join(); t1.join()

for name, fn in (("one stream, 32 frames", one), ("two streams, 16 + 16 frames", two), ("one stream, 32 frames", one),
                 ("two streams, 16 + 16 frames", two)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(8): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 8
    print("%-30s %.2f ms  %.1f frames/s" % (name, dt * 1e3, B / dt))
