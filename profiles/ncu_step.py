"""One step of the bench workload (BASELINE config 2, `--frames` frames) for `ncu --set full`:
warm-up step outside the profiled range, then ONE step between cudaProfilerStart/Stop.

    python profiles/ncu_step.py [--frames 8]                                     # plain run first
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'rodc' -o gpurun_out/r2_step python profiles/ncu_step.py --frames 8

8 frames keep the capture short (about 35 launches) while the dominant shapes are the bench's: the first stage-2
chunk is a full 18 944-window chunk (features S = 48, re-crop, fc1 pair kernel with two waves).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rapidobjectdetectionusingcascadedcnns_b200 as pkg  # noqa: E402
from rapidobjectdetectionusingcascadedcnns_b200 import synth  # noqa: E402
from rapidobjectdetectionusingcascadedcnns_b200.engine import CascadeEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    args = ap.parse_args()
    pkg.cf.reset()
    eng = CascadeEngine(synth.xavier_cascade())
    d = torch.from_numpy(np.stack([synth.synth_frame(480, 640, j) for j in range(args.frames)])).cuda()
    res = eng.run_frames(d)                       # warm-up: plans uploaded, workspace allocated
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    res = eng.run_frames(d)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("frames %d windows %d stage_in %s kept %s boxes %d" % (args.frames, res.n_windows, res.stage_in, res.stage_kept,
                                                                 res.boxes.shape[0]))


if __name__ == "__main__":
    main()
