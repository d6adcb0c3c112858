"""Measured upper bound of what producer fusion / L2 blocking of fc1's A operand could return.

    python profiles/fusion_bound_probe.py            # product library: the step as it is
    RODC_PROBE_A_WRAP=256 python profiles/fusion_bound_probe.py --probe
        # probe build (-DRODC_PROBES): the 48-px stage keeps its whole A operand inside 256 rows (75 MB, L2 resident):
        # the feature kernel still executes every instruction and every store, fc1 still executes every TMA load and
        # every MMA, but the 11.5 GB per chunk of A traffic never reach DRAM.  Results of that stage are WRONG.

Both runs time the bench's resident step (32 frames of 640x480, 10 steps, L2 flushed between steps) and print the
per-kernel table of the library's event pairs.  The difference is everything a scheme that removes A's HBM traffic
could gain while the producer still pays its instructions.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--probe", action="store_true")
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from rapidobjectdetectionusingcascadedcnns_b200 import _lib, build
    if args.probe:
        _lib.use_library(build.LIB_PROBES)
    import rapidobjectdetectionusingcascadedcnns_b200 as pkg
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    from rapidobjectdetectionusingcascadedcnns_b200.engine import CascadeEngine
    pkg.cf.reset()
    eng = CascadeEngine(synth.xavier_cascade())
    d = torch.from_numpy(np.stack([synth.synth_frame(480, 640, j) for j in range(args.frames)])).cuda()
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        res = eng.run_frames(d)
    L = _lib.lib()

    def timed(steps):
        tot = 0.0
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = eng.run_frames(d); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / steps, r
    ms, res = timed(args.steps)
    L.rodc_profile_enable(2)
    ms2, _ = timed(2)
    L.rodc_profile_collect(None, None, None, None, None, None)
    L.rodc_profile_enable(0)
    n = 64
    kc, sub = (C.c_int32 * n)(), (C.c_int32 * n)()
    kms, kby, kfl = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
    kln, krec = (C.c_int64 * n)(), (C.c_int64 * n)()
    rows = L.rodc_profile_table(n, kc, sub, kms, kby, kfl, kln, krec)
    table = sorted(({"kernel": L.rodc_kernel_class_name(int(kc[i])).decode() + (" S=%d" % sub[i] if sub[i] else ""),
                     "ms_per_step": kms[i] / 2} for i in range(rows)), key=lambda r: -r["ms_per_step"])
    print(json.dumps({"probe": bool(args.probe), "a_wrap": os.environ.get("RODC_PROBE_A_WRAP"), "frames": args.frames,
                      "ms_per_step": ms, "frames_per_s": args.frames / ms * 1e3, "stage_kept": res.stage_kept,
                      "kernels": table[:8]}))


if __name__ == "__main__":
    main()
