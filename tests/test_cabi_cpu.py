"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol declared in
include/rodc.h, and its host-only planner reproduces the reference's window table."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import CASCADE_CASES, REPO, case_inputs


@pytest.fixture(scope="module")
def L():
    from rapidobjectdetectionusingcascadedcnns_b200 import _lib
    return _lib.lib()


def test_library_exports_every_declared_symbol(L):
    from rapidobjectdetectionusingcascadedcnns_b200 import _lib
    header = open(os.path.join(REPO, "include", "rodc.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)                  # drop comments
    declared = set(re.findall(r"\b(rodc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(L, name), name
    assert L.rodc_version() == 100


def test_no_cuda_device_fails_loudly(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rapidobjectdetectionusingcascadedcnns_b200 import _lib, native
    a = C.c_int32()
    rc = L.rodc_device_info(C.byref(a), None, None)
    assert rc != 0 and L.rodc_last_error()
    with pytest.raises(_lib.RodcError):
        native.Stage({"conv0_W": np.zeros((3, 3, 3, 32), np.float32), "conv0_b": np.zeros(32, np.float32),
                      "fc1_W": np.zeros((4608, 512), np.float32), "fc1_b": np.zeros(512, np.float32),
                      "fc2_W": np.zeros((512, 2), np.float32), "fc2_b": np.zeros(2, np.float32)})


@pytest.mark.parametrize("name", CASCADE_CASES + ["single_small"])
def test_planner_matches_reference_window_table(name):
    from rapidobjectdetectionusingcascadedcnns_b200 import native
    g, frames, weights, cfg = case_inputs(name)
    win = 48 if g["meta"]["single"] else 12
    rows, scales = [], []
    for fi, fr in enumerate(frames):
        p = native.Plan(fr.shape[1], fr.shape[0], win, cfg.window_scale_factor, cfg.min_window_length)
        assert p.fits
        used = p.nx * p.ny > 0
        assert np.array_equal(p.scale[used], g["level_scale_%d" % fi])          # bit-exact float64 scales
        assert np.array_equal(np.column_stack([p.level_h, p.level_w])[used], g["level_hw_%d" % fi])
        rows.append(np.column_stack([np.full(p.n_windows, fi), p.win_x, p.win_y, p.norm]))
        scales.append(p.scale[p.win_level])
    assert np.array_equal(np.concatenate(rows), g["win_table"])
    assert np.array_equal(np.concatenate(scales), g["win_scale"])


@pytest.mark.parametrize("shape,win,f,ml,levels,windows", [
    ((640, 480), 12, 1.1, 0.075, 24, 5061), ((640, 480), 48, 1.1, 0.075, 24, 3319),
    ((450, 400), 12, 1.1, 0.075, 26, 6461), ((1920, 1080), 12, 1.2, 0.0125, 21, 93645),
    ((3840, 2160), 12, 1.1, 0.0125, 40, 161969), ((450, 400), 12, 1.005, 0.075, 496, 116065)])
def test_planner_counts_of_baseline_configs(shape, win, f, ml, levels, windows):
    """counts measured by running the reference's extract_windows."""
    from oracle import plan as oplan
    from rapidobjectdetectionusingcascadedcnns_b200 import native
    p = native.Plan(shape[0], shape[1], win, f, ml)
    assert (p.n_levels, p.n_windows) == (levels, windows)
    if windows < 100000:
        o = oplan.plan_pyramid(shape[0], shape[1], win, f, ml)
        assert np.array_equal(o.norm, p.norm) and np.array_equal(o.win_x, p.win_x) and np.array_equal(o.win_y, p.win_y)


def test_planner_rejects_bad_arguments(L):
    h = C.c_void_p()
    assert L.rodc_plan_create(0, 10, 12, 1.1, 0.075, C.byref(h)) != 0
    assert L.rodc_plan_create(100, 100, 12, 1.0, 0.075, C.byref(h)) != 0
    assert b"window_scale_factor" in L.rodc_last_error()
    from rapidobjectdetectionusingcascadedcnns_b200 import native
    assert native.Plan(10, 10, 12).n_windows == 0            # frame smaller than the window
    assert L.rodc_compact_scratch_bytes(5061) > 0 and L.rodc_group_scratch_bytes(100, 3) > 0


def test_shipped_library_contains_no_probe_switches():
    """Measurement aids that change numerics (RODC_FC1_PROBE_HI_ONLY, RODC_FC1_KC, RODC_PROBE_A_WRAP) exist only in a
    -DRODC_PROBES build; the shipped librodc.so must not even contain their names."""
    from rapidobjectdetectionusingcascadedcnns_b200 import build
    blob = open(build.build_library(), "rb").read()
    for name in (b"RODC_FC1_PROBE_HI_ONLY", b"RODC_FC1_KC", b"RODC_PROBE_A_WRAP", b"RODC_PROBE_FEAT_OCC3"):
        assert name not in blob, name
    for name in (b"RODC_FC1_2CTA", b"RODC_FC1_BN", b"RODC_FEATURES_GENERIC"):        # bit-identical kernel choices stay
        assert name in blob, name
