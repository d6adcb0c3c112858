"""`cf` -- the global key/value configuration the hot path reads (mirror of the reference's
`config.py:603-625` public interface ``get / set / reset`` with the same key names and the same
defaults for every key of SURVEY.md App. B).

Unlike the reference module this one has no import side effects (no ``config_local.py``, no
dataset-root validation, no directory creation, no ``CUDA_VISIBLE_DEVICES`` mutation): those
belong to the training/data subsystems, which are out of scope.
"""
import os

import numpy as np

NMS_DISABLED = "NMS_DISABLED"                      # config.py:95-96
NMS_OPENCV = "NMS_OPENCV"
FINAL_CONFIDENCE_CALCULATION_LAST_STEP = "FINAL_CONFIDENCE_CALCULATION_LAST_STEP"   # :118-121
FINAL_CONFIDENCE_CALCULATION_AVG = "FINAL_CONFIDENCE_CALCULATION_AVG"
FINAL_CONFIDENCE_CALCULATION_MULT = "FINAL_CONFIDENCE_CALCULATION_MULT"
MIN_SCORE_FOR_FINAL_CONFIDENCE_CALCULATION_MULT = 0.5001

# librodc-specific switches (not in the reference)
FC1_IMPL_TENSOR = "tensor"        # tcgen05 split-fp16 GEMM
FC1_IMPL_SIMT = "simt"            # fp32 CUDA-core GEMM (cross-check)
LATENCY_AUTO = "auto"             # small batches run as one CUDA graph with device-side survivor counts
LATENCY_ALWAYS = "always"
LATENCY_OFF = "off"


def _defaults():
    cf = dict()
    cf["foreground_confidence_threshold"] = 0.5            # config.py:73
    cf["inference_merge"] = True                           # :77
    cf["min_window_length"] = 0.075                        # :84
    cf["window_scale_factor"] = 1.1                        # :92
    cf["nms"] = NMS_OPENCV                                 # :97
    cf["nms_opencv_min_neighbors"] = 1                     # :101
    cf["default_evaluation_model_cascade"] = "cnn_cascade_for_face_detection_Mkd"   # :105
    cf["default_evaluation_model_single"] = "single_cnn_for_face_detection_SfN"     # :106
    cf["vertically_enlarge_bboxes"] = False                # :109
    cf["multi_threaded_step_preparation"] = False          # :113 (no effect: the re-crop is fused on the GPU)
    cf["final_confidence_calculation"] = FINAL_CONFIDENCE_CALCULATION_LAST_STEP     # :122
    cf["batch_size"] = 1200                                # :128
    cf["max_batch_size"] = cf["batch_size"]                # :133
    cf["max_batch_size_original"] = cf["max_batch_size"]   # :582
    cf["fc1_size"] = 512                                   # :243
    cf["conv_filter_sizes"] = [32]                         # :249
    cf["conv_stride"] = 1                                  # :254
    cf["conv_filter_size"] = 3                             # :258
    cf["pooling_size"] = 3                                 # :261
    cf["pooling_stride"] = 1                               # :262
    cf["cascade_n_nets"] = 3                               # :268
    cf["cascade_increasing_input_dimensions"] = True       # :274 (overwritten at run time from the models)
    cf["reuse_bottlenecks"] = True                         # :277
    cf["append_inception"] = False                         # :287 (Inception splice: out of scope)
    cf["graph_final_inference_layer_name"] = "final_softmax"    # :290-294
    cf["graph_input_inference_layer_name"] = "X"
    cf["graph_input_bottleneck_layer_name"] = "bottleneck_in"
    cf["graph_output_bottleneck_layer_name"] = "bottleneck_out"
    cf["img_width"] = 48                                   # :322
    cf["img_height"] = cf["img_width"]
    cf["img_width_max"] = cf["img_width"]                  # :577-578
    cf["img_height_max"] = cf["img_height"]
    cf["img_dtype"] = np.uint8                             # :326
    cf["cascade_scale_patches_individually"] = True        # :351
    cf["cascade_scale_patches_individually_iff_not_cached"] = False   # :358
    cf["log_window_extraction_details"] = False            # :473
    cf["log_cascade_confidence_details"] = False           # :476
    cf["output_graph_dir"] = os.environ.get("RODC_GRAPH_DIR", os.path.join(os.getcwd(), "output", "graph"))   # :523
    # FDDB evaluation caller (`config.py:479,552-558`); the external Perl / gnuplot scoring is out of scope
    cf["foreground_equals_face"] = True
    _root = os.environ.get("RODC_DATA_ROOT", os.path.join(os.getcwd(), "input"))
    cf["fddb_root_dir"] = os.path.join(_root, "fddb")
    cf["fddb_folds_dir"] = os.path.join(cf["fddb_root_dir"], "FDDB-folds")
    cf["fddb_img_base_dir"] = os.path.join(cf["fddb_root_dir"], "images/original/foreground")
    cf["fddb_detection_output_dir"] = os.path.join(os.getcwd(), "output", "fddb_detection_output")
    cf["session_key"] = "session"
    # librodc
    cf["fc1_impl"] = FC1_IMPL_TENSOR
    cf["max_windows_per_launch"] = 1 << 18                 # frame batches are cut so stage 0 stays below this
    # decode pool of the file pipeline: up to 8 workers (measured on a 16-core B200 host: 8 workers 2 084 frames/s of
    # 450x400 JPEGs, 15 workers 1 780-1 880: PIL's per-image Python overhead serialises on the GIL), never more than
    # the cores this process may use minus the calling thread
    cf["decode_threads"] = max(1, min(8, (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)) - 1))
    cf["latency_mode"] = LATENCY_AUTO                      # engine.run_frames_latency for batches of <= latency_max_windows
    cf["latency_max_windows"] = 16384                      # stage-0 windows (one 640x480 frame has 5061)
    return cf


_cf = _defaults()
_cf_default = _cf.copy()


def get(param_name):
    return _cf[param_name]


def set(param_name, param_value):   # noqa: A001  (the reference's name)
    _cf[param_name] = param_value


def reset():
    """Reset the configuration to its default values (`config.py:619-625`)."""
    global _cf
    _cf = _cf_default.copy()
