"""CPU oracle for the cascade-inference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, what the reference
(Johnson145/RapidObjectDetectionUsingCascadedCNNs) computes on the path
pyramid -> windows -> per-stage CNN -> threshold/compaction -> recrop ->
groupRectangles.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package never does, and fails loudly when its CUDA library is missing.

How the oracle is pinned:

* pyramid plan, window grid, pyramid levels, raw windows, ``*_norm`` boxes, the
  cascade loop (threshold / gather / recrop / box building), ``_run_nms`` and
  ``vertically_enlarge_bboxes``: PINNED by running the reference's own Python
  (`data/rectangles.py`, `data/image_info.py`, `app/inference_cascade_app.py`,
  `app/inference_app.py`, `config.py`) in the build container under shims
  (``oracle/ref_shims.py``); the outputs are committed under ``tests/golden/``
  together with the generator ``oracle/gen_golden.py``.
* Pillow ``resize(BILINEAR)``, ``cv2.resize`` and ``cv2.groupRectangles``
  integer restatements: PINNED against the installed Pillow 12.2 / OpenCV 4.13
  (the reference pins Pillow 5.1.0 / OpenCV 3.3.1, which cannot be installed
  here; version drift for these three functions is unverifiable offline).
* the TensorFlow 1.4 stage graph (conv/pool/fc/softmax): **parity unpinned**.
  TensorFlow is absent and uninstallable (no wheel for py3.12, no network) and
  the reference holds no golden vectors for it.  ``oracle/stage_net.py``
  restates the published op semantics (SAME padding, NHWC flatten, softmax) in
  torch-CPU fp32 and fp64; fp64 is the arbiter for score tolerances.
"""
