"""Runtime evaluation caller (mirror of `app/evaluate_runtime_app.py:10-97`).

Runs the cascade and then the single net over the SAME images and logs the runtimes (the three runtime lines of
`app/abstract_inference_app.py:128-154` end up in ``utils.log.log_cache``).  Like the reference it builds each
inference app from its session key -- i.e. from the ``graph_<key>[_<i>].pb|.npz`` files in
``cf output_graph_dir`` -- runs ``run_inference_on_images(imgs, merge=cf inference_merge)``, then ``clean()``s
the app and clears every image's raw cache so that the second run cannot profit from the first (`:93-97`).

Differences, all forced by scope: the images come from the caller (``img_infos``) instead of
``FileListLoader().sample_image_infos`` (`:38`, the dataset layer is out of scope), ``BaseApp`` bookkeeping is
gone, and the CPU / GPU switch is gone (the reference's own CPU variants are commented out, `:50-55`; this
path has no CPU implementation).  ``cascade_models`` / ``single_models`` inject in-memory stage models instead
of files.
"""
from .. import config as cf
from ..utils import log
from ..utils.time_watcher import TimeWatcher
from .inference_app import InferenceApp
from .inference_cascade_app import InferenceCascadeApp


class EvaluateRuntimeApp:
    def __init__(self, cascade_session_key=None, single_session_key=None, max_positive_test_imgs=None,
                 max_negative_test_imgs=None, img_infos=None, cascade_models=None, single_models=None, run_now=True):
        self._cascade_session_key, self._single_session_key = cascade_session_key, single_session_key
        self._max_positive_test_imgs, self._max_negative_test_imgs = max_positive_test_imgs, max_negative_test_imgs
        self._cascade_models, self._single_models = cascade_models, single_models
        cf.set("cache_dataset", False)                      # `:35`: original images, not cached patches
        if img_infos is None:
            raise ValueError("EvaluateRuntimeApp needs img_infos (the reference samples them with FileListLoader, "
                             "which belongs to the dataset layer and is out of scope)")
        self._img_infos = list(img_infos)                   # sampled once so that both apps see the same files (`:37-38`)
        self.results = {}
        self.runtimes = {}
        if run_now:
            self._main()

    def _main(self):
        log.log("Evaluating the runtime of all cascade/single cpu/gpu combinations.")
        log.log(" .. cascade_session_key: {}".format(self._cascade_session_key))
        log.log(" .. single_session_key: {}".format(self._single_session_key))
        log.log(" .. max_positive_test_imgs: {}".format(self._max_positive_test_imgs))
        log.log(" .. max_negative_test_imgs: {}".format(self._max_negative_test_imgs))
        log.log(" .. total number of actually used images: {}".format(len(self._img_infos)))
        self._run_inference(True, True)
        self._run_inference(True, False)

    def _run_inference(self, enable_gpu: bool, use_cascade: bool):
        """`:57-97` for one (processing unit, model) combination; only the GPU exists here."""
        if not enable_gpu:
            raise NotImplementedError("this path has no CPU implementation (the reference's CPU runs are commented out too)")
        log.log("")
        log.log("~~~~~~~~~~~~~~~~~~~~~~~~~~  {} / {}  ~~~~~~~~~~~~~~~~~~".format(
            "Cascade" if use_cascade else "Single Net", "GPU"))
        if use_cascade:
            app_inference = InferenceCascadeApp(self._cascade_session_key, self._cascade_models)
        else:
            app_inference = InferenceApp(self._single_session_key, self._single_models)
        key = "cascade" if use_cascade else "single"
        tw = TimeWatcher("runtime evaluation: {}".format(key))
        self.results[key] = app_inference.run_inference_on_images(self._img_infos, merge=cf.get("inference_merge"))
        tw.stop()
        self.runtimes[key] = tw.elapsed_seconds
        # tidy up so that a later run does not benefit from anything done in a previous run (`:93-97`)
        app_inference.clean()
        app_inference = None
        for img_info in self._img_infos:
            img_info.clear_raw_img_cache()
