"""Runtime evaluation caller (mirror of `app/evaluate_runtime_app.py:10-97`): runs the cascade and
then the single net over the same images and logs the runtimes."""
from .. import config as cf
from ..utils import log
from .inference_app import InferenceApp
from .inference_cascade_app import InferenceCascadeApp


class EvaluateRuntimeApp:
    def __init__(self, cascade_session_key=None, single_session_key=None, img_infos=None,
                 cascade_models=None, single_models=None, run_now=True):
        self._cascade_session_key, self._single_session_key = cascade_session_key, single_session_key
        self._cascade_models, self._single_models = cascade_models, single_models
        self._img_infos = list(img_infos or [])
        self.results = {}
        if run_now:
            self._main()

    def _main(self):
        log.log("Evaluating the runtime of the cascade and the single net.")
        log.log(" .. total number of actually used images: {}".format(len(self._img_infos)))
        self._run_inference(True)
        self._run_inference(False)

    def _run_inference(self, use_cascade: bool):
        log.log("~~~~~~~~~~~~~~  {} / GPU  ~~~~~~~~~~~~~~".format("Cascade" if use_cascade else "Single Net"))
        if use_cascade:
            app = InferenceCascadeApp(self._cascade_session_key, self._cascade_models)
        else:
            app = InferenceApp(self._single_session_key, self._single_models)
        self.results["cascade" if use_cascade else "single"] = app.run_inference_on_images(
            self._img_infos, merge=cf.get("inference_merge"))
        app.clean()
        for info in self._img_infos:
            info.clear_raw_img_cache()
