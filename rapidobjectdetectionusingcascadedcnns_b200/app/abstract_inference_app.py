"""Common root of the inference apps (mirror of `app/abstract_inference_app.py:16-191`).

Same public methods and error behaviour; the body differs: instead of extracting windows with
a thread pool and concatenating them on the host (`:56-87`), images are grouped by shape and
handed to the device engine as frame batches.
"""
import abc
import traceback
from typing import List

import numpy as np

from .. import config as cf
from ..data.rectangles import LabeledBoundingBox, Window
from ..utils import log
from ..utils.time_watcher import TimeWatcher


class AbstractInferenceApp(metaclass=abc.ABCMeta):

    @abc.abstractmethod
    def run_inference_on_windows(self, windows_info: List[Window], windows_raw) -> List[LabeledBoundingBox]:
        return

    @abc.abstractmethod
    def run_inference_on_image(self, image) -> List[LabeledBoundingBox]:
        return

    @abc.abstractmethod
    def _init_tf(self):
        return

    @abc.abstractmethod
    def _run_on_frame_batch(self, infos, frames) -> List[List[LabeledBoundingBox]]:
        """Whole path for a list of host frames (any mix of shapes); one result list per frame."""
        return

    def _extract_windows(self, img, convert_raw_to_np=True):
        """`:158-183`: wrapper of ``Window.extract_windows`` that swallows per-image errors (-> ``([], [])``)."""
        try:
            windows_raw, windows_info = Window.extract_windows(img, convert_raw_to_np)
            if len(windows_raw) < 1:
                raise ValueError("Could not extract any windows from the given image")
            return windows_raw, windows_info
        except FileNotFoundError:
            log.log(" .. Skipped {}, because the file could not be found".format(getattr(img, "path_resized", img)))
            return [], []
        except Exception:
            log.log(" .. Skipped {}, because of an unexpected error:\n{}".format(getattr(img, "path_resized", img),
                                                                               traceback.format_exc()))
            return [], []

    def _load_frame(self, img):
        """Decode/fetch one image; errors are swallowed like `:165-183` (image yields [])."""
        try:
            arr = img.raw_original(True)
            if arr.ndim != 3 or arr.shape[2] != 3:
                raise ValueError("expected an RGB image")
            return np.ascontiguousarray(arr, dtype=np.uint8)
        except FileNotFoundError:
            log.log(" .. Skipped {}, because the file could not be found".format(getattr(img, "path_resized", img)))
        except Exception:
            log.log(" .. Skipped {}, because of an unexpected error:\n{}".format(getattr(img, "path_resized", img),
                                                                               traceback.format_exc()))
        return None

    def run_inference_on_images(self, images: List, merge=True) -> List[List[LabeledBoundingBox]]:
        """`app/abstract_inference_app.py:38-156`.  The outer list has one inner list per input
        image (input order); each inner list holds the foreground boxes of that image."""
        self._init_tf()
        all_results = [[] for _ in images]
        timer_multiple = TimeWatcher("inference_img_multiple: {} imgs".format(len(images)))
        timer_extracting = None
        if merge:
            timer_extracting = TimeWatcher("extract windows from all images and merge them")
            log.log(" -> extract")
            usable = []                                       # (input index, image, frame) with at least one window
            for idx, img in enumerate(images):
                arr = self._load_frame(img)
                if arr is None:
                    continue
                try:
                    n_win = self._windows_per_frame(arr.shape[1], arr.shape[0])      # raises if a window leaves its level
                    if n_win < 1:
                        raise ValueError("Could not extract any windows from the given image")
                except ValueError as e:
                    log.log(" .. Skipped {}: {}".format(getattr(img, "path_resized", img), e))   # `:165-183`
                    continue
                usable.append((idx, img, arr, n_win))
            timer_extracting.stop()
            # cut into launches of at most cf['max_windows_per_launch'] stage-0 windows, input order kept
            budget = int(cf.get("max_windows_per_launch"))
            chunk, acc = [], 0
            chunks = []
            for item in usable:
                if chunk and acc + item[3] > budget:
                    chunks.append(chunk); chunk, acc = [], 0
                chunk.append(item); acc += item[3]
            if chunk:
                chunks.append(chunk)
            for chunk in chunks:
                per_frame = self._run_on_frame_batch([i for _, i, _, _ in chunk], [a for _, _, a, _ in chunk])
                for (idx, _, _, _), boxes in zip(chunk, per_frame):
                    all_results[idx] = boxes
            usable = len(usable)
            if usable == 0:
                raise ValueError("Could not extract any windows from the given image")
        else:
            for idx, img in enumerate(images):
                img_results = []
                try:
                    timer_single = TimeWatcher("inference_img_single")
                    img_results = self.run_inference_on_image(img)
                    timer_single.stop()
                except FileNotFoundError:
                    log.log(" .. Skipped {}, because the file could not be found".format(getattr(img, "path_resized", img)))
                except Exception:
                    log.log(" .. Skipped {}, because of an unexpected error:\n{}".format(
                        getattr(img, "path_resized", img), traceback.format_exc()))
                all_results[idx] = img_results
        timer_multiple.stop()
        n = float(max(len(images), 1))
        if merge:
            t_ext = timer_extracting.elapsed_seconds
            log.log("Runtime window extraction: {} images in {} (avg: {}).".format(
                len(images), TimeWatcher.seconds_to_str(t_ext), TimeWatcher.seconds_to_str(t_ext / n)))
            t_inf = timer_multiple.elapsed_seconds - t_ext
            log.log("Runtime inference only: {} images in {} (avg: {}).".format(
                len(images), TimeWatcher.seconds_to_str(t_inf), TimeWatcher.seconds_to_str(t_inf / n)))
        t_all = timer_multiple.elapsed_seconds
        log.log("Runtime inference including window extraction: {} images in {} (avg: {}).".format(
            len(images), TimeWatcher.seconds_to_str(t_all), TimeWatcher.seconds_to_str(t_all / n)))
        return all_results

    @abc.abstractmethod
    def _windows_per_frame(self, w: int, h: int) -> int:
        return
