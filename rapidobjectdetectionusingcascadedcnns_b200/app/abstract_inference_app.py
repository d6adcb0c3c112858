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

    def _frame_size(self, img):
        """(height, width) of an image WITHOUT decoding it when the object can tell (file header, array shape);
        errors are swallowed like `:165-183` (the image yields [])."""
        try:
            h, w = getattr(img, "img_height_original", None), getattr(img, "img_width_original", None)
            if h is None or w is None:
                arr = img.raw_original(True)
                h, w = arr.shape[0], arr.shape[1]
            return int(h), int(w)
        except FileNotFoundError:
            log.log(" .. Skipped {}, because the file could not be found".format(getattr(img, "path_resized", img)))
        except Exception:
            log.log(" .. Skipped {}, because of an unexpected error:\n{}".format(getattr(img, "path_resized", img),
                                                                               traceback.format_exc()))
        return None

    def _prepare_chunk(self, chunk, pool, slot: int):
        """Decode the images of one chunk (in parallel when a pool is given), drop the ones that fail and upload the
        rest on the CURRENT stream.  Returns ``(input indices, images, device payload, None)``."""
        imgs = [i for _, i, _ in chunk]
        decode = pool is not None and any(getattr(i, "_array", None) is None for i in imgs)     # else: nothing to decode
        frames = list(pool.map(self._load_frame, imgs)) if decode else [self._load_frame(i) for i in imgs]
        keep = [k for k, f in enumerate(frames) if f is not None]
        if not keep:
            return [], [], None, None
        if slot:
            self._engine.staging_slot = slot
        payload = self._upload_frame_batch([frames[k] for k in keep])
        return [chunk[k][0] for k in keep], [imgs[k] for k in keep], payload, None

    def _chunk_feeder(self, chunks):
        """Generator over prepared chunks of FILES: a feeder thread decodes them with a pool of ``cf decode_threads``
        workers and uploads from two alternating pinned staging buffers on its own stream, up to two chunks ahead of
        the consumer.  (In-memory frames do not come here, see ``_run_chunks_in_memory``.)"""
        import torch
        device = self._engine.device
        import queue
        import sys
        import threading
        from concurrent.futures import ThreadPoolExecutor
        q = queue.Queue(maxsize=2)
        stop = threading.Event()

        def work():
            try:
                torch.cuda.set_device(device)
                side = torch.cuda.Stream(device=device)
                n_thr = max(1, int(cf.get("decode_threads")))
                with ThreadPoolExecutor(max_workers=n_thr) as pool:
                    for k, chunk in enumerate(chunks):
                        if stop.is_set():
                            break
                        with torch.cuda.stream(side):
                            ids, infos, payload, _ = self._prepare_chunk(chunk, pool, 1 + (k & 1))
                            ev = torch.cuda.Event()
                            ev.record(side)
                        q.put((ids, infos, payload, ev))
            except BaseException as e:      # noqa: BLE001 -- handed to the consumer, which re-raises it
                q.put(e)
            q.put(None)

        # the launch loop of the caller and the feeder's Python code share the GIL: hand it over every 0.2 ms instead
        # of every 5 ms while both are active, or each side stalls the other for milliseconds at a time
        old_interval = sys.getswitchinterval()
        sys.setswitchinterval(2e-4)
        t = threading.Thread(target=work, name="rodc-chunk-feeder", daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                yield item
        finally:
            stop.set()
            while t.is_alive():             # unblock a feeder that waits on the bounded queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    t.join(timeout=0.05)
            sys.setswitchinterval(old_interval)

    def _run_chunks_in_memory(self, chunks, as_rows: bool):
        """Several chunks of frames that need no decoding: no thread.  While the GPU works on stage 0 of chunk k (tens of
        milliseconds for a full chunk) the host prepares chunk k + 1 and issues its asynchronous upload on a side stream
        -- the engine calls back right after stage 0 has been enqueued, before the first count read.  (A feeder THREAD
        for in-memory frames was measured: its Python loops and the caller's launch loop take turns on the GIL at the
        interpreter's 5 ms switch interval and the GPU starves -- 0.78 instead of 0.95 of the resident rate.)"""
        import torch
        side = torch.cuda.Stream(device=self._engine.device)

        def prepare(k):
            with torch.cuda.stream(side):
                ids, infos, payload, _ = self._prepare_chunk(chunks[k], None, 1 + (k & 1))
                ev = torch.cuda.Event()
                ev.record(side)
            return ids, infos, payload, ev
        nxt = prepare(0)
        for k in range(len(chunks)):
            cur, nxt = nxt, None
            state = {}

            def ahead(k=k, state=state):
                if k + 1 < len(chunks):
                    state["next"] = prepare(k + 1)
            ids, infos, payload, ev = cur
            if ids:
                self._wait_upload(ev, payload)
                done = self._run_uploaded(infos, payload, as_rows=as_rows, while_busy=ahead)
            else:
                done = (np.zeros((0,), np.int64), np.zeros((0, 5), np.float64)) if as_rows else []
            if "next" not in state:                         # the run took a path without the callback (latency graph)
                ahead()
            nxt = state.get("next")
            yield ids, done

    @staticmethod
    def _wait_upload(event, payload):
        """Make the compute stream wait for the feeder's copies and tell the caching allocator that the uploaded
        buffers are now used on this stream."""
        import torch
        cur = torch.cuda.current_stream()
        cur.wait_event(event)
        kind, data = payload
        (data if kind == "same" else data.flat).record_stream(cur)

    def run_inference_rows(self, images: List) -> np.ndarray:
        """``run_inference_on_images(images, merge=True)`` with the detections as ONE array ``[K,6]`` of rows
        ``(input index, xmin, ymin, xmax, ymax, confidence)`` sorted by input index (per-image order kept): what
        `dist.run_sharded` exchanges.  No box object and no per-image list is created."""
        return self.run_inference_on_images(images, merge=True, _as_rows=True)

    def run_inference_on_images(self, images: List, merge=True, _as_rows: bool = False) -> List[List[LabeledBoundingBox]]:
        """`app/abstract_inference_app.py:38-156`.  The outer list has one inner list per input
        image (input order); each inner list holds the foreground boxes of that image."""
        self._init_tf()
        if _as_rows and not merge:
            raise ValueError("row output is a merge-mode feature")
        all_results = [[] for _ in images]
        row_parts = []
        timer_multiple = TimeWatcher("inference_img_multiple: {} imgs".format(len(images)))
        timer_extracting = None
        if merge:
            timer_extracting = TimeWatcher("extract windows from all images and merge them")
            log.log(" -> extract")
            # 1. plan: image sizes from the file HEADERS (no decode), window counts from the planner
            usable = []                                       # (input index, image, stage-0 windows)
            per_shape = {}                                    # (h, w) -> window count or the ValueError of that shape
            for idx, img in enumerate(images):
                hw = self._frame_size(img)
                if hw is None:
                    continue
                n_win = per_shape.get(hw)
                if n_win is None:
                    try:
                        n_win = self._windows_per_frame(hw[1], hw[0])      # raises if a window leaves its level
                        if n_win < 1:
                            raise ValueError("Could not extract any windows from the given image")
                    except ValueError as e:
                        n_win = e
                    per_shape[hw] = n_win
                if isinstance(n_win, ValueError):
                    log.log(" .. Skipped {}: {}".format(getattr(img, "path_resized", img), n_win))   # `:165-183`
                    continue
                usable.append((idx, img, n_win))
            timer_extracting.stop()
            # cut into launches of at most cf['max_windows_per_launch'] stage-0 windows, input order kept
            budget = int(cf.get("max_windows_per_launch"))
            chunk, acc = [], 0
            chunks = []
            for item in usable:
                if chunk and acc + item[2] > budget:
                    chunks.append(chunk); chunk, acc = [], 0
                chunk.append(item); acc += item[2]
            if chunk:
                chunks.append(chunk)
            # 2. decode -> upload -> compute.  In-memory images of a single chunk go straight through; otherwise a
            # feeder thread decodes chunk k + 1 with a thread pool and uploads it on its own stream while the
            # GPU works on chunk k (the reference decodes every file once per pyramid scale on the calling
            # thread, `data/image_info.py:228-240`).
            in_memory = all(getattr(i, "_array", None) is not None for _, i, _ in usable)
            n_done = 0
            if in_memory and len(chunks) > 1:
                feed = self._run_chunks_in_memory(chunks, _as_rows)
            elif in_memory:
                feed = [self._prepare_chunk(chunks[0], None, 0)] if chunks else []
            else:
                feed = self._chunk_feeder(chunks)
            for item in feed:
                if isinstance(item, BaseException):
                    raise item
                if len(item) == 2:                             # (input indices, finished result) from the in-memory pipeline
                    ids, done = item
                    if _as_rows:
                        local, rows = done
                        row_parts.append(np.column_stack([np.asarray(ids, np.float64)[local], rows]).reshape(-1, 6))
                    else:
                        for idx, boxes in zip(ids, done):
                            all_results[idx] = boxes
                    n_done += len(ids)
                    continue
                ids, infos, payload, event = item
                if not ids:
                    continue
                if event is not None:
                    self._wait_upload(event, payload)
                if _as_rows:
                    local, rows = self._run_uploaded(infos, payload, as_rows=True)
                    row_parts.append(np.column_stack([np.asarray(ids, np.float64)[local], rows]).reshape(-1, 6))
                else:
                    per_frame = self._run_uploaded(infos, payload)
                    for idx, boxes in zip(ids, per_frame):
                        all_results[idx] = boxes
                n_done += len(ids)
            usable = [None] * n_done
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
        if _as_rows:
            rows = np.concatenate(row_parts) if row_parts else np.zeros((0, 6), np.float64)
            return rows[np.argsort(rows[:, 0], kind="stable")]
        return all_results

    @abc.abstractmethod
    def _windows_per_frame(self, w: int, h: int) -> int:
        return

    @abc.abstractmethod
    def _upload_frame_batch(self, frames):
        return

    @abc.abstractmethod
    def _run_uploaded(self, infos, payload) -> List[List[LabeledBoundingBox]]:
        return
