"""Image access for the hot path (the slice of `data/image_info.py:197-262` that inference uses).

``ImageInfo`` objects only need ``raw_original(cache)``, ``path_original``, ``full_key``,
``img_height_original`` / ``img_width_original`` and ``clear_raw_img_cache()``; the reference's
own ``ImageInfo`` satisfies that, and so do the two classes below.  The per-scale cache of
``raw_scaled`` is gone: pyramid levels are built on the GPU from the original frame.
"""
import os

import numpy as np


# data pointers of the frames that ``ArrayImageInfo(..., pin=True)`` placed in page-locked memory (a set lookup is
# cheaper than asking the driver about every frame of every batch)
PINNED_FRAME_PTRS = set()


class ArrayImageInfo:
    """An in-memory ``[H,W,3]`` uint8 frame (synthetic frames, video streams)."""

    def __init__(self, array: np.ndarray, key: str, dataset_key: str = "memory", pin: bool = False):
        """``pin=True`` keeps the frame in page-locked host memory (``cudaHostAlloc`` through torch), from where the
        engine uploads it with one asynchronous copy and no staging pass."""
        a = np.asarray(array)
        if a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.uint8:
            raise ValueError("expected an [H,W,3] uint8 array")
        if pin:
            import torch
            self._pinned = torch.empty(a.shape, dtype=torch.uint8).pin_memory()
            self._array = self._pinned.numpy()
            self._array[...] = a
            PINNED_FRAME_PTRS.add(self._pinned.data_ptr())
        else:
            self._array = np.ascontiguousarray(a)
        self._key = key
        self._dataset_key = dataset_key

    def __del__(self):
        p = getattr(self, "_pinned", None)
        if p is not None:
            PINNED_FRAME_PTRS.discard(p.data_ptr())

    @property
    def path_original(self) -> str:
        return self._key

    path_resized = path_original

    @property
    def basename(self) -> str:
        return os.path.basename(self._key)

    @property
    def dataset_key(self) -> str:
        return self._dataset_key

    @property
    def full_key(self) -> str:
        return self._dataset_key + "/" + self._key

    def raw_original(self, cache=False) -> np.ndarray:
        return self._array

    @property
    def img_height_original(self) -> int:
        return int(self._array.shape[0])

    @property
    def img_width_original(self) -> int:
        return int(self._array.shape[1])

    def clear_raw_img_cache(self):
        pass


class FileImageInfo(ArrayImageInfo):
    """A file on disk, decoded ONCE to RGB uint8 (the reference re-decodes it for every
    pyramid scale, `data/image_info.py:228-240`)."""

    def __init__(self, path_original: str, dataset_key: str = "files"):
        self._path = path_original
        self._key = path_original
        self._dataset_key = dataset_key
        self._array = None
        self._wh = None

    def raw_original(self, cache=False) -> np.ndarray:
        if self._array is None:
            from PIL import Image
            arr = np.asarray(Image.open(self._path).convert("RGB"), dtype=np.uint8)
            if not cache:
                return arr
            self._array = np.ascontiguousarray(arr)
        return self._array

    def _size(self):
        """(width, height) from the file header only: planning the launch chunks must not cost a decode."""
        if self._array is not None:
            return int(self._array.shape[1]), int(self._array.shape[0])
        if getattr(self, "_wh", None) is None:
            from PIL import Image
            with Image.open(self._path) as im:
                self._wh = (int(im.size[0]), int(im.size[1]))
        return self._wh

    @property
    def img_height_original(self) -> int:
        return self._size()[1]

    @property
    def img_width_original(self) -> int:
        return self._size()[0]

    def clear_raw_img_cache(self):
        self._array = None
