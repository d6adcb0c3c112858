"""Binary labels of the detector (mirror of `data/db/label.py:12-13,74-92`)."""
IID_BACKGROUND = 0
IID_FOREGROUND = 1
KEY_BACKGROUND = "background"
KEY_FOREGROUND = "foreground"

_by_iid = {}


class Label:
    def __init__(self, iid: int, key: str):
        self._iid, self._key = iid, key

    @property
    def iid(self) -> int:
        return self._iid

    @property
    def key(self) -> str:
        return self._key

    @property
    def name(self) -> str:
        return self._key.title()

    @property
    def is_background(self) -> bool:
        return self._iid == IID_BACKGROUND

    @property
    def is_foreground(self) -> bool:
        return self._iid == IID_FOREGROUND


def get_by_iid(label_iid: int) -> Label:
    """Same object for the same iid; unknown iids raise ValueError like the reference."""
    if label_iid not in _by_iid:
        if label_iid == IID_BACKGROUND:
            _by_iid[label_iid] = Label(IID_BACKGROUND, KEY_BACKGROUND)
        elif label_iid == IID_FOREGROUND:
            _by_iid[label_iid] = Label(IID_FOREGROUND, KEY_FOREGROUND)
        else:
            raise ValueError("Found unknown label_iid: {}".format(label_iid))
    return _by_iid[label_iid]


def get_by_key(label_key: str) -> Label:
    if label_key == KEY_BACKGROUND:
        return get_by_iid(IID_BACKGROUND)
    if label_key == KEY_FOREGROUND:
        return get_by_iid(IID_FOREGROUND)
    raise ValueError("Found unknown label key: {}".format(label_key))
