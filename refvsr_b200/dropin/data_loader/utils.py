"""`data_loader.utils` of the reference with a decoded-frame cache in front of `read_frame`.

Only the plain evaluation read - `read_frame(path)` with every augmentation argument at its default
(data_loader/datasets.py:262-286) - is cached; any other call goes straight to the reference's function.  Cached
arrays are read-only: the datasets only slice and `np.concatenate` them (data_loader/datasets.py:288-291), so the
returned tensors are bit-identical, and an unexpected in-place write raises instead of corrupting a later window."""
import collections
import importlib.util
import os
import threading

from data_loader import REFERENCE_DIR

_spec = importlib.util.spec_from_file_location('_refvsr_reference_data_loader_utils', os.path.join(REFERENCE_DIR, 'utils.py'))
_ref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_ref)
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith('__')})     # `from data_loader.utils import *`

#: decoded frames kept per process: 4 streams x (T = 13 frames of the longest config + slack)
CACHE_FRAMES = int(os.environ.get('REFVSR_FRAME_CACHE', 4 * 16))
_cache = collections.OrderedDict()
_lock = threading.Lock()
stats = {'hits': 0, 'misses': 0}


def read_frame(path, norm_val=None, rotate_val=None, flip_val=None, gauss=None, gamma=0, sat_factor=None):
    if not (norm_val is None and rotate_val is None and flip_val is None and gauss is None and gamma == 0 and sat_factor is None):
        return _ref.read_frame(path, norm_val, rotate_val, flip_val, gauss, gamma, sat_factor)
    try:
        key = (path, os.path.getmtime(path))
    except OSError:
        key = (path, None)
    with _lock:
        hit = _cache.get(key)
        if hit is not None:
            _cache.move_to_end(key)
            stats['hits'] += 1
            return hit
    frame = _ref.read_frame(path)
    frame.setflags(write=False)
    with _lock:
        stats['misses'] += 1
        _cache[key] = frame
        while len(_cache) > CACHE_FRAMES:
            _cache.popitem(last=False)
    return frame


def clear_frame_cache():
    with _lock:
        _cache.clear()
        stats['hits'] = stats['misses'] = 0
