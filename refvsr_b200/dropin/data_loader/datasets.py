"""`data_loader.datasets` of the reference with a per-frame tensor cache in `Test_datasets.__getitem__`.

The reference assembles every evaluation window from scratch (data_loader/datasets.py:222-316): 4T PNG decodes, T
float64 copies of each stream concatenated along the channel axis, then one float64 -> float32 transpose per stream
(data_loader/utils.py:81-93) - ~2.5 s per item at LR 270x480 + 1080p ground truth, T = 7 (tools/loader_bench.py),
far below the rate of the network.  Consecutive windows share T - 1 frames, and everything the reference does to a
frame is element-wise and per frame, so this class converts each FILE once to the float32 (3, H, W) tensor the
reference would produce for it and stacks cached tensors.  Returned items are bit-identical
(tests/test_loader_dropin.py).  Anything but the plain evaluation read (crop_valid, is_use_T) falls through to the
reference implementation; Train_datasets is the reference's class unchanged."""
import collections
import importlib.util
import os
import threading

import numpy as np
import torch

from data_loader import REFERENCE_DIR

_spec = importlib.util.spec_from_file_location('_refvsr_reference_data_loader_datasets', os.path.join(REFERENCE_DIR, 'datasets.py'))
_ref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_ref)
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith('__')})

from data_loader.utils import get_base_name, get_folder_name, read_frame      # noqa: E402  (drop-in utils: cached decode)

CACHE_TENSORS = int(os.environ.get('REFVSR_TENSOR_CACHE', 3 * 16))
_tensors = collections.OrderedDict()
_lock = threading.Lock()
stats = {'hits': 0, 'misses': 0}
_LUT = np.clip(np.arange(256) / 255., 0, 1.0).astype(np.float32)     # float64 division, then the FloatTensor rounding


def _frame_tensor(path):
    """what get_patch(is_crop=False) yields for one frame: FloatTensor of the (3, H, W) transpose of read_frame(path)"""
    try:
        key = (path, os.path.getmtime(path))
    except OSError:
        key = (path, None)
    with _lock:
        hit = _tensors.get(key)
        if hit is not None:
            _tensors.move_to_end(key)
            stats['hits'] += 1
            return hit
    t = None
    try:
        from PIL import Image
        with Image.open(path) as img:
            if img.mode == 'RGB':
                # read_frame computes uint8 / 255. in float64, clips to [0,1] (a no-op) and the caller rounds to float32
                # (data_loader/utils.py:12-41,81-93): the result depends on the 8-bit value only -> 256-entry table,
                # bit-identical and without the two float64 passes over the frame
                u8 = np.asarray(img)
                t = torch.from_numpy(np.ascontiguousarray(np.transpose(_LUT[u8], (2, 0, 1))))
    except Exception:
        t = None
    if t is None:                      # 16-bit / palette / grey inputs: the reference's own conversion
        t = torch.FloatTensor(np.ascontiguousarray(np.transpose(read_frame(path), (2, 0, 1))))
    with _lock:
        stats['misses'] += 1
        _tensors[key] = t
        while len(_tensors) > CACHE_TENSORS:
            _tensors.popitem(last=False)
    return t


class Test_datasets(_ref.Test_datasets):
    def __getitem__(self, index):
        if self.is_use_T or (self.config.is_crop_valid is True and self.is_valid):
            return super().__getitem__(index)
        video_idx = self.idx_video[index]
        frame_offset = self.idx_frame_flat[index] - self.frame_half
        lr_uw, lr_w, lr_t = (self.LR_UW_file_path_list[video_idx], self.LR_REF_W_file_path_list[video_idx],
                             self.LR_REF_T_file_path_list[video_idx])
        hr_uw = self.HR_UW_file_path_list[video_idx]
        sampled = np.arange(frame_offset, frame_offset + self.frame_num + self.frame_itr_num - 1).clip(min=0, max=len(lr_uw) - 1)
        video_name = lr_uw[sampled[self.frame_half]].split(os.sep)[-2]
        if self.vid_name is not None and video_name not in self.vid_name:
            return {'is_continue': True, 'is_first': True, 'video_name': video_name}
        for s in sampled:                                                       # the reference's consistency checks
            assert get_folder_name(str(lr_uw[s])) == get_folder_name(str(lr_w[s])) == get_folder_name(str(lr_t[s])) \
                == get_folder_name(str(hr_uw[s]))
            assert get_base_name(lr_uw[s]) == get_base_name(lr_w[s]) == get_base_name(lr_t[s])
        LR_UW = torch.stack([_frame_tensor(str(lr_uw[s])) for s in sampled])
        LR_W = torch.stack([_frame_tensor(str(lr_w[s])) for s in sampled])
        HR_UW = torch.stack([_frame_tensor(str(hr_uw[s])) for s in sampled])
        is_first = True
        if len(self.idx_video) > 1 and self.idx_video[index] == self.idx_video[index - 1]:
            is_first = False
        return {'LR_UW': LR_UW, 'LR_REF_W': LR_W, 'LR_REF_T': LR_W,      # (sic) data_loader/utils.py:103: W returned twice
                'HR_UW': HR_UW, 'HR_REF_W': HR_UW, 'HR_REF_T': HR_UW,
                'is_first': is_first,
                'video_len': len(self.LR_UW_file_path_list),
                'frame_len': len(self.LR_UW_file_path_list[video_idx]),
                'video_idx': video_idx,
                'frame_idx': sampled[self.frame_half],
                'video_name': video_name,
                'frame_name': os.path.basename(lr_uw[sampled[self.frame_half]])}
