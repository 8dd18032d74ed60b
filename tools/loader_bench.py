"""CPU measurement of the evaluation data path: items / s of the reference's Test_datasets on a
synthetic RealMCVSR-shaped clip (LR 270x480 x3 streams + 1080p ground truth, PNG), with the reference's own
`data_loader.utils` and with the drop-in decoded-frame cache (refvsr_b200/dropin/data_loader).  Needs /root/reference.
    python tools/loader_bench.py [frames=24] [T=7]"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_loader_dropin import CODE          # the same subprocess body as the parity test


def make_clip(tmp, n):
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(1)
    base = rng.integers(0, 256, (1080 // 8, 1920 // 8, 3), dtype=np.uint8)
    for sub, stream, (h, w) in (('LR', 'UW', (270, 480)), ('LR', 'W', (270, 480)), ('LR', 'T', (270, 480)),
                                ('HR', 'UW', (1080, 1920)), ('HR', 'W', (1080, 1920)), ('HR', 'T', (1080, 1920))):
        d = os.path.join(tmp, sub, stream, '0001')
        os.makedirs(d, exist_ok=True)
        for k in range(n):
            img = Image.fromarray(np.roll(base, 3 * k, axis=1)).resize((w, h), Image.BICUBIC)     # compressible like a photo, not noise
            img.save(os.path.join(d, f'{k:04d}.png'))


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    code = CODE.replace('cfg.frame_num = 5', f'cfg.frame_num = {T}')
    with tempfile.TemporaryDirectory() as tmp:
        make_clip(tmp, n)
        for flag, name in (('0', 'reference data_loader.utils'), ('1', 'drop-in frame cache')):
            r = subprocess.run([sys.executable, '-c', code, ROOT, tmp, flag], capture_output=True, text=True, timeout=3600)
            assert r.returncode == 0, r.stderr[-2000:]
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(f'{name:32s}: {d["items"]} items in {d["seconds"]:.2f} s = {d["items"] / d["seconds"]:.2f} items/s'
                  f'  sha {d["sha"][:12]}  decode cache {d["stats"]}  tensor cache {d.get("tensor_stats")}')
