"""Frames/s of the reference's OWN evaluation driver (run.py -> eval.py -> evaluation/eval_qual_quan.py -> Trainer.evaluation) on a
synthetic RealMCVSR-style tree at the benchmark size (LR 270x480 -> 1080p), three ways:
  A  unmodified reference (eager PyTorch network, its own loop)
  B  refvsr_b200 network + data-loader drop-ins, the reference's loop        (REFVSR_DROPIN_EVAL=0)
  C  B + the evaluation-loop drop-in (device SSIM, async D2H, threaded writers, no gc / empty_cache)
Prints one JSON line; per-frame seconds are the driver's own "(...sec)" figures, wall = whole process incl. start-up."""
import json, os, re, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_dropin_runpy as T

ref = T._ref_root()
assert ref, 'stage the reference first: python oracle/build_ref.py'
frames = int(os.environ.get('EVAL_FRAMES', 12))
tmp = tempfile.mkdtemp(prefix='evalbench_')
T.make_tree(os.path.join(tmp, 'data'), 270, 480, videos=2, frames=frames)
T.make_ckpt(os.path.join(tmp, 'ckpt.pytorch'), 'config_RefVSR_MFID')
out = {}
for tag, dropin, env in (('A_reference', False, {}), ('B_network_dropin', True, {'REFVSR_DROPIN_EVAL': '0'}), ('C_network_and_eval_dropin', True, {})):
    t0 = time.time()
    rows, log = T.run_py(ref, tmp, tag, 'config_RefVSR_MFID', dropin=dropin, cpu=False, frame_num=7, extra_env=env, timeout=3000)
    wall = time.time() - t0
    secs = [float(x) for x in re.findall(r'\[EVAL [^\n]*\(([\d.]+)sec\)', log)]
    steady = sorted(secs[2:])[len(secs[2:]) // 2] if len(secs) > 2 else None
    out[tag] = dict(frames=len(rows), wall_s=round(wall, 1), median_frame_s=steady, fps_loop=(1.0 / steady if steady else None),
                    psnr_first=rows[0][2], ssim_first=rows[0][3])
    print(tag, out[tag], flush=True)
print(json.dumps(out))
