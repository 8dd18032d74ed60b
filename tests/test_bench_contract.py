"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the unmodified reference, or the
oracle port when no checkout is staged, timed on host cores) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, REFVSR_CPU_THREADS='4', REFVSR_BENCH_LR='32x48')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 0 and d['value'] > 0
    assert d['metric'].startswith('frames/sec 4x SR')
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] == 4 and cb['value'] == d['value'] and 'sample' in cb
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_our_arm_fails_loudly_without_a_gpu():
    """no CPU fallback: on a box without CUDA the product arm must exit non-zero, not print a number"""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('needs a CUDA-less host')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0', '--no-cpu-baseline'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not any(ln.strip().startswith('{') for ln in r.stdout.splitlines())
