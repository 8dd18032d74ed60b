"""Turns the round-2 GPU artefacts under gpurun_out/ (scratch) into the committed summaries under profiles/:
  r02_parity_fullsize.md   <- r02_parity_fullsize.jsonl (tests/test_gpu_fullsize.py)
  r02_ncu_kernels.md/.json <- r02_*.ncu-rep (ncu --set full captures), via `ncu -i ... --page raw --csv`
  r02_launch_shares*.txt   <- r02_launches_window*.csv (ncu gpu__time_duration launch lists of ONE steady window)
"""
import collections
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
           'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
           'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']

# algorithmic bytes / flops per launch at the captured shapes (bench.kernel_rooflines), peak 6575 GB/s / 1705 TF/s
ALGO = {
    'warp_vec': ('hbm', 100.6e6, 'K1 flow warp of the (2h, 2w, C) feature with the x2 flow upsample fused'),
    'warp3': ('hbm', 126.5e6, 'K1 fused warp of feat + conf + feat_UP from one flow read'),
    'aligned_sample2': ('hbm', 101.1e6, 'K4 AlignedConv2d resampling, tiled kernel (ks = 2)'),
    'reconstruct4': ('hbm', 59.6e6, 'K7 conv_last output + bicubic x4 base + clamp, one thread per LR pixel'),
    'match_tc': ('tensor', 1.209e12, 'K2 matching GEMM + argmax, single pass (default mode)'),
    'gather_cells': ('hbm', 25.4e6, 'K3 aa1 gather, one thread per (cell, vector) (first gather launch of profile_kernels.py)'),
    'gather_blocks': ('hbm', 25.4e6, 'K3 aa1 gather (first gather_blocks launch of profile_kernels.py)'),
    'aligned_sample': ('hbm', 101.1e6, 'K4 AlignedConv2d resampling of the gathered 2x feature'),
    'reconstruct': ('hbm', 59.6e6, 'K7 conv_last output + bicubic x4 base + clamp'),
    'conv_chain': ('tensor', 60 * 5.374e9, 'K5 60-layer trunk as one persistent launch (opt-in kernel)'),
    'conv_tc': ('tensor', 5.374e9, 'K5 conv3x3 48->48 @270x480 +ReLU +residual, one launch (the roofline kernel of the bench line)'),
    'match': ('tensor', 1.209e12, 'K2 matching GEMM + argmax, single pass'),
}
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def ncu_raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return None
    h, u, v = rows[0], rows[1], rows[2]
    d = {'kernel': v[h.index('Kernel Name')]}
    for m in METRICS:
        if m in h:
            i = h.index(m)
            try:
                d[m] = (float(v[i].replace(',', '')), u[i])
            except ValueError:
                pass
    return d


def kernels():
    lines, js = ['# ncu --set full summaries, round 2 (B200, --clock-control none, one launch each, cold cache under the profiler)', '',
                 'Captured with `ncu --set full --clock-control none --import-source on -k regex:<kernel> --launch-skip N --launch-count 1 '
                 'python tools/profile_kernels.py` (tools/gpu_r02_*.sh); the `.ncu-rep` files stay in gpurun_out/ (scratch, 3 MB each). '
                 'Graph-timed (warm) durations of the same launches are in the bench line (`roofline_other`).', ''], {}
    for f in sorted(os.listdir(G)):
        if not (f.startswith('r02_') and f.endswith('.ncu-rep')):
            continue
        key = f[4:-8]
        d = ncu_raw(os.path.join(G, f))
        if d is None:
            continue
        bound, algo, what = ALGO.get(key, ('hbm', None, ''))
        t, tu = d['gpu__time_duration.sum']
        t_us = t / 1000.0 if tu in ('ns', 'nsecond') else (t * 1000.0 if tu in ('ms', 'msecond') else t)
        traffic = sum(d[m][0] * UNIT[d[m][1]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum') if m in d)
        lines += [f'## {key}: {what}', f'`{d["kernel"][:150]}`', '', '| metric | value | unit |', '|---|---:|---|']
        for m in METRICS:
            if m in d:
                lines.append(f'| {m} | {d[m][0]:.6g} | {d[m][1]} |')
        if algo:
            if bound == 'hbm':
                lines.append(f'| algorithmic bytes | {algo/1e6:.1f} | MB  -> {algo/t_us/1e3:.0f} GB/s = {algo/t_us/1e3/6575.1:.2f} of the measured copy peak (cold, under ncu) |')
            else:
                lines.append(f'| algorithmic flops | {algo/1e9:.1f} | GFLOP -> {algo/t_us/1e6:.0f} TFLOP/s = {algo/t_us/1e6/1705.3:.2f} of the measured bf16 peak (cold, under ncu) |')
        lines.append(f'| DRAM traffic (read + write) | {traffic/1e6:.1f} | MB |')
        lines.append('')
        js[key] = {'duration_us': t_us, 'traffic': traffic, 'kernel': d['kernel'][:120], **{m: d[m][0] for m in METRICS if m in d}}
    open(os.path.join(P, 'r02_ncu_kernels.md'), 'w').write('\n'.join(lines) + '\n')
    json.dump(js, open(os.path.join(P, 'r02_ncu_kernels.json'), 'w'), indent=1)
    if 'conv_tc' in js:      # the bench line's roofline.traffic (dominant kernel: conv_tc 3x3 48->48 @270x480 +ReLU +residual)
        json.dump({'traffic': js['conv_tc']['traffic'], 'duration_us_cold': js['conv_tc']['duration_us'],
                   'note': 'dram__bytes_read.sum + dram__bytes_write.sum of ONE conv_tc launch (tools/ncu_conv.py under ncu --set full, round 2)'},
                  open(os.path.join(P, 'r02_ncu_conv_lr.json'), 'w'))


def launches(name, out):
    p = os.path.join(G, name)
    if not os.path.isfile(p):
        return
    rows = list(csv.reader(open(p)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hdr]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi or not r[vi].replace(',', '').replace('.', '').isdigit():
            continue
        nm = r[ki].split('(')[0].split('<')[0].replace('void ', '').replace('rv::', '').replace('(anonymous namespace)::', '')
        if nm.startswith('at::') or 'elementwise' in nm or 'emset' in nm:
            nm = 'torch copy/fill/index kernels + memsets'
        us = float(r[vi].replace(',', '')) / (1000.0 if r[ui] in ('ns', 'nsecond') else 1.0)
        agg[nm][0] += 1
        agg[nm][1] += us
    tot = sum(v[1] for v in agg.values())
    lines = [f'# ncu launch list of ONE steady-state window, RefVSR_MFID 270x480 bf16, eager (tools/profile_window.py), {name}',
             '# --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised durations: compare SHARES',
             f'{"kernel":44s} {"n":>4s} {"total us":>10s} {"share":>7s} {"avg us":>8s}']
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'{k:44s} {v[0]:4d} {v[1]:10.1f} {100*v[1]/tot:6.1f}% {v[1]/v[0]:8.1f}')
    lines.append(f'total {tot/1000:.2f} ms over {sum(v[0] for v in agg.values())} launches')
    open(os.path.join(P, out), 'w').write('\n'.join(lines) + '\n')


def parity():
    p = os.path.join(G, 'r02_parity_fullsize.jsonl')
    if not os.path.isfile(p):
        return
    recs = {}
    for ln in open(p):
        r = json.loads(ln)
        recs[(r['case'], r['precision'])] = r          # last run wins
    lines = ['# Full-size parity at the BASELINE configurations (tests/test_gpu_fullsize.py on B200)', '',
             'Golden vectors: `mfid_270x480` = the UNMODIFIED reference on CPU (fp32), config_RefVSR_MFID, 30 blocks, T = 7, LR = Ref = 270x480 -> '
             '1080x1920 (BASELINE configs[2]); `small_mfid_270x480_ref2x` = the oracle port (pinned to the reference at fixture sizes), '
             'config_RefVSR_small_MFID, Ref 540x960 (configs[1]; the reference\'s 67 GB similarity matrix does not fit the build container). '
             'reset_branch overridden to 2 so that 3-4 calls cover first / steady / forced-reset / steady-after-reset windows. '
             'Generator: tests/golden/make_fullsize.py.  PSNR / errors are OUR output vs the reference output (both in [0, 1]).', '',
             '| case | path | window | PSNR vs ref (dB) | max abs | 99.9 % abs | rel L2 |', '|---|---|---|---:|---:|---:|---:|']
    for (case, prec), r in sorted(recs.items()):
        for w in r['windows']:
            lines.append(f'| {case} | {prec} ({r["match_mode"]} matching) | {w["window"]} {w["kind"]} | {w["psnr_db"]:.1f} | {w["max_abs"]:.2e} | {w["p999_abs"]:.2e} | {w["rel_l2"]:.2e} |')
    lines += ['', '| case | path | index_map flip rate vs ref | conf max abs | flow max abs (px) |', '|---|---|---:|---:|---:|']
    for (case, prec), r in sorted(recs.items()):
        if 'index_flip_rate' in r:
            lines.append(f'| {case} | {prec} | {100*r["index_flip_rate"]:.4f} % | {r["conf_max_abs"]:.2e} | {r["flow_max_abs_px"]:.2e} |')
    lines += ['', 'Reading: the fp32 path is within 1e-3 everywhere (max abs <= 7.5e-4 incl. the few pixels behind near-tie argmax flips; '
              '99.9 % of the pixels within 1.4e-5).  The benchmarked bf16 path sits at 77 dB with max abs < 1e-3; its 2 % index flips are '
              'near-ties of the bf16 feature extractor, not of the matching GEMM (profiles/r01_match_mode.log).  PSNR-vs-ground-truth '
              'deltas through the reference\'s own run.py: tests/test_dropin_runpy.py (-m gpu).', '']
    open(os.path.join(P, 'r02_parity_fullsize.md'), 'w').write('\n'.join(lines))


if __name__ == '__main__':
    kernels()
    launches('r02_launches_window.csv', 'r02_launch_shares_nochain.txt')
    launches('r02_launches_window_chain.csv', 'r02_launch_shares.txt')
    parity()
    print(sorted(f for f in os.listdir(P) if f.startswith('r02')))
