#!/usr/bin/env python
"""bench.py - frames/sec of RefVSR 4x SR (LR 270x480 -> 1080x1920) through refvsr_b200 on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3                # our arm (CUDA path)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --steps K --warmup W          # reference algorithm on host cores

A "step" is one call of SRNet.forward on the next sliding window of a synthetic clip = one output frame.
Steps are consecutive windows of one stream, so the timed region contains the model's own periodic forward-
branch resets (reset_branch = 9, RefVSR.py:168-170) in their natural proportion.  Workload: BASELINE.json
configs[2] (RefVSR_MFID, T=7, Ref 270x480) - the configuration the north-star target is quoted on;
`--workload small_mfid` selects configs[1] (RefVSR_small_MFID, Ref 540x960).
Multi-GPU: each rank processes its own reset-aligned segments of the clip (they are independent: SURVEY
8e / appendix B), no data-path collective -> weak scaling.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_REAL_STDOUT = sys.stdout

WORKLOADS = {
    'mfid': dict(config='config_RefVSR_MFID', ref_scale=1, precision='bf16', baseline_cfg=2),
    'small_mfid': dict(config='config_RefVSR_small_MFID', ref_scale=2, precision='fp16', baseline_cfg=1),
}
H, W, T = 270, 480, 7


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tensor_burst=d['bf16_tflops'], tensor_sustained=d['bf16_tflops_sustained'],
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled while the timed region runs (NVML every 10 ms, else nvidia-smi every 200 ms)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_ev = index, [], threading.Event()

    def _nvml_loop(self):
        """NVML directly (nvidia_ml_py): one sample every 10 ms, so even a 0.2 s timed region gets ~20 samples."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'), (0x20, 'sw_thermal_slowdown'), (0x4, 'sw_power_cap'))
        while not self._stop_ev.is_set():
            r = int(get_reasons(h))
            flags = ['Active' if r & bit else 'Not Active' for bit, _ in bits]
            self.rows.append([str(self.index), str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), '0'] + flags)
            self._stop_ev.wait(0.01)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            pass                                             # no NVML binding: fall back to the nvidia-smi CLI below
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                for ln in out.strip().splitlines():
                    self.rows.append([c.strip() for c in ln.split(',')])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=5)
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx[0] if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_model(workload, precision, device, graphs=True, fuse=False):
    import torch
    from refvsr_b200 import SRNet, get_config
    from refvsr_b200.modules import seeded_test_weights
    wl = WORKLOADS[workload]
    cfg = get_config(wl['config'], device=device, b200_precision=precision or wl['precision'], b200_cuda_graphs=graphs,
                     b200_fuse_resblocks=fuse)
    net = SRNet(cfg).eval()
    seeded_test_weights(net, seed=1234)          # random-init weights of the named architecture (no checkpoints offline)
    return cfg, net.to(device)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, from the committed
    `ncu --set full` capture (profiles/r01_ncu_conv_lr.json); None when the summary is absent."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_ncu_conv_lr.json')) as f:
            return float(json.load(f)['traffic'])
    except Exception:
        return None


def host_threads():
    """threads for the CPU legs: all cores up to 16 - beyond that the oracle's many small ATen ops lose to
    oversubscription (measured on the 128-core GPU host: 128 threads were >10x slower than 8)."""
    return int(os.environ.get('REFVSR_CPU_THREADS', min(os.cpu_count() or 1, 16)))


def window_indices(k, n):
    return [min(max(k - T // 2 + j, 0), n - 1) for j in range(T)]      # data_loader/datasets.py:233-234


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def kernel_rooflines(net, peaks):
    """Isolated timings of the path's dominant kernels at the BASELINE shapes (CUDA events on the launching
    stream, warm-up, buffers rotated through > L2 worth of memory so every launch misses L2 like in the step)."""
    import torch
    from refvsr_b200 import packing
    from refvsr_b200.lib import ACT_RELU
    ops = net.Network.ops
    dt = net.Network.act_dtype
    C = net.Network.mid_channels
    dev = 'cuda'
    out = {}
    e = 2 if dt != torch.float32 else 4
    nrot = max(2, int(160e6 // (H * W * C * e)) + 1)          # rotate through ~160 MB > 126 MB L2

    def timeit(fn, iters=30, warm=5):
        # the `iters` launches are captured into one CUDA graph so the number is device time per launch, not the
        # Python / ctypes / tensor-map-encode launch rate (which is what a plain loop measures for ~25 us kernels)
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(iters):
                fn(i)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    # K5: conv3x3 C->C + ReLU + residual at LR resolution (the ResidualBlockNoBN half-step; 305 of ~700 launches)
    g = torch.Generator().manual_seed(0)
    w = (torch.rand((C, C, 3, 3), generator=g) - 0.5) * 0.1
    layer = packing.pack_conv('bench.rb', w, torch.zeros(C), [(C, C)], 1, 1, dt, dev, net.Network.prefer_tc)
    xs = [torch.randn((H, W, C), device=dev).to(dt) for _ in range(nrot)]
    ys = [torch.empty((H, W, C), device=dev, dtype=dt) for _ in range(nrot)]
    tconv = timeit(lambda i: ops.conv2d(layer, xs[i % nrot], None, ys[i % nrot], res=xs[(i + 1) % nrot], act_pre=ACT_RELU))
    flops = 2.0 * 9 * C * C * H * W
    out['conv3x3_lr'] = dict(bound='tensor', achieved=flops / tconv / 1e12, peak=peaks['tensor_burst'], unit='TFLOP/s',
                             frac=flops / tconv / 1e12 / peaks['tensor_burst'], seconds=tconv,
                             algorithmic_flops=flops, bytes_min=3.0 * C * H * W * e)
    # fused residual block (rv_resblock): conv3x3 -> ReLU -> conv3x3 -> + x in one launch
    if net.Network.fuse_resblocks and dt != torch.float32:
        rb = packing.pack_resblock('bench.rbf', w, torch.zeros(C), w, torch.zeros(C), C, dt, dev)
        trb = timeit(lambda i: ops.resblock(rb, xs[i % nrot], ys[i % nrot], ACT_RELU))
        out['resblock_lr'] = dict(bound='tensor', achieved=2 * flops / trb / 1e12, peak=peaks['tensor_burst'], unit='TFLOP/s',
                                  frac=2 * flops / trb / 1e12 / peaks['tensor_burst'], seconds=trb,
                                  algorithmic_flops=2 * flops, bytes_min=2.0 * C * H * W * e)
    # K1: flow warp of the (C, 2h, 2w) feature with the on-the-fly x2 flow upsample (largest of the 3 warps)
    flow = torch.randn((H, W, 2), device=dev)
    nrot2 = max(2, int(160e6 // (4 * H * W * C * e)) + 1)
    fs = [torch.randn((2 * H, 2 * W, C), device=dev).to(dt) for _ in range(nrot2)]
    fo = [torch.empty((2 * H, 2 * W, C), device=dev, dtype=dt) for _ in range(nrot2)]
    twarp = timeit(lambda i: ops.warp(fs[i % nrot2], flow, fo[i % nrot2], flow_up2=True))
    byt = 2.0 * 4 * H * W * C * e + 8.0 * H * W
    out['warp_up'] = dict(bound='hbm', achieved=byt / twarp / 1e9, peak=peaks['hbm'], unit='GB/s',
                          frac=byt / twarp / 1e9 / peaks['hbm'], seconds=twarp, algorithmic_bytes=byt)
    # K2: matching GEMM + argmax (one per frame with reuse)
    split = net.Network.match_mode == 'split'
    kpad = 448 if split else 192
    A = torch.randn((H * W, kpad), device=dev).half()
    R = (H // 2) * (W // 2)
    B = torch.randn((R, kpad), device=dev).half()
    conf = torch.empty((H * W,), device=dev)
    idx = torch.empty((H * W,), device=dev, dtype=torch.int32)
    tm = timeit(lambda i: ops.match_argmax(A, B, conf, idx, impl=1), iters=10, warm=2)
    fl = 2.0 * 144 * (3 if split else 1) * H * W * R
    out['match_argmax'] = dict(bound='tensor', achieved=fl / tm / 1e12, peak=peaks['tensor_burst'], unit='TFLOP/s',
                               frac=fl / tm / 1e12 / peaks['tensor_burst'], seconds=tm, algorithmic_flops=fl,
                               issued_k=kpad, mode='split-fp16 (3 passes)' if split else 'fp16 single pass')
    return out


def cpu_baseline_sample(workload, threads=None):
    """The oracle port (CPU restatement of the reference's algorithm, as written: all 2(T-1) flows, matching
    for every frame of the window) on this box's host cores, one steady-state window at HALF resolution
    (136x240 LR; ~20-30 s), scaled to the full-size metric by the pixel ratio (4x; optimistic for the CPU because
    the matching GEMM grows 16x)."""
    import torch
    from oracle.refvsr_oracle import OracleRefVSR
    from refvsr_b200 import SRNet, get_config
    from refvsr_b200.modules import seeded_test_weights
    from refvsr_b200.synth import make_clip
    cores = threads or host_threads()
    torch.set_num_threads(cores)
    wl = WORKLOADS[workload]
    cfg = get_config(wl['config'], device='cpu')
    net = SRNet(cfg).eval()
    seeded_test_weights(net, seed=1234)
    orc = OracleRefVSR(cfg, net.state_dict())
    orc.compute_all_flows = True
    h2, w2 = 136, 240
    lrs, refs = make_clip(T + 1, h2, w2, wl['ref_scale'], seed=1234)
    orc.forward(lrs[window_indices(T // 2, T + 1)].unsqueeze(0), refs[window_indices(T // 2, T + 1)].unsqueeze(0), True)
    t0 = time.perf_counter()
    orc.forward(lrs[window_indices(T // 2 + 1, T + 1)].unsqueeze(0), refs[window_indices(T // 2 + 1, T + 1)].unsqueeze(0), False)
    dt = time.perf_counter() - t0
    area = (H * W) / float(h2 * w2)
    return {'value': 1.0 / dt / area, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'sample': f'1 steady-state window (T={T}) of {wl["config"]} at {h2}x{w2} LR -> {4*h2}x{4*w2}, fp32, '
                      f'{dt:.1f} s on {cores} threads; scaled to 270x480 by the pixel ratio {area:.2f}'}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from refvsr_b200.dist import exchange_halo, plan_segments
    from refvsr_b200.lib import load_library
    from refvsr_b200.synth import make_clip_range
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = load_library()
    peaks = measured_peaks()
    cfg, net = make_model(args.workload, args.precision, dev, graphs=not args.no_graphs, fuse=args.fuse)
    wl = WORKLOADS[args.workload]
    K, Wm = args.steps, args.warmup
    # One clip of world * n_own frames; rank r owns (decodes) frames [r*n_own, (r+1)*n_own), n_own a multiple of
    # reset_branch so every rank starts on a segment boundary.  The T//2 input-halo frames on each
    # side come from the neighbouring ranks over NCCL send/recv - the only communication of the whole job.
    rb = cfg.reset_branch
    n_own = (Wm + K + rb - 1) // rb * rb
    n_total = world * n_own
    plan = plan_segments(n_total, rb, world)
    assert plan[rank] == (rank * n_own, (rank + 1) * n_own)
    own_l, own_r = make_clip_range(rank * n_own, n_own, H, W, wl['ref_scale'], seed=1234)
    halo_ms = 0.0
    if world > 1:
        dl, dr = own_l.to(dev), own_r.to(dev)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        lrs_d, first = exchange_halo(dl, plan, rank, T // 2)
        refs_d, _ = exchange_halo(dr, plan, rank, T // 2)
        torch.cuda.synchronize()
        halo_ms = (time.perf_counter() - t0) * 1e3
        lrs, refs = lrs_d.cpu(), refs_d.cpu()
    else:
        lrs, refs, first = own_l, own_r, 0
        lrs_d, refs_d = lrs.to(dev), refs.to(dev)
    lrs_p, refs_p = lrs.pin_memory(), refs.pin_memory()
    f0 = rank * n_own                     # absolute index of this rank's first output frame
    n_frames = n_total                    # clamp windows at the ends of the CLIP, not of the rank's range
    off = f0 - first                      # position of frame f0 inside the local (halo-extended) arrays

    def local_ids(k):                     # window of output frame f0+k, as indices into the local arrays
        return [i - first for i in window_indices(f0 + k, n_frames)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_resident(k0, k1):
        for k in range(k0, k1):
            idx = torch.tensor(local_ids(k), device=dev)
            out = net(lrs_d.index_select(0, idx).unsqueeze(0), refs_d.index_select(0, idx).unsqueeze(0), k == 0, False, False)
        return out

    # End-to-end serving loop through the public API (SRNet.forward), double-buffered like any throughput-oriented
    # caller: while window k computes, the frames of window k+1 are uploaded from the pinned host clip on a copy stream,
    # and frame k-1 is downloaded on a second copy stream.  Every step still moves its full 7-frame window host ->
    # device and its 1080p result device -> host inside the timed region; the caller reads frame k-1 at step k.
    NB_ = 2
    dev_l = [torch.empty((1, T, 3, H, W), device=dev) for _ in range(NB_)]
    dev_r = [torch.empty((1, T, 3) + tuple(refs.shape[2:]), device=dev) for _ in range(NB_)]
    res_dev = [torch.empty((1, 3, 4 * H, 4 * W), device=dev) for _ in range(NB_)]
    res_host = [torch.empty((1, 3, 4 * H, 4 * W)).pin_memory() for _ in range(NB_)]
    h2d_s, d2h_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run_e2e(k0, k1):
        main = torch.cuda.current_stream()
        ev_up = [torch.cuda.Event() for _ in range(NB_)]
        ev_done = [None] * NB_
        ev_down = [None] * NB_

        prof = {} if os.environ.get('REFVSR_E2E_PROFILE') else None

        def tick(name, t0):
            if prof is not None:
                prof.setdefault(name, []).append(time.perf_counter() - t0)
            return time.perf_counter()

        def upload(k):
            b = k % NB_
            t = time.perf_counter()
            if ev_done[b] is not None:
                ev_done[b].synchronize()                     # step k-2 no longer reads dev_l[b] / host slot b is free
            t = tick('wait_prev_done', t)
            with torch.cuda.stream(h2d_s):                   # the window's 7 + 7 frames go straight from the pinned
                for j, i in enumerate(local_ids(k)):         # host clip into the device window (no host-side stacking)
                    dev_l[b][0, j].copy_(lrs_p[i], non_blocking=True)
                    dev_r[b][0, j].copy_(refs_p[i], non_blocking=True)
                ev_up[b].record(h2d_s)
            tick('h2d_enqueue', t)

        upload(k0)
        for k in range(k0, k1):
            b = k % NB_
            t = time.perf_counter()
            main.wait_event(ev_up[b])
            out = net(dev_l[b], dev_r[b], k == 0, False, False)
            t = tick('net_call', t)
            if ev_down[b] is not None:
                main.wait_event(ev_down[b])                  # res_dev[b] of step k-2 has been downloaded
            res_dev[b].copy_(out['result'])                  # the engine reuses its output buffer on the next call
            ev_done[b] = torch.cuda.Event()
            ev_done[b].record(main)
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(ev_done[b])
                res_host[b].copy_(res_dev[b], non_blocking=True)
                ev_down[b] = torch.cuda.Event()
                ev_down[b].record(d2h_s)
            t = tick('result_copies', t)
            if k + 1 < k1:
                upload(k + 1)
            t = time.perf_counter()
            if k > k0:
                ev_down[(k - 1) % NB_].synchronize()         # the caller reads frame k-1
            tick('wait_frame', t)
        ev_down[(k1 - 1) % NB_].synchronize()
        main.wait_stream(d2h_s)
        main.wait_stream(h2d_s)
        if prof is not None and rank == 0:
            print('[e2e profile, ms/step] ' + ' '.join(f'{k}={1e3 * sum(v) / len(v):.2f}' for k, v in prof.items()), file=sys.stderr)

    def timed(fn):
        net.Network.reset_state()
        fn(0, Wm)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        n0 = net.Network.executed_kernels
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(Wm, Wm + K)
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), net.Network.executed_kernels - n0, clocks

    # one-time CUDA-graph capture (one graph per ring phase x window kind, ~16 of them) in an untimed pre-pass over
    # the same windows, so that the K timed steps measure the steady serving rate rather than capture cost
    net.Network.reset_state()
    run_resident(0, Wm + K)
    ms_res, launches, clocks = timed(run_resident)
    ms_e2e, _, clocks_e2e = timed(run_e2e)
    bad = {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    if bad & set(clocks['reasons']):                          # re-measure once (timing rules)
        ms_res, launches, clocks = timed(run_resident)

    line = None
    if rank == 0:
        fps = world * K / (ms_res * 1e-3)
        fps_e2e = world * K / (ms_e2e * 1e-3)
        roof = kernel_rooflines(net, peaks)
        h2d = 2 * T * 3 * H * W * 4 if wl['ref_scale'] == 1 else T * 3 * H * W * 4 * 5
        line = {
            'metric': 'frames/sec 4x SR (270x480 -> 1080p)', 'value': fps, 'unit': 'frames/s', 'n_gpus': world,
            'steps': K, 'warmup': Wm, 'ms_per_step': ms_res / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': net.Network.precision, 'data': 'synthetic',
            'config': {'workload': f'{wl["config"]} 4x inference, T={T}, LR 270x480 + Ref '
                                   f'{270*wl["ref_scale"]}x{480*wl["ref_scale"]} -> 1080x1920 (BASELINE.json configs[{wl["baseline_cfg"]}]), '
                                   'seeded random-init weights, steady stream incl. reset_branch=9 resets',
                       'parallelism': f'clip sharded into reset-aligned frame ranges x{world}, {T//2}-frame input halo via NCCL '
                                      f'send/recv ({halo_ms:.1f} ms once per clip), no collective inside the forward',
                       'l2': 'working set per step (~1.5 GB of activations) >> 126 MB L2; no explicit flush',
                       'match_mode': net.Network.match_mode},
            'e2e': {'value': fps_e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': 3 * 16 * H * W * 4, 'ms_per_step': ms_e2e / K,
                    'note': 'pinned host frames -> device window -> SRNet.forward -> pinned host frame; double-buffered copy streams, the caller reads frame k-1 while window k computes'},
            'gpu_launches': int(launches),
            'gpu_launches_note': 'kernels of librefvsr_b200.so executed in the timed region on rank 0 (eager launches + '
                                 'kernel nodes replayed by the per-window CUDA graphs)',
            'clocks': clocks, 'clocks_e2e': clocks_e2e,
            'roofline': dict(roof['conv3x3_lr'], kernel='conv_tc_kernel 3x3 C->C @270x480 (+ReLU+residual)',
                             peak_source=peaks['source'], traffic=ncu_traffic()),
            'roofline_other': {k: v for k, v in roof.items() if k != 'conv3x3_lr'},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line['cpu_baseline'] = cpu_baseline_sample(args.workload)
            except Exception as ex:  # the checker must never take the bench down
                line['cpu_baseline'] = {'error': repr(ex)}
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle.refvsr_oracle import OracleRefVSR
    from refvsr_b200 import SRNet, get_config
    from refvsr_b200.modules import seeded_test_weights
    from refvsr_b200.synth import make_clip
    cores = host_threads()
    torch.set_num_threads(cores)
    wl = WORKLOADS[args.workload]
    cfg = get_config(wl['config'], device='cpu')
    net = SRNet(cfg).eval()
    seeded_test_weights(net, seed=1234)
    orc = OracleRefVSR(cfg, net.state_dict())
    orc.compute_all_flows = True
    # bounded sample: full-size windows are minutes each on CPU, so a step is one steady-state window of the
    # same stream at 136x240 LR; the rate is scaled to 270x480 by the pixel ratio (optimistic for the CPU)
    h2, w2 = 136, 240
    budget_s = float(os.environ.get('REFVSR_REF_BUDGET_S', '120'))
    n_frames = args.warmup + args.steps + 1
    lrs, refs = make_clip(min(n_frames, 24), h2, w2, wl['ref_scale'], seed=1234)
    nf = lrs.shape[0]
    t_start = time.perf_counter()
    done, t_timed = 0, 0.0
    for k in range(min(args.warmup + args.steps, nf)):
        ids = window_indices(k, nf)
        t0 = time.perf_counter()
        orc.forward(lrs[ids].unsqueeze(0), refs[ids].unsqueeze(0), k == 0)
        dt = time.perf_counter() - t0
        if k >= min(args.warmup, 1):
            done += 1
            t_timed += dt
        if time.perf_counter() - t_start > budget_s and done >= 1:
            break
    area = (H * W) / float(h2 * w2)
    fps = done / t_timed / area
    sample = (f'{done} consecutive windows (T={T}) of {wl["config"]} at {h2}x{w2} LR on {cores} host threads, fp32, '
              f'oracle port of the reference algorithm as written; rate scaled to 270x480 by pixel ratio {area:.2f}')
    line = {'impl': 'reference', 'metric': 'frames/sec 4x SR (270x480 -> 1080p)', 'value': fps, 'unit': 'frames/s',
            'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': done, 'warmup': min(args.warmup, 1),
            'ms_per_step': 1e3 * t_timed / done * area, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': f'{wl["config"]} 4x inference, T={T} (bounded CPU sample, see cpu_baseline.sample)'},
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def main():
    # libraries (NCCL's version banner, ...) may write to the C-level stdout; the contract is ONE JSON line on stdout,
    # so fd 1 is pointed at stderr and the JSON goes to a private copy of the original stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=18)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='mfid', choices=list(WORKLOADS))
    ap.add_argument('--precision', default=None, choices=[None, 'fp32', 'fp16', 'bf16'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graphs', action='store_true', help='eager kernel launches (for ncu launch lists)')
    ap.add_argument('--fuse', action='store_true', help='use the fused residual-block kernel (rv_resblock)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
