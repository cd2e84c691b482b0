#!/usr/bin/env python
"""Benchmark of NeRF-Det's lifting hot path on B200 (BASELINE.json metric: voxel-view samples / s at 1/2/4/8 GPUs and
achieved HBM GB/s against the measured peak).

  python bench.py [--gpus N] [--steps K] [--warmup W]                    the headline: fused lift, configs[1]
  python bench.py --scaling strong --views 100 ...                        configs[3]: ~100 views sharded over N GPUs
  python bench.py --workload render ...                                   configs[2]: render_rays (R2-R7 + the shared MLP)
  python bench.py --workload sweep ...                                    configs[4]: 8 scenes x 50 views over four voxel grids
  python bench.py --impl reference [same flags]                           the reference's own code on the host CPU

workload "lift" (default).  One step = one fused lift (project + nearest gather + masked mean + all-view variance +
count; reference nerfdet.py:164-181) of one synthetic ScanNet-shaped scene: nv views of [256, 60, 80] fp32 stride-4
features passed as the [:, :, :59, :80] slice, 40x40x16 voxels.  N = 1: nv = 50 (BASELINE.json configs[1]).  N > 1,
weak scaling (default): every rank lifts 50 views of one scene (50 N views in total), the per-rank accumulators
(S1, S2, count) meet in one exchange step over NVLink and every rank ends with mean / exp(-var) / count finalised
with the global view count.  N > 1, strong scaling: --views V views in total, sharded.

`value` counts device time only (inputs resident in HBM, three input sets rotated so that no step finds its features
in L2); `e2e` is the same metric through the Python API with pinned host features -> device -> lift -> results ->
pinned host inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NV_PER_GPU = 50
N_VOXELS = (40, 40, 16)
VOXEL_SIZE = (0.16, 0.16, 0.2)
CHANNELS = 256
FEAT_HW_PAD = (60, 80)
FEAT_HW = (59, 80)
OVERLAP_SMS = 28           # N > 1, pipelined: SMs that carry the exchange kernel while the others accumulate the next scene
N_INPUT_SETS = 3           # rotated so that no step finds its features in L2
HEAD_START_CYCLES = 400_000  # torch.cuda._sleep ahead of the timed region of the lift: ~200 us at 1965 MHz
SWEEP_GRIDS = [((40, 40, 16), (.16, .16, .2)), ((56, 56, 16), (.16, .16, .2)), ((64, 64, 24), (.1, .1, .13)),
               ((80, 80, 32), (.08, .08, .08))]
REF_COPY = os.path.join(ROOT, 'baseline', '_ref')     # the four reference files, copied by __graft_entry__.build()


def algorithmic_bytes(nv, c, hf, wf, n_vox, elt=4):
    """every sliced feature element read once, mean and cov written once, int64 count,
    projection matrices."""
    return nv * c * hf * wf * elt + 2 * c * n_vox * 4 + n_vox * 8 + nv * 48


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as fh:
            d = json.load(fh)
        return d, 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is busy.  The sampler runs from before the warm-up to the end
    of a sustained phase of LOAD_SECONDS of the very same steps that follows the timed region, so that it sees the GPU
    under this load even when the timed region itself lasts a few milliseconds."""
    Q = ('clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    LOAD_SECONDS = 1.2

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms',
                 os.environ.get('BENCH_SMI_MS', '200'), '-i', str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                clk, mx, util = float(parts[0]), float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax = mx
            if util > 0:
                sm.append(clk)
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith('active'):
                        reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax,
                'reasons': sorted(reasons), 'samples': len(self.lines), 'samples_under_load': len(sm),
                'note': 'sampled every 200 ms from the warm-up to the end of the sustained phase that follows the timed region'}


def bind_host_to_gpu(index):
    """Pins this process to the CPUs NVML names as local to GPU `index` BEFORE any pinned buffer is allocated, so that the
    end-to-end staging buffers land on the GPU's NUMA node (first touch).  Returns what was done, for the JSON line."""
    info = {'cpus': None, 'numa_node': None}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and w * 64 + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info['cpus'] = f'{allowed[0]}-{allowed[-1]} ({len(allowed)})'
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = f'/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node'
        if os.path.isfile(path):
            info['numa_node'] = int(open(path).read().strip())
    except Exception as exc:                                   # no NVML / no permission: leave the scheduler alone
        info['error'] = str(exc)[:80]
    return info


def build_scene(seed, nv, grid=N_VOXELS, vsize=VOXEL_SIZE):
    from nerfdet_b200 import lifting
    from nerfdet_b200.synthetic import SceneConfig, make_scene
    cfg = SceneConfig(n_views=nv, n_voxels=grid, voxel_size=vsize, channels=CHANNELS)
    sc = make_scene(cfg, seed=seed, with_images=False, with_features=False)
    proj = lifting.compute_projection(sc.img_meta, 4)
    pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin'])
    return proj, pts


def host_features(seed, nv, channels=CHANNELS):
    from nerfdet_b200.synthetic import make_features
    rs = np.random.RandomState(seed)
    return torch.from_numpy(make_features(rs, (nv, channels) + FEAT_HW_PAD))


def device_timed(fn, steps, barrier, warm=0):
    """`steps` calls of fn(i) between two CUDA events on the current stream, bracketed by barrier + synchronize.
    ``warm`` untimed iterations of the very same loop come first (the same tensors stay alive as in the timed loop, so the
    caching allocator has every block it will need before the region starts: a cudaMalloc inside a 2 ms region shows)."""
    out = None
    for i in range(warm):
        out = fn(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for i in range(steps):
        out = fn(i)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps, out


def per_step_times(fn, steps):
    """Event-separated step times (an event between two launches ends their overlap: these are single-step latencies)."""
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    torch.cuda.synchronize()
    evs[0].record()
    for i in range(steps):
        fn(i)
        evs[i + 1].record()
    torch.cuda.synchronize()
    t = np.array([evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(steps)])
    return {'min': float(t.min()), 'median': float(np.median(t)), 'p90': float(np.percentile(t, 90)), 'steps': steps,
            'note': 'one CUDA event between consecutive steps (no overlap of step i + 1 with the tail of step i)'}


# ------------------------------------------------------------------------------------------
# the reference's own code on the host CPU (and, for the same-box comparison, on the GPU with torch eager)
# ------------------------------------------------------------------------------------------
def load_reference():
    """The UNMODIFIED reference files (baseline/_ref, copied from /root/reference by __graft_entry__.build(); the copy
    is git-ignored and travels to the GPU box) behind oracle/ref_loader.py; None when they are not there."""
    root = REF_COPY if os.path.isfile(os.path.join(REF_COPY, 'mmdet3d/models/detectors/nerfdet.py')) else None
    if root is None and os.path.isdir('/root/reference/mmdet3d'):
        root = '/root/reference'
    if root is None:
        return None
    os.environ['NERFDET_REFERENCE_ROOT'] = root
    from oracle import ref_loader
    ref_loader.REFERENCE_ROOT = root
    return ref_loader.load()


def reference_lift(ns, features, points, projection, voxel_size):
    """reference backproject (nerfdet.py:393-420, called unmodified) + nerfdet.py:171-181 verbatim (those lines are
    inline in extract_feat, not a function)."""
    volume, valid = ns.nerfdet.backproject(features, points, projection, None, voxel_size)
    volume_sum = volume.sum(dim=0)
    valid = valid.sum(dim=0)
    volume_mean = volume_sum / (valid + 1e-8)
    volume_mean[:, valid[0] == 0] = .0
    volume_cov = torch.sum((volume - volume_mean.unsqueeze(0)) ** 2, dim=0) / (valid + 1e-8)
    volume_cov[:, valid[0] == 0] = 1e6
    volume_cov = torch.exp(-volume_cov)
    return volume_mean, volume_cov, valid


def cpu_lift_pass(ns, feats, pts, proj):
    if ns is not None:
        return reference_lift(ns, feats[:, :, :FEAT_HW[0], :FEAT_HW[1]], pts, proj, VOXEL_SIZE)
    from oracle import lift_oracle
    return lift_oracle.lift_mean_var(feats[:, :, :FEAT_HW[0], :FEAT_HW[1]], pts, proj)


def time_cpu_lift(steps, warmup, nv=NV_PER_GPU, grid=N_VOXELS, vsize=VOXEL_SIZE, channels=CHANNELS):
    """Returns (voxel-view samples / s scaled to the full channel count, s per pass, kind, description).  `channels` <
    CHANNELS runs a channel subset (the reference materialises [nv, C, N] fp32 plus two temporaries of that size: 31 GB
    at 80x80x32) and scales the time by CHANNELS / channels -- the work is proportional to the channel count."""
    torch.set_num_threads(os.cpu_count() or 1)
    ns = load_reference()
    proj, pts = build_scene(1000, nv, grid, vsize)
    feats = host_features(2000, nv, channels)
    with torch.no_grad():
        for _ in range(warmup):
            cpu_lift_pass(ns, feats, pts, proj)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_lift_pass(ns, feats, pts, proj)
        dt = (time.perf_counter() - t0) / steps * (CHANNELS / channels)
    n_vox = int(np.prod(grid))
    kind = 'reference' if ns is not None else 'port'
    what = ('the unmodified reference backproject + nerfdet.py:171-181 (baseline/_ref)' if ns is not None
            else 'oracle/lift_oracle.py (torch-CPU port of nerfdet.py:164-181; baseline/_ref not found)')
    return nv * n_vox / dt, dt, kind, what


def gpu_eager_reference(dev, feats_dev, pts_dev, proj_dev, steps=5):
    """The same reference code with CUDA tensors (torch eager; cuBLAS bmm, index_put, ...) on this very GPU: the honest
    same-box comparison.  Needs the 1.3 GB per-view volume the fused kernel never builds."""
    ns = load_reference()
    if ns is None:
        return None
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            f = feats_dev[:, :, :FEAT_HW[0], :FEAT_HW[1]]
            for _ in range(2):
                out = reference_lift(ns, f, pts_dev, proj_dev, VOXEL_SIZE)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = reference_lift(ns, f, pts_dev, proj_dev, VOXEL_SIZE)
            e1.record()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        return {'ms_per_step': ms, 'value': feats_dev.shape[0] * pts_dev[0].numel() / (ms * 1e-3), 'unit': 'samples/s',
                'steps': steps, 'what': 'unmodified reference backproject + nerfdet.py:171-181 with CUDA tensors (torch eager, TF32 off)'}, out
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


def run_reference(args, rank):
    """--impl reference: the reference's own implementation on the host cores, bounded samples of the same workload."""
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.workload == 'lift':
        nv = args.views if args.scaling == 'strong' else NV_PER_GPU
        steps_eff, warm_eff = min(steps, 20 if nv <= 50 else 8), min(warmup, 2)
        value, dt, kind, what = time_cpu_lift(steps_eff, warm_eff, nv=nv)
        sample = (f'{steps_eff} timed + {warm_eff} warm-up passes over ONE scene of {nv} views '
                  f'(C={CHANNELS}, 59x80, 40x40x16) with {what}')
        config = {'workload': 'nerfdet_res50_2x_low_res lift: backproject + mean/var/count (nerfdet.py:164-181)',
                  'views': nv, 'channels': CHANNELS, 'feature_hw': list(FEAT_HW), 'n_voxels': list(N_VOXELS),
                  'ran': f'one scene of {nv} views on the host CPU, whatever --gpus says (the reference lifts one scene per process)'}
        metric, unit = 'voxel_view_samples_per_sec', 'samples/s'
    elif args.workload == 'sweep':
        t_total, samples = 0.0, 0
        kind = what = None
        for grid, vs in SWEEP_GRIDS:
            v, dt, kind, what = time_cpu_lift(1, 1 if grid == SWEEP_GRIDS[0][0] else 0, grid=grid, vsize=vs, channels=32)
            t_total += dt
            samples += NV_PER_GPU * int(np.prod(grid))
        value, dt = samples / t_total, t_total
        sample = (f'one scene of {NV_PER_GPU} views per voxel grid of the sweep (4 lifts), once, 32 of the 256 channels with the '
                  f'time scaled by 8 (the reference materialises [nv, C, N] fp32 and two temporaries: 31 GB at 80x80x32), with {what}')
        config = {'workload': 'voxel-grid sweep of the lift (BASELINE.json configs[4])', 'grids': [list(g) for g, _ in SWEEP_GRIDS],
                  'views': NV_PER_GPU, 'ran': 'one scene per grid on the host CPU'}
        metric, unit = 'voxel_view_samples_per_sec', 'samples/s'
    else:
        value, dt, kind, what, config = time_cpu_render(min(steps, 3), min(warmup, 1))
        sample = f'{min(steps, 3)} timed passes of render_rays_func with {what}'
        metric, unit = 'rays_per_sec', 'rays/s'
    line = {
        'impl': 'reference', 'metric': metric, 'value': value, 'unit': unit,
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config,
        'cpu_baseline': {'value': value, 'unit': unit, 'cores': torch.get_num_threads(), 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# workload: lift
# ------------------------------------------------------------------------------------------
def lift_config(n_gpus, scaling, views_total, views_per_gpu, exchange, lanes, want_cov=True, overlap_sms=0, owner='every rank'):
    if n_gpus == 1:
        part = 'single GPU'
    elif exchange == 'multicast':
        part = (f'views sharded over {n_gpus} GPUs; (S1,S2,count) reduced in the NVSwitch and finalised by one kernel per rank '
                '(nd_lift_finalize_peers: multimem.ld_reduce of the channel slice, multimem.st of the rows)')
    elif exchange == 'peer':
        part = (f'views sharded over {n_gpus} GPUs; (S1,S2,count) reduced and finalised by one kernel per rank over '
                'NVLink peer memory (nd_lift_finalize_peers: P2P loads of the channel slice, P2P stores of the rows)')
    else:
        part = f'views sharded over {n_gpus} GPUs, 1 NCCL all-reduce of (S1,S2,count)'
    return {
        'workload': ('nerfdet_res50_2x_low_res lift: fused backproject + mean/var/count (nerfdet.py:164-181)'
                     if scaling == 'weak' else
                     'nerfdet_res101_2x_low_res_depth_sp test-time lift: ~100 views of one scene sharded over the GPUs'),
        'views_per_gpu': views_per_gpu, 'views_total': views_total, 'channels': CHANNELS,
        'feature_hw': list(FEAT_HW), 'feature_hw_padded': list(FEAT_HW_PAD), 'n_voxels': list(N_VOXELS),
        'feature_layout': 'NCHW fp32, non-contiguous [:, :, :59, :80] slice (reference layout)',
        'l2_policy': f'inputs exceed the 126 MB L2 and {N_INPUT_SETS} input sets are rotated',
        'geometry': 'the geometry plan (pixel offsets, counts, work distribution; depends on the cameras only) is built '
                    'once per scene geometry and reused by the steps (ops.cached_lift_plan); fresh_geometry times the '
                    'step with the plan rebuilt every call',
        'partitioning': part, 'scenes_in_flight': lanes if lanes else 1, 'want_cov': want_cov, 'exchange_sms': overlap_sms if lanes else 'all',
        'result_owner': owner,
    }


def bench_lift(args, rank, local_rank, world):
    import torch.distributed as dist
    from nerfdet_b200 import distributed as nd_dist
    from nerfdet_b200 import lifting, ops

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    all_cpus = os.sched_getaffinity(0)
    binding = bind_host_to_gpu(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = world
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    n_vox = int(np.prod(N_VOXELS))
    strong = args.scaling == 'strong'
    want_cov = not args.no_cov
    if strong:
        views_total = args.views
        v0, v1 = nd_dist.view_shard(views_total, rank, world)
        nv_local = v1 - v0
        proj_all, pts = build_scene(1000, views_total)
        proj = proj_all[v0:v1].contiguous()
    else:
        views_total, nv_local = NV_PER_GPU * n_gpus, NV_PER_GPU
        proj, pts = build_scene(1000 + rank, NV_PER_GPU)
    proj_d, pts_d = proj.to(dev), pts.to(dev)
    seeds = [2000 + 10 * rank + i for i in range(N_INPUT_SETS)]
    if strong:                                   # every rank holds ITS slice of the same scene's views
        host_sets = [host_features(2000 + i, views_total)[v0:v1].contiguous().pin_memory() for i in range(N_INPUT_SETS)]
    else:
        host_sets = [host_features(s, NV_PER_GPU).pin_memory() for s in seeds]
    dev_sets = [h.to(dev) for h in host_sets]

    # N > 1: one peer-mapped segment per scene in flight and one per end-to-end lane
    exchange = args.exchange if n_gpus > 1 else 'none'
    pipeline = n_gpus > 1 and exchange != 'nccl' and not args.no_pipeline
    overlap_sms = args.overlap_sms if pipeline else 0
    n_lanes = max(2, args.lanes) if pipeline else 1
    peers = None
    if exchange in ('auto', 'multicast'):
        err = ''
        try:
            peers = [nd_dist.PeerLift(CHANNELS, n_vox, dev, transport='multicast', want_cov=want_cov, overlap_sms=overlap_sms if i < n_lanes else 0)
                     for i in range(n_lanes + 2)]
        except Exception as e:                                 # no NVLS on this box / symmetric memory unavailable
            err = f'{type(e).__name__}: {e}'
        ok = torch.tensor([1 if peers is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)              # all ranks take the same path
        if int(ok.item()) == 1:
            exchange = 'multicast'
        else:
            if args.exchange == 'multicast':
                raise SystemExit(f'--exchange multicast is not available on this box ({err})')
            peers, exchange = None, 'peer'
    if exchange == 'peer':
        try:
            peers = [nd_dist.PeerLift(CHANNELS, n_vox, dev, want_cov=want_cov, overlap_sms=overlap_sms if i < n_lanes else 0)
                     for i in range(n_lanes + 2)]
        except RuntimeError as e:
            # PeerLift fails on ALL ranks together when CUDA IPC / peer access is not available between these GPUs; the
            # all-reduce form of the same exchange (NCCL + our finalise kernel) still runs
            if rank == 0:
                print(f'[bench] peer-memory exchange unavailable ({e}); using the NCCL all-reduce form', file=sys.stderr, flush=True)
            peers, exchange, pipeline, n_lanes = None, 'nccl', False, 1
    use_peer = peers is not None

    owner_rotate = bool(getattr(args, 'owner', 'all') == 'rotate') and n_gpus > 1

    def step(feats, lane=0, scene=-1):
        f = feats[:, :, :FEAT_HW[0], :FEAT_HW[1]]
        if n_gpus == 1:
            return lifting.lift_mean_var(f, pts_d, proj_d, want_cov=want_cov)
        if use_peer:
            # --owner rotate: scene i belongs to rank i % N (a data-parallel detector runs neck and heads of a scene on one
            # GPU); only the owner receives the finished volumes
            return peers[lane](f, pts_d, proj_d, views_total, owner=scene % n_gpus if owner_rotate and scene >= 0 else -1)
        return nd_dist.lift_mean_var_view_sharded(f, pts_d, proj_d, n_views_total=views_total, want_cov=want_cov)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    dev_lanes = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)] if pipeline else None

    def run_steps(count):
        """`count` steps; pipelined: scene i on lane i % n_lanes (its own stream and peer segment), all joined at the end."""
        out = None
        if not pipeline:
            for i in range(count):
                out = step(dev_sets[i % N_INPUT_SETS], 0, i)
            return out
        cur = torch.cuda.current_stream()
        for st in dev_lanes:
            st.wait_stream(cur)
        for i in range(count):
            with torch.cuda.stream(dev_lanes[i % n_lanes]):
                out = step(dev_sets[i % N_INPUT_SETS], i % n_lanes, i)
        for st in dev_lanes:
            cur.wait_stream(st)
        return out

    # ---- warm-up, then the device-resident timing: EXACTLY `steps` steps between two events ----
    run_steps(warmup)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # a ~200 us delay kernel ahead of the first event: the host enqueues the first steps while it spins, so the region
    # between the events is K steps of DEVICE time and not the host latency of the first launch on an idle GPU (30-45 us
    # of a 1.8 ms region at K = 20 -- and more on a slower or busier host; tools/prewarm_probe.py)
    torch.cuda._sleep(HEAD_START_CYCLES)
    ev0.record()
    out = run_steps(steps)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / steps
    value = views_total * n_vox / (ms_per_step * 1e-3)

    # ---- the same steps for ClockSampler.LOAD_SECONDS right behind the timed region: the 200 ms sampler cannot see a
    # region of K x 90 us, so the clocks and throttle reasons are taken under this identical load, and the sustained
    # step time (the GPU reaches its power cap within a few hundred milliseconds of this kernel) is reported beside
    # the timed one ----
    # (step counts derived from the all-reduced step time: the same on every rank, as the exchange steps require)
    run_steps(max(20, int(0.3 / (ms_per_step * 1e-3))))            # reach the steady state first
    n_sus = max(20, int(0.9 * ClockSampler.LOAD_SECONDS / (ms_per_step * 1e-3)))
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    run_steps(n_sus)
    s1.record()
    barrier()
    sus = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sus, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    sustained = {'ms_per_step': float(sus.item()) / n_sus, 'steps': n_sus,
                 'note': 'the same steps run for about a second right after the timed region (power-capped steady state); '
                         'clocks were sampled here'}

    extras = {}
    if n_gpus == 1:
        extras['step_times_us'] = per_step_times(lambda i: step(dev_sets[i % N_INPUT_SETS]), min(max(steps, 20), 200))
        f_views = [d[:, :, :FEAT_HW[0], :FEAT_HW[1]] for d in dev_sets]
        for i in range(3):                                  # (the first call of an entry point pays its one-off costs)
            ops.direct.lift_mean_var(f_views[i], pts_d, proj_d, None, True, 0)
        var_steps = max(50, min(steps, 200))                # labelled variants: not the contract's K, long enough to be stable
        fresh_ms, _ = device_timed(lambda i: ops.direct.lift_mean_var(f_views[i % N_INPUT_SETS], pts_d, proj_d, None, True, 0),
                                   var_steps, barrier, warm=5)
        extras['fresh_geometry'] = {'ms_per_step': fresh_ms, 'steps': var_steps, 'value': views_total * n_vox / (fresh_ms * 1e-3),
                                    'note': 'the one-shot entry nd_lift_mean_var: geometry plan (3 kernels) + lift, every step'}
        bf_sets = [d.to(torch.bfloat16) for d in dev_sets]
        bf_views = [d[:, :, :FEAT_HW[0], :FEAT_HW[1]] for d in bf_sets]
        for i in range(3):
            lifting.lift_mean_var(bf_views[i], pts_d, proj_d)
        bf_ms, _ = device_timed(lambda i: lifting.lift_mean_var(bf_views[i % N_INPUT_SETS], pts_d, proj_d), var_steps, barrier, warm=5)
        bfb = algorithmic_bytes(nv_local, CHANNELS, FEAT_HW[0], FEAT_HW[1], n_vox, 2)
        extras['variants'] = {'bf16_features': {'ms_per_step': bf_ms, 'steps': var_steps, 'value': views_total * n_vox / (bf_ms * 1e-3),
                                                'algorithmic_bytes_per_step': bfb,
                                                'note': 'same values rounded to bf16, fp32 accumulation; labelled variant, not the headline'}}
        del bf_sets, bf_views

    # ---- N > 1: parity of the exchanged result, outside the timed region ----
    parity = None
    if n_gpus > 1:
        parity = multi_gpu_parity(dev, rank, world, step, dev_sets[0], pts, proj, pts_d, views_total, dist, strong, proj_all if strong else None)

    # ---- end to end: pinned host features -> device, lift, results -> host, every step ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = min(steps, 30)
        lanes = []
        for _ in range(2):
            lanes.append({
                'stream': torch.cuda.Stream(device=dev),
                'stage': torch.empty_like(dev_sets[0]),
                'host_out': [torch.empty((CHANNELS, n_vox), dtype=torch.float32).pin_memory() for _ in range(2)],
                'host_cnt': torch.empty((n_vox,), dtype=torch.int64).pin_memory(),
            })

        def e2e_step(i):
            ln = lanes[i % 2]
            with torch.cuda.stream(ln['stream']):
                ln['stage'].copy_(host_sets[i % N_INPUT_SETS], non_blocking=True)
                mean, cov, cnt = step(ln['stage'], n_lanes + i % 2)
                ln['host_out'][0].copy_(mean.view(CHANNELS, -1), non_blocking=True)
                if cov is not None:
                    ln['host_out'][1].copy_(cov.view(CHANNELS, -1), non_blocking=True)
                ln['host_cnt'].copy_(cnt.view(-1), non_blocking=True)
                for t in (mean, cov, cnt):                  # allocated on this side stream: keep alive until it is done
                    if t is not None:
                        t.record_stream(ln['stream'])

        def e2e_join():
            for ln in lanes:
                torch.cuda.current_stream().wait_stream(ln['stream'])

        for ln in lanes:
            ln['stream'].wait_stream(torch.cuda.current_stream())
        for i in range(4):
            e2e_step(i)
        e2e_join()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for ln in lanes:
            ln['stream'].wait_event(e0)
        for i in range(e2e_steps):
            e2e_step(i)
        e2e_join()
        e1.record()
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e_ms = float(ems.item()) / e2e_steps
        h2d = torch.tensor([int(host_sets[0].numel() * 4)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(h2d)
        e2e = {'value': views_total * n_vox / (e2e_ms * 1e-3), 'unit': 'samples/s',
               'h2d_bytes_per_step': int(h2d.item()),
               'd2h_bytes_per_step': int((2 if want_cov else 1) * CHANNELS * n_vox * 4 + n_vox * 8) * n_gpus,
               'ms_per_step': e2e_ms, 'steps': e2e_steps,
               'note': 'pinned host features -> device, fused lift through the Python API, mean / cov / count -> pinned host, '
                       'every step; two streams so that consecutive steps overlap copy and compute',
               'host_binding': binding}
        if n_gpus == 1:
            # labelled variant, not the headline: the same loop with the features stored as bf16 on the host (half the
            # host-to-device bytes; results within the 1e-2 bar of the fp32 reference)
            host16 = [h.to(torch.bfloat16).pin_memory() for h in host_sets]
            stage16 = [torch.empty_like(dev_sets[0], dtype=torch.bfloat16) for _ in range(2)]

            def e2e16_step(i):
                ln = lanes[i % 2]
                with torch.cuda.stream(ln['stream']):
                    stage16[i % 2].copy_(host16[i % N_INPUT_SETS], non_blocking=True)
                    mean, cov, cnt = step(stage16[i % 2], n_lanes + i % 2)
                    ln['host_out'][0].copy_(mean.view(CHANNELS, -1), non_blocking=True)
                    if cov is not None:
                        ln['host_out'][1].copy_(cov.view(CHANNELS, -1), non_blocking=True)
                    ln['host_cnt'].copy_(cnt.view(-1), non_blocking=True)
                    for t in (mean, cov, cnt):
                        if t is not None:
                            t.record_stream(ln['stream'])
            for i in range(4):
                e2e16_step(i)
            e2e_join()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for ln in lanes:
                ln['stream'].wait_event(e0)
            for i in range(e2e_steps):
                e2e16_step(i)
            e2e_join()
            e1.record()
            torch.cuda.synchronize()
            ms16 = e0.elapsed_time(e1) / e2e_steps
            e2e['variants'] = {'bf16_features': {'value': views_total * n_vox / (ms16 * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms16,
                                                 'h2d_bytes_per_step': int(host16[0].numel() * 2),
                                                 'note': 'features kept as bf16 on the host: half the host-to-device bytes (1e-2 bar)'}}
            del host16, stage16

    if peers is not None:
        for p in peers:
            p.check()                                   # a peer that missed an exchange step invalidates the run

    # ---- CPU baseline and the reference on this GPU (rank 0, N = 1 only) ----
    cpu_baseline = gpu_eager = None
    os.sched_setaffinity(0, all_cpus)                  # the pinned buffers exist: give the CPU arm every core back
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        v, dt, kind, what = time_cpu_lift(4, 1)
        cpu_baseline = {'value': v, 'unit': 'samples/s', 'cores': torch.get_num_threads(), 'kind': kind,
                        'sample': f'4 timed + 1 warm-up passes of the full workload (nv=50, C=256, 59x80, 40x40x16) with {what}',
                        'ms_per_step': dt * 1e3}
        try:
            res = gpu_eager_reference(dev, dev_sets[0], pts_d, proj_d)
            if res is not None:
                gpu_eager, ref_out = res
                mean, _, cnt = step(dev_sets[0])
                torch.cuda.synchronize()
                gpu_eager['count_equal_to_ours'] = bool(torch.equal(ref_out[2].view(-1), cnt.view(-1)))
                gpu_eager['max_abs_mean_diff'] = float((ref_out[0] - mean).abs().max())
                gpu_eager['speedup_ours'] = gpu_eager['ms_per_step'] / ms_per_step
                del ref_out
        except torch.cuda.OutOfMemoryError:
            gpu_eager = {'unavailable': 'out of memory for the 1.3 GB per-view volume'}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        peak = float(peaks['hbm_gbs'])
        bytes_per_step = algorithmic_bytes(nv_local, CHANNELS, FEAT_HW[0], FEAT_HW[1], n_vox) - (0 if want_cov else CHANNELS * n_vox * 4)
        achieved = bytes_per_step / (ms_per_step * 1e-3) / 1e9
        launches = 1 + (0 if n_gpus == 1 else 2 if use_peer else 1)        # cached plan: the lift kernel (+ exchange)
        line = {
            'metric': 'voxel_view_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': n_gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': lift_config(n_gpus, args.scaling, views_total, nv_local, exchange, n_lanes if pipeline else 0, want_cov, overlap_sms,
                                  'the rank the scene belongs to (scene i -> rank i % N)' if owner_rotate and use_peer else 'every rank'),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ncu_traffic(), 'peak_source': peak_src,
                         'algorithmic_bytes_per_step': bytes_per_step,
                         'kernel': 'k_lift_quads (the only kernel of a step with a cached geometry plan)',
                         'note': 'per GPU; CUDA events around the K steps on the launching stream; consecutive steps overlap '
                                 '(programmatic dependent launch), step_times_us has the event-separated single-step time; a '
                                 '~200 us delay kernel runs ahead of the first event so that the region holds device time only '
                                 '(not the host latency of the first launch after the synchronize)'},
            'cpu_baseline': cpu_baseline, 'gpu_eager_baseline': gpu_eager, 'e2e': e2e,
            'gpu_launches': launches * steps, 'clocks': clocks,
        }
        line['sustained'] = sustained
        line.update(extras)
        if parity is not None:
            line['parity_ok'] = parity['ok']
            line['parity'] = parity
        print(json.dumps(line), flush=True)
    if peers is not None:
        for p in peers:
            p.close()
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_parity(dev, rank, world, step, feats_local, pts, proj_local, pts_d, views_total, dist, strong, proj_all):
    """Rank 0 compares the exchanged result of one step with a single-GPU CUDA lift over ALL ranks' views on 16 channels
    (the kernel the GPU tests pin on the C oracle; the exchange itself meets the oracle in tests/test_peer_gpu.py); every
    rank's result must be bit-identical to rank 0's.  Tolerance 1e-4 relative + 1e-5 max|ref| (fp32), counts exact."""
    n_chk = 16
    mean, cov, cnt = step(feats_local)
    torch.cuda.synchronize()
    has_cov = cov is not None
    mean, cnt = mean.reshape(CHANNELS, -1)[:n_chk].clone(), cnt.reshape(-1).clone()
    cov = cov.reshape(CHANNELS, -1)[:n_chk].clone() if has_cov else torch.zeros_like(mean)
    f_chk = feats_local[:, :n_chk].contiguous()
    nv_max = torch.tensor([f_chk.shape[0]], device=dev)
    dist.all_reduce(nv_max, op=dist.ReduceOp.MAX)
    nvm = int(nv_max.item())
    pad = torch.zeros((nvm,) + tuple(f_chk.shape[1:]), device=dev)
    pad[:f_chk.shape[0]] = f_chk
    nvs = torch.tensor([f_chk.shape[0]], device=dev)
    all_nv = [torch.zeros_like(nvs) for _ in range(world)]
    dist.all_gather(all_nv, nvs)
    all_f = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(all_f, pad)
    pj = torch.zeros((nvm, 3, 4), device=dev)
    pj[:proj_local.shape[0]] = proj_local.to(dev)
    all_p = [torch.zeros_like(pj) for _ in range(world)]
    dist.all_gather(all_p, pj)
    # every rank holds the same bits?
    same = torch.tensor([1], device=dev)
    ref_m, ref_c, ref_n = mean.clone(), cov.clone(), cnt.clone()
    dist.broadcast(ref_m, 0)
    dist.broadcast(ref_c, 0)
    dist.broadcast(ref_n, 0)
    if not (torch.equal(ref_m, mean) and torch.equal(ref_c, cov) and torch.equal(ref_n, cnt)):
        same[0] = 0
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    res = None
    if rank == 0:
        from nerfdet_b200 import ops
        feats = torch.cat([all_f[g][:int(all_nv[g].item())] for g in range(world)])
        projs = torch.cat([all_p[g][:int(all_nv[g].item())] for g in range(world)])
        assert feats.shape[0] == views_total
        fs = feats[:, :, :FEAT_HW[0], :FEAT_HW[1]]

        def bad(a, b):
            a, b = a.double().cpu().numpy(), np.asarray(b, dtype=np.float64)
            tol = 1e-4 * np.abs(b) + 1e-5 * np.abs(b).max()
            return int((np.abs(a - b) > tol).sum()), float(np.abs(a - b).max())
        m1, c1, n1 = ops.lift_mean_var(fs, pts_d, projs, None, True, 0)            # single GPU, all views
        torch.cuda.synchronize()
        bm, em = bad(mean, m1.cpu().numpy())
        bc, ec = bad(cov, c1.cpu().numpy()) if has_cov else (0, 0.0)
        cnt_ok = bool(torch.equal(cnt, n1))
        res = {'ok': bool(bm == 0 and bc == 0 and cnt_ok and int(same.item()) == 1),
               'vs_single_gpu_lift_16_channels': {'mean_outside_tol': bm, 'cov_outside_tol': bc, 'max_abs_err': [em, ec]},
               'note': 'the single-GPU lift is the one the GPU tests pin on the C oracle; the exchange itself is checked against '
                       'the oracle in tests/test_peer_gpu.py (in-process ranks)',
               'counts_equal': cnt_ok, 'all_ranks_bit_identical': bool(int(same.item()) == 1), 'cov_checked': has_cov,
               'tolerance': '|a-b| <= 1e-4 |ref| + 1e-5 max|ref|; counts exact', 'views_total': views_total}
    dist.barrier()
    return res


def ncu_traffic():
    """dram bytes per step from the committed ncu capture of the same command, if present."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.isfile(path):
        try:
            with open(path) as fh:
                return json.load(fh).get('dram_bytes_per_step')
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------
# workload: sweep (BASELINE.json configs[4]): 8 scenes x 50 views over four voxel grids
# ------------------------------------------------------------------------------------------
def bench_sweep(args, rank, local_rank, world):
    """Scene-sharded replicas: the 8 scenes of a batch are independent single-scene problems, so
    rank r lifts scenes r, r + N, ... with no collective; one step = the whole batch over all four grids."""
    import torch.distributed as dist
    from nerfdet_b200 import lifting
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_scenes = 8
    mine = list(range(rank, n_scenes, world))
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    dev_sets = [host_features(2000 + 10 * rank + i, NV_PER_GPU).to(dev) for i in range(N_INPUT_SETS)]
    geo = []
    for grid, vs in SWEEP_GRIDS:
        per_scene = []
        for s in mine:
            proj, pts = build_scene(1000 + s, NV_PER_GPU, grid, vs)
            per_scene.append((proj.to(dev), pts.to(dev)))
        geo.append(per_scene)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # caller-owned outputs, two sets per grid used alternately (at 80x80x32 mean + cov are 420 MB: allocating them per call
    # puts cudaMalloc inside the step)
    outs = []
    for grid, _ in SWEEP_GRIDS:
        n = int(np.prod(grid))
        outs.append([(torch.empty((CHANNELS, n), device=dev), torch.empty((CHANNELS, n), device=dev),
                      torch.empty((n,), dtype=torch.int64, device=dev)) for _ in range(2)])

    def one_grid(gi, i):
        out = None
        for k, (pr, pt) in enumerate(geo[gi]):
            out = lifting.lift_mean_var(dev_sets[(i + k) % N_INPUT_SETS][:, :, :FEAT_HW[0], :FEAT_HW[1]], pt, pr, out=outs[gi][k % 2])
        return out

    def step(i):
        for gi in range(len(SWEEP_GRIDS)):
            one_grid(gi, i)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(warmup):
        step(i)
    ms_per_step, _ = device_timed(step, steps, barrier)
    per_grid = []
    from nerfdet_b200 import ops as _ops
    for gi, (grid, vs) in enumerate(SWEEP_GRIDS):
        builds0 = _ops.plan_builds
        one_grid(gi, 0)
        g_ms, _ = device_timed(lambda i: one_grid(gi, i), max(5, min(steps, 20)), barrier)
        builds = _ops.plan_builds - builds0
        n_vox = int(np.prod(grid))
        byts = algorithmic_bytes(NV_PER_GPU, CHANNELS, FEAT_HW[0], FEAT_HW[1], n_vox) * len(mine)
        per_grid.append({'grid': list(grid), 'ms': g_ms, 'scenes_on_this_rank': len(mine),
                         'gbs': byts / (g_ms * 1e-3) / 1e9 if mine else 0.0, 'plans_built_during_timing': builds})
    t = torch.tensor([ms_per_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    samples = n_scenes * NV_PER_GPU * sum(int(np.prod(g)) for g, _ in SWEEP_GRIDS)
    value = samples / (ms_per_step * 1e-3)
    if rank == 0:
        peaks, peak_src = measured_peaks()
        peak = float(peaks['hbm_gbs'])
        for pg in per_grid:
            pg['frac_of_hbm_peak'] = pg['gbs'] / peak
        byts = sum(algorithmic_bytes(NV_PER_GPU, CHANNELS, FEAT_HW[0], FEAT_HW[1], int(np.prod(g))) for g, _ in SWEEP_GRIDS) * len(mine)
        achieved = byts / (ms_per_step * 1e-3) / 1e9
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            v, dt, kind, what = time_cpu_lift(1, 1)
            cpu_baseline = {'value': v, 'unit': 'samples/s', 'cores': torch.get_num_threads(), 'kind': kind,
                            'sample': f'one scene of 50 views at 40x40x16 (1 timed + 1 warm-up pass) with {what}', 'ms_per_step': dt * 1e3}
        line = {'metric': 'voxel_view_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': steps,
                'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': 'batch of 8 scenes x 50 views, fused lift over the voxel grids 40x40x16, 56x56x16, 64x64x24, '
                                       '80x80x32 (BASELINE.json configs[4])',
                           'partitioning': f'scene-sharded replicas: rank r lifts scenes r, r + {world}, ...; no collective',
                           'channels': CHANNELS, 'feature_hw': list(FEAT_HW), 'views_per_scene': NV_PER_GPU,
                           'l2_policy': f'{N_INPUT_SETS} input sets of 245 MB rotated',
                           'geometry': 'one cached geometry plan per (scene, grid), built during the warm-up (the headline bench times the '
                                       'fresh-geometry step); outputs are caller-owned buffers'},
                'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                             'traffic': None, 'peak_source': peak_src, 'per_grid_rank0': per_grid},
                'cpu_baseline': cpu_baseline, 'e2e': None, 'gpu_launches': 4 * len(mine) * steps, 'clocks': clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# workload: render (BASELINE.json configs[2]): render_rays train branch, N_rand = 2048, N_samples = 64, nt = 10
# ------------------------------------------------------------------------------------------
def render_scene(n_views=NV_PER_GPU, nt=10):
    from nerfdet_b200.synthetic import SceneConfig, make_mlp_state, make_scene
    cfg = SceneConfig(n_views=n_views, n_voxels=N_VOXELS, voxel_size=VOXEL_SIZE, channels=32, n_target_views=nt)
    sc = make_scene(cfg, seed=1000, with_images=True, with_features=True)
    return cfg, sc, make_mlp_state(191)


def time_cpu_render(steps, warmup, n_rays=256):
    """The reference render_rays_func (projection + grid_sample + statistics + MLP + compositing) on the host CPU."""
    torch.set_num_threads(os.cpu_count() or 1)
    ns = load_reference()
    cfg, sc, state = render_scene()
    rb = sc.ray_batch
    rs = np.random.RandomState(5)
    sel = rs.choice(rb['ray_o'].view(-1, 3).shape[0], n_rays, replace=False)
    ray_o, ray_d = rb['ray_o'].view(-1, 3)[sel].float(), rb['ray_d'].view(-1, 3)[sel].float()
    f2d = sc.features[:, :, :FEAT_HW[0], :FEAT_HW[1]].contiguous()
    imgs = sc.denorm_images[0]
    mlp_state = {k: v for k, v in state.items() if not k.startswith('mapping.')}
    if ns is not None:
        field = ns.nerf_mlp.VanillaNeRFRadianceField(net_depth=4, net_width=256, skip_layer=3, feature_dim=70,
                                                     net_depth_condition=1, net_width_condition=128)
        field.load_state_dict(mlp_state)
        proj = ns.projection.Projector()

        def once():
            return ns.render_ray.render_rays_func(ray_o, ray_d, None, None, f2d, imgs, cfg.aabb, cfg.near_far_range, 64,
                                                  n_rays, field, sc.img_meta, proj, 'image', 3, False, 0, True)
        kind, what = 'reference', 'the unmodified reference render_rays_func (baseline/_ref)'
    else:
        from oracle import mlp_oracle, render_oracle
        field = mlp_oracle.FieldOracle(mlp_state)

        def once():
            return render_oracle.render_image_mode(ray_o, ray_d, f2d, imgs, cfg.near_far_range, 64, field, sc.img_meta, det=True)
        kind, what = 'port', 'oracle/render_oracle.py (torch-CPU port; baseline/_ref not found)'
    with torch.no_grad():
        for _ in range(warmup):
            once()
        t0 = time.perf_counter()
        for _ in range(steps):
            once()
        dt = (time.perf_counter() - t0) / steps
    config = {'workload': 'render_rays_func: 64 samples per ray, 50 source views, deterministic sampling', 'rays_per_step': n_rays,
              'ran': f'{n_rays} of the 2048 rays of a batch per pass (bounded sample of the same workload)'}
    return n_rays / dt, dt, kind, what + f', {n_rays} rays per pass', config


def bench_render(args, rank, local_rank, world):
    """Rays are sharded over the ranks with no collective (the depth clamp bounds are batch-global and computed before
    sharding); one step = render_rays_func over this rank's share of the 2048 selected rays."""
    import torch.distributed as dist
    from nerfdet_b200 import render
    from nerfdet_b200.nerf_mlp import VanillaNeRFRadianceField
    from nerfdet_b200.projection import Projector
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    cfg, sc, state = render_scene()
    n_rand, n_samples = 2048, 64
    prec = args.mlp
    field = VanillaNeRFRadianceField(4, 256, 3, 70, 1, 128, precision=prec)
    field.load_state_dict({k: v for k, v in state.items() if not k.startswith('mapping.')})
    field = field.to(dev)
    imgs = sc.denorm_images[0].to(dev)
    f2d_host = sc.features[:, :, :FEAT_HW[0], :FEAT_HW[1]].contiguous()
    # the mapped 2-D features as the gather kernel wants them (channels-last, what live.map_features_2d produces)
    f2d = f2d_host.to(dev).contiguous(memory_format=torch.channels_last)
    rb = sc.ray_batch
    rs = np.random.RandomState(5)
    sels = [rs.choice(rb['ray_o'].view(-1, 3).shape[0], n_rand, replace=False) for _ in range(N_INPUT_SETS)]
    r0, r1 = rank * n_rand // world, (rank + 1) * n_rand // world
    rays = [(rb['ray_o'].view(-1, 3)[s][r0:r1].float().to(dev), rb['ray_d'].view(-1, 3)[s][r0:r1].float().to(dev)) for s in sels]
    proj = Projector()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        ro, rd = rays[i % N_INPUT_SETS]
        return render.render_rays_func(ro, rd, None, None, f2d, imgs, cfg.aabb, cfg.near_far_range, n_samples, n_rand, field,
                                       sc.img_meta, proj, 'image', 3, False, 0, True)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        while rank == 0 and world == 1 and time.perf_counter() - t0 < ClockSampler.LOAD_SECONDS:
            for i in range(20):
                step(i)
            torch.cuda.synchronize()
        ms_per_step, out = device_timed(step, steps, barrier)
        t = torch.tensor([ms_per_step], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())
        clocks = sampler.stop() if rank == 0 else None
        # the dominant kernel alone: the shared MLP over this rank's points
        ro, rd = rays[0]
        pts, z = render.sample_along_camera_ray(ro, rd, cfg.near_far_range, n_samples, det=True)
        glob = torch.randn(pts.shape[0], n_samples, 70, device=dev)
        mlp_ms, _ = device_timed(lambda i: field(pts, rd, glob), max(5, min(steps, 50)), barrier)
        # the other precision of the same kernel, as a labelled variant
        other = 'fp32' if prec == 'bf16' else 'bf16'
        field2 = VanillaNeRFRadianceField(4, 256, 3, 70, 1, 128, precision=other)
        field2.load_state_dict({k: v for k, v in state.items() if not k.startswith('mapping.')})
        field2 = field2.to(dev)

        def step2(i):
            ro2, rd2 = rays[i % N_INPUT_SETS]
            return render.render_rays_func(ro2, rd2, None, None, f2d, imgs, cfg.aabb, cfg.near_far_range, n_samples, n_rand, field2,
                                           sc.img_meta, proj, 'image', 3, False, 0, True)
        for i in range(3):
            step2(i)
        other_ms, _ = device_timed(step2, max(5, min(steps, 300)), barrier)
        other_mlp_ms, _ = device_timed(lambda i: field2(pts, rd, glob), max(5, min(steps, 50)), barrier)
        # end to end: host ray batch -> selection (host, like the reference) -> device -> render -> rgb / depth -> host
        e2e = None
        if not args.no_e2e and world == 1:
            host_rgb = torch.empty((n_rand, 3)).pin_memory()
            host_depth = torch.empty((n_rand,)).pin_memory()

            def e2e_step(i):
                ret = render.render_rays(rb, None, None, f2d, imgs, cfg.aabb, cfg.near_far_range, n_samples, n_rand, field,
                                         sc.img_meta, proj, 'image', 3, False, 0, False, True)
                host_rgb.copy_(ret['outputs_coarse']['rgb'], non_blocking=True)
                host_depth.copy_(ret['outputs_coarse']['depth'], non_blocking=True)
            for i in range(2):
                e2e_step(i)
            e_ms, _ = device_timed(e2e_step, min(steps, 10), barrier)
            e2e = {'value': n_rand / (e_ms * 1e-3), 'unit': 'rays/s', 'ms_per_step': e_ms,
                   'h2d_bytes_per_step': int(2 * n_rand * 3 * 4 + n_rand * 3 * 8 + n_rand * 8), 'd2h_bytes_per_step': n_rand * 16,
                   'note': 'render_rays train branch: host-side gt_depth > 0 filter and rng.choice over 660 000 rays (numpy, like the '
                           'reference), selected rays -> device, render, rgb / depth -> pinned host'}
    if rank == 0:
        peaks, peak_src = measured_peaks()
        n_pts = (r1 - r0) * n_samples
        flops = n_pts * 734474.0
        achieved = flops / (mlp_ms * 1e-3) / 1e12
        peak = float(peaks['bf16_tflops'])
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            v, dt, kind, what, _ = time_cpu_render(2, 1)
            cpu_baseline = {'value': v, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': kind, 'sample': f'2 timed + 1 warm-up passes with {what}',
                            'ms_per_step': dt * 1e3}
        line = {'metric': 'rays_per_sec', 'value': n_rand / (ms_per_step * 1e-3), 'unit': 'rays/s', 'n_gpus': world, 'steps': steps,
                'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'bf16 operands / f32 accumulate' if prec == 'bf16' else 'f32-grade (hi + lo bf16 operands / f32 accumulate)', 'data': 'synthetic',
                'config': {'workload': 'nerfdet_res50_2x_low_res_depth_sp render_rays_func: 2048 rays x 64 samples, 50 source views '
                                       '(R2-R7 + the shared MLP, BASELINE.json configs[2])',
                           'mlp_precision': prec, 'rays_per_step': n_rand, 'samples_per_ray': n_samples,
                           'partitioning': 'rays sharded contiguously over the ranks, no collective' if world > 1 else 'single GPU',
                           'point_samples_per_sec': n_rand * n_samples / (ms_per_step * 1e-3),
                           'avoided': 'the reference\'s [rays, samples, views, 35] tensor (917 MB) is never built'},
                'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                             'traffic': None, 'peak_source': peak_src + ' bf16 burst',
                             'kernel': 'the shared MLP on tcgen05 (k_nerf_mlp_tc), 734 474 algorithmic FLOP per point'
                                       + (' (the fp32-grade precision issues three bf16 products per multiply)' if prec == 'fp32' else ''),
                             'kernel_ms': mlp_ms,
                             'flops_per_launch': flops},
                'cpu_baseline': cpu_baseline, 'e2e': e2e, 'gpu_launches': 5 * steps, 'clocks': clocks,
                'variants': {f'mlp_{other}': {'ms_per_step': other_ms, 'value': n_rand / (other_ms * 1e-3), 'mlp_kernel_ms': other_mlp_ms,
                                              'note': ('fp32-grade: hi + lo bf16 operands, three tensor-core products per multiply (1e-4)'
                                                       if other == 'fp32' else 'plain bf16 operands (1e-2)')}}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='lift', choices=['lift', 'render', 'sweep'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='lift, N > 1: weak = 50 views per GPU (default), strong = --views in total, sharded')
    ap.add_argument('--views', type=int, default=100, help='--scaling strong: views of the scene (BASELINE.json configs[3]: ~100)')
    ap.add_argument('--mlp', default='bf16', choices=['fp32', 'bf16'], help='render: precision of the shared MLP kernel')
    ap.add_argument('--no-cov', action='store_true',
                    help='lift: mean and count only (the live path never reads the 256-channel volume_cov, SURVEY.md section 0.5); '
                         'a labelled variant of the workload: half the bytes are written, and half cross the links at N > 1')
    ap.add_argument('--owner', default='all', choices=['all', 'rotate'],
                    help='N > 1: which ranks receive the finished volumes of a scene: all of them (all-gather), or the rank the '
                         'scene belongs to, rotating with the scene index (a data-parallel detector; halves the exchange bytes)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--lanes', type=int, default=2, help='N > 1, pipelined: scenes in flight (streams / peer segments)')
    ap.add_argument('--overlap-sms', type=int, default=OVERLAP_SMS, help='N > 1, pipelined: SMs left to the exchange kernel')
    ap.add_argument('--no-pipeline', action='store_true',
                    help='N > 1: one scene at a time (default: two scenes in flight on two streams, the exchange of scene i '
                         'on OVERLAP_SMS SMs beside the accumulate of scene i + 1 on the others)')
    ap.add_argument('--exchange', default='peer', choices=['peer', 'multicast', 'auto', 'nccl'],
                    help='N > 1: how the per-rank accumulators meet: our kernel over per-peer P2P loads / stores (default), our '
                         'kernel over NVLS multicast, or NCCL all-reduce + finalise')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the lifting ops have no CPU fallback')
    if args.workload == 'lift':
        bench_lift(args, rank, local_rank, world)
    elif args.workload == 'sweep':
        bench_sweep(args, rank, local_rank, world)
    else:
        bench_render(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
