#!/usr/bin/env python
"""Benchmark of the lifting hot path (BASELINE.json metric: voxel-view samples / s and
achieved HBM GB/s against the measured peak).

One "step" = one pass of the fused lift (project + nearest gather + masked mean +
all-view variance + count; reference nerfdet.py:164-181) over one synthetic
ScanNet-shaped scene: nv=50 views of [256, 60, 80] fp32 stride-4 features passed as the
[:, :, :59, :80] slice, 40x40x16 voxels (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]        our CUDA path
  python bench.py --impl reference ...                        reference algorithm on the host CPU

N > 1 (launched by torch.distributed.run, one rank per GPU): views are sharded, every rank
lifts 50 views of the same scene, accumulators are combined with one NCCL all-reduce and
finalised with the global view count (weak scaling: 50*N views in total).

Prints ONE JSON line on rank 0.
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
OVERLAP_SMS = 20           # N > 1, pipelined: SMs that carry the exchange kernel while the others accumulate the next scene
N_INPUT_SETS = 3          # rotated so that no step finds its features in L2


def algorithmic_bytes(nv, c, hf, wf, n_vox, elt=4):
    """every sliced feature element read once, mean and cov written once,
    int64 count, projection matrices."""
    return nv * c * hf * wf * elt + 2 * c * n_vox * 4 + n_vox * 8 + nv * 48


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is busy."""
    Q = ('clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms',
                 os.environ.get('BENCH_SMI_MS', '200'),      # the recipe's period; every query perturbs the GPU for a few ms
                 '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
                'reasons': sorted(reasons), 'samples': len(self.lines), 'samples_under_load': len(sm)}


def build_scene(seed, nv):
    from nerfdet_b200 import lifting
    from nerfdet_b200.synthetic import SceneConfig, make_scene
    cfg = SceneConfig(n_views=nv, n_voxels=N_VOXELS, voxel_size=VOXEL_SIZE, channels=CHANNELS)
    sc = make_scene(cfg, seed=seed, with_images=False, with_features=False)
    proj = lifting.compute_projection(sc.img_meta, 4)
    pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin'])
    return proj, pts


def host_features(seed, nv):
    from nerfdet_b200.synthetic import make_features
    rs = np.random.RandomState(seed)
    return torch.from_numpy(make_features(rs, (nv, CHANNELS) + FEAT_HW_PAD))


# ------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port (torch-CPU restatement of nerfdet.py:164-181,
# materialising the per-view volume exactly like the reference) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_pass(feats, pts, proj):
    from oracle import lift_oracle
    return lift_oracle.lift_mean_var(feats[:, :, :FEAT_HW[0], :FEAT_HW[1]], pts, proj)


def time_cpu_reference(steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    proj, pts = build_scene(1000, NV_PER_GPU)
    feats = host_features(2000, NV_PER_GPU)
    for _ in range(warmup):
        cpu_reference_pass(feats, pts, proj)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pass(feats, pts, proj)
    dt = (time.perf_counter() - t0) / steps
    n_vox = int(np.prod(N_VOXELS))
    return NV_PER_GPU * n_vox / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    # bounded: at most ~60 full-scene passes (about 1 s each on 8 cores)
    steps_eff, warm_eff = min(steps, 40), min(warmup, 3)
    value, dt = time_cpu_reference(steps_eff, warm_eff)
    cores = torch.get_num_threads()
    sample = (f'{steps_eff} timed + {warm_eff} warm-up passes of the full workload '
              f'(nv={NV_PER_GPU}, C={CHANNELS}, 59x80, 40x40x16) with the torch-CPU port of '
              f'nerfdet.py:164-181 (oracle/lift_oracle.py)')
    line = {
        'impl': 'reference', 'metric': 'voxel_view_samples_per_sec', 'value': value, 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # the same `config` as our arm prints for these flags (the CPU port itself lifts one scene of views_per_gpu views)
        'config': workload_config(args.gpus, args.exchange if args.exchange != 'auto' else 'peer',
                                  max(2, args.lanes) if args.gpus > 1 and args.exchange != 'nccl' and not args.no_pipeline else 0),
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, exchange='peer', pipeline=False):
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
        'workload': 'nerfdet_res50_2x_low_res lift: fused backproject + mean/var/count (nerfdet.py:164-181)',
        'views_per_gpu': NV_PER_GPU, 'views_total': NV_PER_GPU * n_gpus, 'channels': CHANNELS,
        'feature_hw': list(FEAT_HW), 'feature_hw_padded': list(FEAT_HW_PAD), 'n_voxels': list(N_VOXELS),
        'feature_layout': 'NCHW fp32, non-contiguous [:, :, :59, :80] slice (reference layout)',
        'l2_policy': f'inputs (241.7 MB/step) exceed the 126 MB L2 and {N_INPUT_SETS} input sets are rotated',
        'partitioning': part,
        'scenes_in_flight': pipeline if pipeline else 1,
    }


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)   # 140 ms timed region: a clocks query landing in a 28 ms one costs 25 %
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--lanes', type=int, default=2, help='N > 1, pipelined: scenes in flight (streams / peer segments)')
    ap.add_argument('--no-pipeline', action='store_true',
                    help='N > 1: one scene at a time (default: two scenes in flight on two streams, the exchange of scene i '
                         'on OVERLAP_SMS SMs beside the accumulate of scene i + 1 on the others)')
    ap.add_argument('--exchange', default='peer', choices=['peer', 'multicast', 'auto', 'nccl'],
                    help='N > 1: how the per-rank accumulators meet: our kernel over per-peer P2P loads / stores (default: the '
                         'fastest at 2 and at 8 GPUs, profiles/r01_s6_multigpu_exchange.txt), our kernel over NVLS multicast '
                         '(in-switch reduction; auto = multicast if the box supports it, else peer), or NCCL all-reduce + finalise')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from nerfdet_b200 import distributed as nd_dist
    from nerfdet_b200 import lifting, ops

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the lifting ops have no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = world
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    n_vox = int(np.prod(N_VOXELS))

    # ---- inputs: this rank's 50 views, N_INPUT_SETS different feature sets resident in HBM ----
    proj, pts = build_scene(1000 + rank, NV_PER_GPU)
    proj_d, pts_d = proj.to(dev), pts.to(dev)
    host_sets = [host_features(2000 + 10 * rank + i, NV_PER_GPU).pin_memory() for i in range(N_INPUT_SETS)]
    dev_sets = [h.to(dev) for h in host_sets]
    views_total = NV_PER_GPU * n_gpus

    # N > 1: one peer-mapped segment for the device-resident loop and one per end-to-end lane (their results are views of it)
    exchange = args.exchange if n_gpus > 1 else 'none'
    pipeline = n_gpus > 1 and exchange != 'nccl' and not args.no_pipeline
    overlap_sms = OVERLAP_SMS if pipeline else 0
    n_lanes = max(2, args.lanes) if pipeline else 1
    peers = None
    if exchange in ('auto', 'multicast'):
        err = ''
        try:
            peers = [nd_dist.PeerLift(CHANNELS, n_vox, dev, transport='multicast', overlap_sms=overlap_sms if i < n_lanes else 0)
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
            if rank == 0:
                print(f'[bench] multicast transport unavailable ({err}); using per-peer P2P', file=sys.stderr, flush=True)
            peers = None
            exchange = 'peer'
    if exchange == 'peer':
        try:
            # the last two segments serve the end-to-end lanes: host-link bound, exchange on the whole GPU (4.71 vs 5.43 ms/step at 2 GPUs)
            peers = [nd_dist.PeerLift(CHANNELS, n_vox, dev, overlap_sms=overlap_sms if i < n_lanes else 0)
                     for i in range(n_lanes + 2)]
        except RuntimeError as e:
            # PeerLift fails on ALL ranks together (it reduces a success flag) when CUDA IPC / peer access is not available
            # between these GPUs; the all-reduce form of the same exchange (NCCL + our finalise kernel) still runs
            if rank == 0:
                print(f'[bench] peer-memory exchange unavailable ({e}); using the NCCL all-reduce form', file=sys.stderr, flush=True)
            peers, exchange, pipeline, n_lanes = None, 'nccl', False, 1
    use_peer = peers is not None

    def step(feats, lane=0):
        f = feats[:, :, :FEAT_HW[0], :FEAT_HW[1]]
        if n_gpus == 1:
            return lifting.lift_mean_var(f, pts_d, proj_d)
        if use_peer:
            return peers[lane](f, pts_d, proj_d, views_total)
        return nd_dist.lift_mean_var_view_sharded(f, pts_d, proj_d, n_views_total=views_total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- device-resident timing ----
    dev_lanes = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)] if pipeline else None

    def run_steps(count):
        """`count` steps; pipelined: scene i on lane i % 2 (its own stream and peer segment), all joined at the end."""
        out = None
        if not pipeline:
            for i in range(count):
                out = step(dev_sets[i % N_INPUT_SETS])
            return out
        cur = torch.cuda.current_stream()
        for st in dev_lanes:
            st.wait_stream(cur)
        for i in range(count):
            with torch.cuda.stream(dev_lanes[i % n_lanes]):
                out = step(dev_sets[i % N_INPUT_SETS], i % n_lanes)
        for st in dev_lanes:
            cur.wait_stream(st)
        return out

    out = run_steps(warmup)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    out = run_steps(steps)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / steps
    value = views_total * n_vox / (ms_per_step * 1e-3)

    # ---- end to end: pinned host features -> device, lift, results -> host, every step ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = min(steps, 30)
        # two streams, each with its own device staging buffer and pinned result buffers: step i + 1's host->device copy
        # runs while step i computes and copies its results back (the copies of EVERY step are inside the timed region;
        # this is how a streaming caller would drive the op)
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
                ln['host_out'][1].copy_(cov.view(CHANNELS, -1), non_blocking=True)
                ln['host_cnt'].copy_(cnt.view(-1), non_blocking=True)
                # the outputs were allocated on this side stream: keep them alive until it has used them
                for t in (mean, cov, cnt):
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
        e2e = {'value': views_total * n_vox / (e2e_ms * 1e-3), 'unit': 'samples/s',
               'h2d_bytes_per_step': int(host_sets[0].numel() * 4) * n_gpus,
               'd2h_bytes_per_step': int(2 * CHANNELS * n_vox * 4 + n_vox * 8) * n_gpus,
               'ms_per_step': e2e_ms, 'steps': e2e_steps,
               'note': 'pinned host features -> device, fused lift through the Python API, mean / cov / count -> pinned host, '
                       'every step; two streams so that consecutive steps overlap copy and compute'}

    clocks = sampler.stop() if rank == 0 else None
    if peers is not None:
        for p in peers:
            p.check()                                   # a peer that missed an exchange step invalidates the run

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        v, dt = time_cpu_reference(3, 1)
        cpu_baseline = {'value': v, 'unit': 'samples/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                        'sample': '3 timed + 1 warm-up passes of the full workload (nv=50, C=256, 59x80, '
                                  '40x40x16) with oracle/lift_oracle.py (torch-CPU port of nerfdet.py:164-181)',
                        'ms_per_step': dt * 1e3}

    if rank == 0:
        peak, peak_src = measured_peaks()
        bytes_per_step = algorithmic_bytes(NV_PER_GPU, CHANNELS, FEAT_HW[0], FEAT_HW[1], n_vox)
        achieved = bytes_per_step / (ms_per_step * 1e-3) / 1e9
        # launches per step: 1 pixel-index pre-pass + (stage, gather) per channel chunk (+ finalize when sharded)
        # (+ when sharded: finalise-over-peers + wait kernel, or our finalise kernel after NCCL's all-reduce)
        launches = ops.lift_launch_count(dev_sets[0][:, :, :FEAT_HW[0], :FEAT_HW[1]], n_vox) + \
            (0 if n_gpus == 1 else 2 if use_peer else 1)
        line = {
            'metric': 'voxel_view_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': n_gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(n_gpus, exchange, n_lanes if pipeline else 0),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ncu_traffic(), 'peak_source': peak_src,
                         'algorithmic_bytes_per_step': bytes_per_step,
                         'note': 'per-GPU; the fused lift (its phase launches) timed as one unit with CUDA events on '
                                 'the launching stream'},
            'cpu_baseline': cpu_baseline, 'e2e': e2e, 'gpu_launches': launches * steps, 'clocks': clocks,
        }
        print(json.dumps(line), flush=True)
    if peers is not None:
        for p in peers:
            p.close()
    if world > 1:
        dist.destroy_process_group()


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


if __name__ == '__main__':
    main()
