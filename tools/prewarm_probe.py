"""Why does a 20-step region (the driver's --steps 20 --warmup 5) print ~91.5 us/step when a 1000-step region prints ~88 us?
One process, the bench's own step (lifting.lift_mean_var, cached plan, three rotating input sets): idle the GPU for a second,
run P untimed steps + 5 warm-up steps, synchronize, time 20 steps between two events; P swept.  Prints the SM clock NVML
reports right before each region."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerfdet_b200 import lifting  # noqa: E402


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:
        return -1


def main():
    dev = torch.device('cuda', 0)
    proj, pts = bench.build_scene(1000, bench.NV_PER_GPU)
    proj_d, pts_d = proj.to(dev), pts.to(dev)
    sets = [bench.host_features(2000 + i, bench.NV_PER_GPU).to(dev) for i in range(3)]
    views = [s[:, :, :bench.FEAT_HW[0], :bench.FEAT_HW[1]] for s in sets]

    def steps(n, base=0):
        out = None
        for i in range(n):
            out = lifting.lift_mean_var(views[(base + i) % 3], pts_d, proj_d)
        return out

    steps(10)
    torch.cuda.synchronize()
    for idle_s, pre in [(1.0, 0), (1.0, 0), (1.0, 20), (1.0, 50), (1.0, 100), (1.0, 250), (1.0, 500), (1.0, 1000), (0.0, 0), (1.0, 0)]:
        time.sleep(idle_s)
        c0 = sm_clock()
        steps(pre)
        steps(5)
        torch.cuda.synchronize()
        c1 = sm_clock()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps(20)
        e1.record()
        torch.cuda.synchronize()
        print(f'idle {idle_s:.1f} s, {pre:5d} untimed steps + 5 warm-up: 20 steps at {e0.elapsed_time(e1) / 20 * 1e3:6.1f} us/step   '
              f'(SM clock before {c0} MHz, at the start of the region {c1} MHz)', flush=True)
    # the 1000-step region for comparison
    time.sleep(1.0)
    steps(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps(1000)
    e1.record()
    torch.cuda.synchronize()
    print(f'idle 1.0 s, 20 warm-up: 1000 steps at {e0.elapsed_time(e1):.2f} ms = {e0.elapsed_time(e1):6.1f} us/step / 1000', flush=True)
    # the same 1000 steps in slices of 20 (events only, no synchronize): where inside the second does the step time move?
    time.sleep(1.0)
    steps(5)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(51)]
    evs[0].record()
    for k in range(50):
        steps(20)
        evs[k + 1].record()
    torch.cuda.synchronize()
    print('slices of 20 steps, us/step: ' + ' '.join(f'{evs[k].elapsed_time(evs[k + 1]) / 20 * 1e3:.1f}' for k in range(50)), flush=True)


if __name__ == '__main__':
    main()
