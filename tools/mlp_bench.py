#!/usr/bin/env python
"""Times the shared MLP on the GPU: the tcgen05 kernel at its two precisions (fp32-grade
split operands, plain bf16) and the fp32 FFMA kernel, at the
render shape (2048 rays x 64 samples) and the density-volume shape (40 x 40 x 16 voxels).  CUDA events, 20 launches
after 5 warm-ups.  Prints one JSON line per (shape, precision)."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfdet_b200.nerf_mlp import VanillaNeRFRadianceField  # noqa: E402

FLOP_FULL, FLOP_DENSITY = 734474, 462090


def main():
    dev = 'cuda'
    torch.manual_seed(0)
    for prec in ('fp32', 'bf16', 'fp32_ffma'):
        field = VanillaNeRFRadianceField(4, 256, 3, 70, 1, 128, precision=prec).to(dev)
        for name, rays, spr, full in (('render 2048x64', 2048, 64, True), ('density 40x40x16', 25600, 1, False)):
            p = rays * spr
            x = torch.empty(p, 3, device=dev).uniform_(-3.2, 3.2)
            f = torch.randn(p, 70, device=dev)
            d = torch.randn(rays, 3, device=dev)
            def run():
                if full:
                    return field(x.view(rays, spr, 3), d, f.view(rays, spr, 70))
                return field.query_density(x, f, return_alpha=True)
            for _ in range(5):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            # the same launches replayed from a CUDA graph: no host launch overhead between kernels
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(10):
                        run()
            torch.cuda.current_stream().wait_stream(side)
            g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            us_graph = e0.elapsed_time(e1) / 50 * 1e3
            flop = p * (FLOP_FULL if full else FLOP_DENSITY)
            print(json.dumps({'shape': name, 'precision': prec, 'points': p, 'us_eager': round(us, 1),
                              'us': round(us_graph, 1), 'tflops': round(flop / us_graph / 1e6, 1),
                              'mpoints_per_s': round(p / us_graph, 1)}), flush=True)


if __name__ == '__main__':
    main()
