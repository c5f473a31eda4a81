#!/usr/bin/env python
"""Developer probe: a stream of pinned 4K RGB8 host frames through fixca_cuda_frames_multi (one process, frames sharded
by index over the GPUs of the box).  usage: python scripts/e2e_frames_multi.py [frames] [max_gpus]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
import numpy as np
import torch
import fixca

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
maxg = int(sys.argv[2]) if len(sys.argv) > 2 else fixca.device_count()
h, w = 2160, 3840
src_t = torch.randint(0, 256, (nf, h, w, 3), dtype=torch.uint8).pin_memory()
dst_t = torch.empty((nf, h, w, 3), dtype=torch.uint8).pin_memory()
frames = [src_t[k].numpy() for k in range(nf)]
outs = [dst_t[k].numpy() for k in range(nf)]
p = fixca.FixCaParams(interpolation=2, lens_x=w // 2, lens_y=h // 2, blue=1.0, red=-1.5, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
ref = None
n = 1
while n <= maxg:
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        fixca.correct_frames(frames, p, flags=fixca.PRECISION_FAST, devices=list(range(n)), outs=outs)
        best = min(best, time.perf_counter() - t0)
    if ref is None:
        ref = dst_t.clone()
    print("%d pinned 4K RGB8 frames, Cubic FAST, one call over %d GPU(s): best %.2f ms = %.0f MP/s (%.1f GB/s each way)  identical to 1 GPU: %s"
          % (nf, n, best * 1e3, nf * h * w / 1e6 / best, nf * h * w * 3 / best / 1e9, bool(torch.equal(ref, dst_t))), flush=True)
    n *= 2
