#!/usr/bin/env python
"""Developer probe: ONE host call over several GPUs of the box (fixca_cuda_region_multi, one process -- what the
plug-in would call on a multi-GPU host): each GPU moves its band over its own PCIe link.
usage: python scripts/e2e_multi.py [pinned|pageable] [max_gpus]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
import numpy as np
import torch
import fixca

mode = sys.argv[1] if len(sys.argv) > 1 else "pinned"
maxg = int(sys.argv[2]) if len(sys.argv) > 2 else fixca.device_count()
h, w = 8192, 12288
nbytes = h * w * 6
if mode == "pinned":
    src_t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst_t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    img, out = src_t.numpy().view(np.uint16).reshape(h, w, 3), dst_t.numpy().view(np.uint16).reshape(h, w, 3)
else:
    img, out = np.empty((h, w, 3), np.uint16), np.zeros((h, w, 3), np.uint16)
rng = np.random.default_rng(1)
img[:] = rng.integers(0, 65535, size=(64, w, 3), dtype=np.uint16).repeat(h // 64, axis=0)
kw = dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9, lens_x=w // 2, lens_y=h // 2, interpolation=2)
p = fixca.FixCaParams(**kw)
ref = None
n = 1
while n <= maxg:
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        fixca.correct(img, p, out=out, flags=fixca.PRECISION_FAST, devices=list(range(n)))
        best = min(best, time.perf_counter() - t0)
    if ref is None:
        ref = out.copy()
    same = bool((out == ref).all())
    print("%s caller, 100 MP RGB16 Cubic FAST, one call over %d GPU(s): best %.2f ms = %.0f MP/s  identical to 1 GPU: %s"
          % (mode, n, best * 1e3, h * w / 1e6 / best, same), flush=True)
    n *= 2
