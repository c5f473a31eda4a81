#!/usr/bin/env python
"""Developer probe: is a stream of small launches bound by the host?  Times the host side of N back-to-back
fixca_cuda_region_dev calls (no synchronisation) beside the GPU time of the same N launches."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
import torch  # noqa: E402

import fixca  # noqa: E402

KW = dict(blue=1.0, red=-1.5, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
for name, h, w, ch, bpc, interp in (("4K rgb8 cubic", 2160, 3840, 3, 1, 2), ("1080p rgba8 cubic", 1080, 1920, 4, 1, 2),
                                    ("768-row rgb f32 band", 768, 8192, 3, -4, 2)):
    bpp = ch * abs(bpc)
    pitch = (w * bpp + 127) // 128 * 128
    nsets = 8
    srcs = [torch.randint(0, 255, (h, pitch), dtype=torch.uint8, device="cuda") for _ in range(nsets)]
    dsts = [torch.empty_like(srcs[0]) for _ in range(nsets)]
    p = fixca.FixCaParams(interpolation=interp, lens_x=w // 2, lens_y=h // 2, **KW)
    st = torch.cuda.current_stream().cuda_stream
    n = 2000
    for i in range(50):
        fixca.fix_ca_region_dev(srcs[i % nsets].data_ptr(), pitch, 0, h, dsts[i % nsets].data_ptr(), pitch, 0, w, h, bpp, bpc, p, 0, h, fixca.PRECISION_FAST, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(n):
        fixca.fix_ca_region_dev(srcs[i % nsets].data_ptr(), pitch, 0, h, dsts[i % nsets].data_ptr(), pitch, 0, w, h, bpp, bpc, p, 0, h, fixca.PRECISION_FAST, st)
    b.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("%-24s host %.2f us per launch, GPU %.2f us per launch (%s)" % (name, t_host / n * 1e6, a.elapsed_time(b) / n * 1e3,
                                                                             "HOST-BOUND" if t_host / n * 1e3 > 0.9 * a.elapsed_time(b) / n else "gpu-bound"))
