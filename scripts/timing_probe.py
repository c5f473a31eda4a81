#!/usr/bin/env python
"""Developer probe for the FIXCA_EXP_TIMING build (make EXP=timing EXPFLAGS=-DFIXCA_EXP_TIMING; FIXCA_LIB=.../libfixca_cuda_timing.so):
one launch of a batch (default) or of a single image; a few CTAs print where their compute warps' cycles went.
    python scripts/timing_probe.py [batch|single4k|single24mp]"""
import os
import sys

ROOT = os.environ.get("GRAFT_REPO_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
import torch  # noqa: E402

import fixca  # noqa: E402

KW = dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
which = sys.argv[1] if len(sys.argv) > 1 else "batch"
nf, h, w = {"batch": (32, 2160, 3840), "single4k": (1, 2160, 3840), "single24mp": (1, 4000, 6000)}[which]
bpp = 3
pitch = (w * bpp + 127) // 128 * 128
src = torch.randint(0, 255, (nf, h, pitch), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
p = fixca.FixCaParams(interpolation=2, lens_x=w // 2, lens_y=h // 2, **KW)
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if nf > 1:
        fixca.fix_ca_frames_dev(src.data_ptr(), pitch, pitch * h, dst.data_ptr(), pitch, pitch * h, nf, w, h, bpp, 1, p, fixca.PRECISION_FAST, st)
    else:
        fixca.fix_ca_region_dev(src.data_ptr(), pitch, 0, h, dst.data_ptr(), pitch, 0, w, h, bpp, 1, p, 0, h, fixca.PRECISION_FAST, st)
    b.record()
    torch.cuda.synchronize()
    print("launch %d: %.1f us" % (rep, a.elapsed_time(b) * 1e3), fixca.last_kernel(), flush=True)
