#!/usr/bin/env python
"""Run one workload a few times (for ncu): python scripts/profile_one.py <fast|exact|linear|none|rgb8|rgb8lin|rgba16|rgb8_4k> [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
import torch  # noqa: E402

import fixca  # noqa: E402

KW = dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
which = sys.argv[1] if len(sys.argv) > 1 else "fast"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = {
    "fast":   (8192, 12288, 3, 2, 2, 2, fixca.PRECISION_FAST),
    "exact":  (8192, 12288, 3, 2, 2, 2, fixca.PRECISION_EXACT),
    "linear": (8192, 12288, 3, 2, 2, 1, fixca.PRECISION_FAST),
    "none":   (8192, 12288, 3, 2, 2, 0, fixca.PRECISION_EXACT),
    "rgb8":   (4000, 6000, 3, 1, 1, 2, fixca.PRECISION_FAST),
    "rgb8lin": (4000, 6000, 3, 1, 1, 1, fixca.PRECISION_FAST),
    "rgba16": (4320, 7680, 4, 2, 2, 2, fixca.PRECISION_FAST),
    "rgb8_4k": (2160, 3840, 3, 1, 1, 2, fixca.PRECISION_FAST),
}[which]
h, w, ch, es, bpc, interp, flags = cfg
bpp = ch * es
pitch = (w * bpp + 127) // 128 * 128
src = torch.randint(0, 255, (h, pitch), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
p = fixca.FixCaParams(interpolation=interp, lens_x=w // 2, lens_y=h // 2, **KW)
for _ in range(reps):
    fixca.fix_ca_region_dev(src.data_ptr(), pitch, 0, h, dst.data_ptr(), pitch, 0, w, h, bpp, bpc, p, 0, h, flags,
                            torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(which, fixca.last_kernel(), "ok")
