#!/usr/bin/env python
"""Run the 128-frame 4K RGB8 Cubic batch a few times (for ncu): python scripts/profile_batch.py [frames] [interp]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
import torch  # noqa: E402

import fixca  # noqa: E402

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 128
interp = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h, w, bpp = 2160, 3840, 3
pitch = (w * bpp + 127) // 128 * 128
src = torch.randint(0, 255, (nf, h, pitch), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
p = fixca.FixCaParams(interpolation=interp, lens_x=w // 2, lens_y=h // 2, blue=1.0, red=-1.5, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
for _ in range(3):
    fixca.fix_ca_frames_dev(src.data_ptr(), pitch, pitch * h, dst.data_ptr(), pitch, pitch * h, nf, w, h, bpp, 1, p,
                            fixca.PRECISION_FAST, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("batch", nf, fixca.last_kernel(), "ok")
