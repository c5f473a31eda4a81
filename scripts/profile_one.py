#!/usr/bin/env python
"""Run one workload a few times (for ncu):
    python scripts/profile_one.py <fast|exact|linear|none|rgb8|rgb8lin|rgb8exact|rgb8linexact|rgba16|rgb8_4k> [reps]
    python scripts/profile_one.py wl:<bench.py workload name>[:exact] [reps]      (bench.py's own shapes and parameters)
A 512 MB buffer is rewritten between launches, so every launch starts with an L2 that holds none of its input
(what bench.py's rotating buffer sets arrange): the captured DRAM traffic is the kernel's own."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fixca  # noqa: E402

KW = dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
which = sys.argv[1] if len(sys.argv) > 1 else "fast"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
frames = 0
if which.startswith("wl:"):
    import bench
    name = which.split(":")[1]
    w, h, ch, dts, interp, kw, lens = bench.WORKLOADS[name]
    es = np.dtype(dts).itemsize
    bpc = -es if dts.startswith("f") else es
    flags = fixca.PRECISION_EXACT if which.endswith(":exact") else fixca.PRECISION_FAST
    lx, ly = (w // 2, h // 2) if lens == "centre" else lens
    frames = min(32, bench.WORKLOAD_FRAMES.get(name, 0))
else:
    cfg = {
        "fast":   (8192, 12288, 3, 2, 2, 2, fixca.PRECISION_FAST),
        "exact":  (8192, 12288, 3, 2, 2, 2, fixca.PRECISION_EXACT),
        "linear": (8192, 12288, 3, 2, 2, 1, fixca.PRECISION_FAST),
        "none":   (8192, 12288, 3, 2, 2, 0, fixca.PRECISION_EXACT),
        "rgb8":   (4000, 6000, 3, 1, 1, 2, fixca.PRECISION_FAST),
        "rgb8lin": (4000, 6000, 3, 1, 1, 1, fixca.PRECISION_FAST),
        "rgb8exact": (4000, 6000, 3, 1, 1, 2, fixca.PRECISION_EXACT),
        "rgb8linexact": (4000, 6000, 3, 1, 1, 1, fixca.PRECISION_EXACT),
        "rgba16": (4320, 7680, 4, 2, 2, 2, fixca.PRECISION_FAST),
        "rgb8_4k": (2160, 3840, 3, 1, 1, 2, fixca.PRECISION_FAST),
    }[which]
    h, w, ch, es, bpc, interp, flags = cfg
    kw, lx, ly = KW, w // 2, h // 2
bpp = ch * es
pitch = (w * bpp + 127) // 128 * 128
nf = max(1, frames)
if bpc == -4:
    src = torch.rand((nf * h, pitch // 4), dtype=torch.float32, device="cuda").view(torch.uint8)
else:
    src = torch.randint(0, 255, (nf * h, pitch), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
p = fixca.FixCaParams(interpolation=interp, lens_x=lx, lens_y=ly, **kw)
st = torch.cuda.current_stream().cuda_stream
for i in range(reps):
    flush.fill_(i)
    if frames:
        fixca.fix_ca_frames_dev(src.data_ptr(), pitch, pitch * h, dst.data_ptr(), pitch, pitch * h, nf, w, h, bpp, bpc, p, flags, st)
    else:
        fixca.fix_ca_region_dev(src.data_ptr(), pitch, 0, h, dst.data_ptr(), pitch, 0, w, h, bpp, bpc, p, 0, h, flags, st)
torch.cuda.synchronize()
print(which, fixca.last_kernel(), "frames", frames, "ok")
