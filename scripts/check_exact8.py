#!/usr/bin/env python
"""Developer probe (small enough for a sanitizer where one is available): the 8-bit EXACT forms (deferred: stream kernel + repair_patch_kernel; the
whole-region path of an image of exact ties; a frame batch; a band) on small images, checked against the compiled reference.
    python scripts/check_exact8.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fixca  # noqa: E402
import oracle as orc  # noqa: E402

chk = orc.best_checker()
KW = dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
for (h, w, ch, interp, kw) in ((301, 517, 3, 2, KW), (97, 640, 4, 1, KW), (203, 1031, 3, 1, dict(blue=0.0, red=0.0, x_blue=0.5, x_red=0.0, y_blue=0.0, y_red=-0.5)),
                               (64, 300, 4, 2, dict(blue=0.0, red=0.0, x_blue=0.5, x_red=-0.5, y_blue=0.5, y_red=0.5))):
    img = orc.synth_image(h, w, ch, "u1", seed=5 + h)
    p = dict(kw, lens_x=w // 2, lens_y=h // 2, interpolation=interp)
    want = chk.region(img, orc.Params(**p))
    got = fixca.correct(img, fixca.FixCaParams(**p), flags=fixca.PRECISION_EXACT)
    assert got.tobytes() == want.tobytes(), (h, w, ch, interp)
    print("ok", h, w, ch, interp, fixca.last_kernel(), flush=True)
# a batch of frames in one launch, and a band into a taller destination
h, w, ch, nf = 130, 517, 3, 3
p = fixca.FixCaParams(interpolation=2, lens_x=250, lens_y=60, **KW)
frames = [orc.synth_image(h, w, ch, "u1", seed=40 + k) for k in range(nf)]
bpp = ch
pitch = (w * bpp + 127) // 128 * 128
src = torch.zeros((nf, h, pitch), dtype=torch.uint8, device="cuda")
dst = torch.zeros_like(src)
for k, fr in enumerate(frames):
    src[k, :, :w * bpp] = torch.from_numpy(fr.reshape(h, w * bpp)).cuda()
st = torch.cuda.current_stream().cuda_stream
fixca.fix_ca_frames_dev(src.data_ptr(), pitch, pitch * h, dst.data_ptr(), pitch, pitch * h, nf, w, h, bpp, 1, p, fixca.PRECISION_EXACT | fixca.PADDING_SCRATCH, st)
torch.cuda.synchronize()
for k, fr in enumerate(frames):
    want = chk.region(fr, orc.Params(interpolation=2, lens_x=250, lens_y=60, **KW))
    assert dst[k, :, :w * bpp].cpu().numpy().tobytes() == want.tobytes(), k
print("ok batch", fixca.last_kernel(), flush=True)
y1, y2 = 40, 101
band = torch.full((h, pitch), 0x5A, dtype=torch.uint8, device="cuda")
fixca.fix_ca_region_dev(src[0].data_ptr(), pitch, 0, h, band.data_ptr(), pitch, 0, w, h, bpp, 1, p, y1, y2, fixca.PRECISION_EXACT | fixca.PADDING_SCRATCH, st)
torch.cuda.synchronize()
want = chk.region(frames[0], orc.Params(interpolation=2, lens_x=250, lens_y=60, **KW))
got = band[:, :w * bpp].cpu().numpy()
assert got[y1:y2].tobytes() == want[y1:y2].tobytes() and (got[:y1] == 0x5A).all() and (got[y2:] == 0x5A).all()
print("ok band", fixca.last_kernel(), flush=True)
