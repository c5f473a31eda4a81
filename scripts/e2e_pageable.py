#!/usr/bin/env python
"""Developer probe: the drop-in call as the plug-in makes it -- pageable host buffers (g_new in fix-ca.c:366-367)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
import numpy as np
import fixca
h, w = 8192, 12288
rng = np.random.default_rng(1)
img = rng.integers(0, 65535, size=(h, w, 3), dtype=np.uint16)
out = np.zeros_like(img)
kw = dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9, lens_x=w // 2, lens_y=h // 2, interpolation=2)
p = fixca.FixCaParams(**kw)
for name, flags in (("EXACT (the ABI default)", fixca.PRECISION_EXACT), ("FAST", fixca.PRECISION_FAST)):
    for rep in range(3):
        t0 = time.perf_counter()
        fixca.correct(img, p, out=out, flags=flags)
        dt = time.perf_counter() - t0
        print("%-24s pageable 100 MP RGB16 Cubic: %.1f ms = %.0f MP/s (%s)" % (name, dt * 1e3, h * w / 1e6 / dt, fixca.last_kernel()), flush=True)
