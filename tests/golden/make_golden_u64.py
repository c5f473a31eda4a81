#!/usr/bin/env python
"""Golden vectors for Linear / Cubic on u64 samples (bpc = 8, fix-ca.c:728-733, :759-761).

The reference decodes 64-bit samples through the x87's 80-bit long double and encodes them through
`roundl(d * 18446744073709551615UL)` -- whose constant becomes the double 2^64, so d = 1.0 converts an out-of-range
value to uint64_t: undefined in C; the reference as compiled here (gcc, x86-64: oracle/_ref/libfixca_ref.so, the
unmodified fix-ca.c) wraps it to 0.  The CUDA library restates both steps in integer arithmetic
(gimp-fix-ca_b200/csrc/fixca_kernels.cuh: u64_get_pixel / u64_set_pixel); these digests pin it: a seeded synthetic
suite (RGB / RGBA, Linear and Cubic, the parameter and lens sets of make_golden.py; every second case draws from
the extremes -- 0, 1, 2^53 +- 1, 2^63, 2^64 - 2^11 +- 1, 2^64 - 1 -- where the two roundings and the wrap live) plus
preview (show_progress = FALSE) cases with the saturation boost.  Runs only where /root/reference is mounted.

    python tests/golden/make_golden_u64.py
"""
from __future__ import annotations

import hashlib
import itertools
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import oracle as orc  # noqa: E402
from make_golden import LENSES, PARAMS, case_params, max_dim  # noqa: E402

GOLDEN_JSON = os.path.join(HERE, "golden_u64.json")
SHAPES = [(1, 1), (3, 2), (5, 40), (67, 131), (150, 200)]
synth_u64 = orc.synth_u64


def cases():
    n = 0
    for (h, w), ch, (pname, pk), (lname, lens), interp in itertools.product(
            SHAPES, (3, 4), PARAMS.items(), LENSES.items(), (1, 2)):
        n += 1
        if n % 3:
            continue
        lx, ly = (w // 2, h // 2) if lens is None else lens
        m = max_dim(w, h, lx, ly)
        if m + pk.get("blue", 0.0) == 0 or m + pk.get("red", 0.0) == 0:
            continue
        yield dict(name="%dx%d-u64x%d-%s-%s-i%d" % (w, h, ch, pname, lname, interp), h=h, w=w, ch=ch, dtype="u64",
                   seed=90000 + n, extremes=(n % 2 == 0), interpolation=interp, lens_x=float(lx), lens_y=float(ly), **pk)


def preview_cases():
    n = 0
    for (h, w), ch, sat, interp in itertools.product([(7, 40), (90, 131)], (3, 4), (0.0, 35.0, -100.0), (0, 1, 2)):
        n += 1
        yield dict(name="preview-%dx%d-u64x%d-sat%g-i%d" % (w, h, ch, sat, interp), h=h, w=w, ch=ch, dtype="u64",
                   seed=91000 + n, extremes=(n % 2 == 0), interpolation=interp, lens_x=float(w // 2), lens_y=float(h // 2),
                   saturation=sat, blue=3.0, red=-2.0, x_blue=0.7, y_red=-0.9)


def main():
    ref = orc.Reference()
    out = {"generator": "tests/golden/make_golden_u64.py",
           "source": "oracle/_ref/libfixca_ref.so (the unmodified reference fix-ca.c, gcc x86-64)",
           "suite": [], "preview": []}
    for key, gen, prev in (("suite", cases, False), ("preview", preview_cases, True)):
        for c in gen():
            src = synth_u64(c["h"], c["w"], c["ch"], c["seed"], c["extremes"])
            got = ref.region(src, case_params(c), preview=prev)
            c["md5"] = hashlib.md5(got.tobytes()).hexdigest()
            out[key].append(c)
    with open(GOLDEN_JSON, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote %s: %d cases + %d preview cases" % (GOLDEN_JSON, len(out["suite"]), len(out["preview"])))


if __name__ == "__main__":
    main()
