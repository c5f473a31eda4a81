#!/usr/bin/env python
"""Golden vectors of the u15 extension (bpc = 15).

The reference only rejects 15-bit samples (fix-ca.c:694-695), so the checker is the reference's own row loop with
the u15 rows written into color_size() / get_pixel() / set_pixel() in a scratch copy, in the pattern of its other
unsigned types (oracle/patch_u15.py -> oracle/_ref/libfixca_ref_u15.so); everything else is the unmodified
reference source.  Runs only where /root/reference is mounted; records md5 digests of the outputs for a seeded
synthetic suite (u15 RGB / RGBA in uint16 storage, all three interpolations, the parameter and lens sets of
make_golden.py; every second case also holds out-of-range codes 32769..65535) plus preview
(show_progress = FALSE) cases in tests/golden/golden_u15.json.

    python tests/golden/make_golden_u15.py
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

GOLDEN_JSON = os.path.join(HERE, "golden_u15.json")
SHAPES = [(1, 1), (3, 2), (5, 40), (67, 131), (150, 200), (260, 389)]


def cases():
    n = 0
    for (h, w), ch, (pname, pk), (lname, lens), interp in itertools.product(
            SHAPES, (3, 4), PARAMS.items(), LENSES.items(), (0, 1, 2)):
        n += 1
        if n % 4:       # (not 3: the interpolation is the fastest axis)
            continue
        lx, ly = (w // 2, h // 2) if lens is None else lens
        m = max_dim(w, h, lx, ly)
        if m + pk.get("blue", 0.0) == 0 or m + pk.get("red", 0.0) == 0:
            continue
        yield dict(name="%dx%d-u15x%d-%s-%s-i%d" % (w, h, ch, pname, lname, interp), h=h, w=w, ch=ch, dtype="u15",
                   seed=80000 + n, wide=(n % 8 == 0), interpolation=interp, lens_x=float(lx), lens_y=float(ly), **pk)


def preview_cases():
    n = 0
    for (h, w), ch, sat, interp in itertools.product([(7, 40), (90, 131)], (3, 4), (0.0, 35.0, -100.0), (0, 1, 2)):
        n += 1
        yield dict(name="preview-%dx%d-u15x%d-sat%g-i%d" % (w, h, ch, sat, interp), h=h, w=w, ch=ch, dtype="u15",
                   seed=81000 + n, wide=False, interpolation=interp, lens_x=float(w // 2), lens_y=float(h // 2),
                   saturation=sat, blue=3.0, red=-2.0, x_blue=0.7, y_red=-0.9)


def main():
    ref = orc.ReferenceU15()
    out = {"generator": "tests/golden/make_golden_u15.py",
           "source": "oracle/_ref/libfixca_ref_u15.so (reference fix-ca.c with u15 rows written in by oracle/patch_u15.py)",
           "suite": [], "preview": []}
    for key, gen, prev in (("suite", cases, False), ("preview", preview_cases, True)):
        for c in gen():
            src = orc.synth_u15(c["h"], c["w"], c["ch"], c["seed"], c["wide"])
            got = ref.region(src, case_params(c), preview=prev, bpc=orc.BPC_U15)
            c["md5"] = hashlib.md5(got.tobytes()).hexdigest()
            out[key].append(c)
    with open(GOLDEN_JSON, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote %s: %d cases + %d preview cases" % (GOLDEN_JSON, len(out["suite"]), len(out["preview"])))


if __name__ == "__main__":
    main()
