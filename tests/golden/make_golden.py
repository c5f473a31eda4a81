#!/usr/bin/env python
"""Generate the golden vectors for the parity tests from the reference's own code.

Runs only where /root/reference is mounted (this container): it drives
oracle/_ref/libfixca_ref.so -- the reference's unmodified fix-ca.c -- over a seeded
synthetic suite and records the md5 of every output in tests/golden/golden.json
(inputs are regenerated from the seed by oracle.synth_image, so only digests are
stored).  It also decodes the reference's test fixture img-fix-ca/full-branches.jpg
the way GIMP does (libjpeg float DCT) into oracle/_ref/full-branches.rgb, which is
git-ignored but travels to the GPU box, and records the md5 chain of
tests/test1.md5.

    python tests/golden/make_golden.py                 # everything
    python tests/golden/make_golden.py --fixture-only  # just the decoded fixture
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

import numpy as np  # noqa: E402

import oracle as orc  # noqa: E402
from fixture_io import decode_jpeg_float_dct, encode_gimp_bmp24  # noqa: E402

REF_JPEG = "/root/reference/img-fix-ca/full-branches.jpg"
REF_MD5 = "/root/reference/tests/test1.md5"
FIXTURE_RAW = os.path.join(ROOT, "oracle", "_ref", "full-branches.rgb")
GOLDEN_JSON = os.path.join(HERE, "golden.json")

SHAPES = [(1, 1), (2, 3), (3, 2), (5, 40), (40, 5), (67, 131), (150, 200), (260, 389)]
FORMATS = [("u1", 3), ("u1", 4), ("u2", 3), ("u2", 4), ("u4", 3), ("u8", 4), ("f4", 3), ("f4", 4), ("f8", 3)]
PARAMS = {
    "zero": dict(),
    "lateral": dict(blue=6.0, red=-2.4),
    "directional": dict(x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9),
    "both": dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9),
    "extreme": dict(blue=30.0, red=-30.0, x_blue=30.0, x_red=-30.0, y_blue=-30.0, y_red=30.0),
    "mixed": dict(blue=-1.25, red=10.5, x_blue=-7.25, x_red=12.5, y_blue=2.5, y_red=-0.5),
}
LENSES = {"centre": None, "origin": (0, 0), "reset": (-1, -1), "off": (1000, -50), "quirk": (-858993459, 3)}


def max_dim(w, h, lx, ly):
    xc, yc = int(lx), int(ly)
    return max(xc, yc, w - xc, h - yc)


def cases():
    n = 0
    for (h, w), (dt, ch), (pname, pk), (lname, lens), interp in itertools.product(
            SHAPES, FORMATS, PARAMS.items(), LENSES.items(), (0, 1, 2)):
        n += 1
        # thin the matrix deterministically: 1 in 3 of the tiny shapes, 1 in 4 of the others
        if (n % 3) if h * w <= 300 else (n % 4):
            continue
        if dt == "u8" and interp != 0:
            continue            # u64 Linear/Cubic: 80-bit long double, unsupported by the CUDA path
        lx, ly = (w // 2, h // 2) if lens is None else lens
        m = max_dim(w, h, lx, ly)
        if m + pk.get("blue", 0.0) == 0 or m + pk.get("red", 0.0) == 0:
            continue            # the reference itself reads out of bounds here (scale = inf)
        yield dict(name="%dx%d-%sx%d-%s-%s-i%d" % (w, h, dt, ch, pname, lname, interp), h=h, w=w, ch=ch,
                   dtype=dt, seed=n, wide=(dt[0] == "f" and n % 2 == 0), interpolation=interp,
                   lens_x=float(lx), lens_y=float(ly), **pk)


PREVIEW_SHAPES = [(1, 1), (7, 40), (64, 64), (90, 131), (150, 203)]
PREVIEW_FORMATS = [("u1", 3), ("u1", 4), ("u2", 3), ("u2", 4), ("u4", 3), ("f4", 3), ("f4", 4), ("f8", 3)]
PREVIEW_SATURATION = (0.0, 35.0, -50.0, 100.0, -100.0)


def preview_cases():
    """The show_progress = FALSE call (fix-ca.c:656-657, :1322-1327): correction + saturate + centerline."""
    n = 0
    for (h, w), (dt, ch), sat, (lname, lens), interp in itertools.product(
            PREVIEW_SHAPES, PREVIEW_FORMATS, PREVIEW_SATURATION, LENSES.items(), (0, 1, 2)):
        n += 1
        if n % 5:
            continue
        lx, ly = (w // 2, h // 2) if lens is None else lens
        m = max_dim(w, h, lx, ly)
        if m + 3.0 == 0 or m - 2.0 == 0:
            continue
        yield dict(name="preview-%dx%d-%sx%d-sat%g-%s-i%d" % (w, h, dt, ch, sat, lname, interp), h=h, w=w, ch=ch,
                   dtype=dt, seed=50000 + n, wide=(dt[0] == "f" and n % 2 == 0), interpolation=interp,
                   lens_x=float(lx), lens_y=float(ly), saturation=sat, blue=3.0, red=-2.0, x_blue=0.7, y_red=-0.9)


def case_params(c) -> orc.Params:
    keys = ("blue", "red", "lens_x", "lens_y", "interpolation", "saturation", "x_blue", "x_red", "y_blue", "y_red")
    return orc.Params(**{k: c[k] for k in keys if k in c})


def write_fixture():
    img = decode_jpeg_float_dct(REF_JPEG)
    os.makedirs(os.path.dirname(FIXTURE_RAW), exist_ok=True)
    img.tofile(FIXTURE_RAW)
    return img


def main():
    img = write_fixture()
    if "--fixture-only" in sys.argv:
        print("wrote", FIXTURE_RAW, img.shape)
        return
    ref = orc.Reference()
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libfixca_ref.so (reference fix-ca.c, -O2 -ffp-contract=off)"}

    # --- the reference's own known-answer test ---
    h, w, _ = img.shape
    want_bmp = open(REF_MD5).read().split()[0]
    fixture = {"shape": [h, w, 3], "input_md5": hashlib.md5(img.tobytes()).hexdigest(), "test1_bmp_md5": want_bmp, "outputs": {}}
    for interp, lens in itertools.product((0, 1, 2), ((0, 0), (658, 1280))):
        p = orc.Params(blue=6.0, red=-2.4, lens_x=lens[0], lens_y=lens[1], interpolation=interp)
        got = ref.region(img, p)
        fixture["outputs"]["i%d-lens%d,%d" % (interp, lens[0], lens[1])] = hashlib.md5(got.tobytes()).hexdigest()
        if interp == 1 and lens == (0, 0):
            bmp = hashlib.md5(encode_gimp_bmp24(got)).hexdigest()
            assert bmp == want_bmp, (bmp, want_bmp)
            fixture["test1_raw_md5"] = hashlib.md5(got.tobytes()).hexdigest()
    out["fixture"] = fixture

    # --- seeded synthetic suite ---
    suite = []
    for c in cases():
        src = orc.synth_image(c["h"], c["w"], c["ch"], c["dtype"], c["seed"], c["wide"])
        got = ref.region(src, case_params(c))
        c["md5"] = hashlib.md5(got.tobytes()).hexdigest()
        suite.append(c)
    out["suite"] = suite
    # --- preview suite: the reference's own saturate()/centerline() on top of its pass; the HSV pair behind
    #     saturate() is libgimpcolor's, restated in oracle/ref_harness.c (parity unpinned for that pair) ---
    preview = []
    for c in preview_cases():
        src = orc.synth_image(c["h"], c["w"], c["ch"], c["dtype"], c["seed"], c["wide"])
        got = ref.region(src, case_params(c), preview=True)
        c["md5"] = hashlib.md5(got.tobytes()).hexdigest()
        preview.append(c)
    out["preview"] = preview
    with open(GOLDEN_JSON, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote %s: %d synthetic cases + %d preview cases + fixture chain (%s)" % (GOLDEN_JSON, len(suite), len(preview), want_bmp))


if __name__ == "__main__":
    main()
