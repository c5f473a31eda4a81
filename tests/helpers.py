"""Shared helpers for the parity tests."""
import hashlib
import json
import os

import numpy as np

import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "golden.json")
FIXTURE_RAW = os.path.join(ROOT, "oracle", "_ref", "full-branches.rgb")

GOLDEN_HALF = os.path.join(ROOT, "tests", "golden", "golden_half.json")
GOLDEN_U15 = os.path.join(ROOT, "tests", "golden", "golden_u15.json")
GOLDEN_U64 = os.path.join(ROOT, "tests", "golden", "golden_u64.json")
_golden = None
_golden_half = None
_golden_u15 = None
_golden_u64 = None


def golden_u64():
    """Digests of Linear / Cubic on u64 samples (tests/golden/make_golden_u64.py)."""
    global _golden_u64
    if _golden_u64 is None:
        with open(GOLDEN_U64) as f:
            _golden_u64 = json.load(f)
    return _golden_u64


def golden_u15():
    """Digests of the u15 extension (tests/golden/make_golden_u15.py)."""
    global _golden_u15
    if _golden_u15 is None:
        with open(GOLDEN_U15) as f:
            _golden_u15 = json.load(f)
    return _golden_u15


def golden_half():
    """Digests of the half-precision extension (tests/golden/make_golden_half.py)."""
    global _golden_half
    if _golden_half is None:
        with open(GOLDEN_HALF) as f:
            _golden_half = json.load(f)
    return _golden_half


def golden():
    global _golden
    if _golden is None:
        with open(GOLDEN) as f:
            _golden = json.load(f)
    return _golden


PKEYS = ("blue", "red", "lens_x", "lens_y", "interpolation", "saturation", "x_blue", "x_red", "y_blue", "y_red")


def oracle_params(c) -> orc.Params:
    return orc.Params(**{k: c[k] for k in PKEYS if k in c})


def fx_params(fx, c):
    return fx.FixCaParams(**{k: c[k] for k in PKEYS if k in c})


def case_image(c) -> np.ndarray:
    if c["dtype"] == "u64":     # full-range 64-bit samples, half of them at rounding boundaries when c["extremes"]
        return orc.synth_u64(c["h"], c["w"], c["ch"], c["seed"], c.get("extremes", False))
    if c["dtype"] == "u15":     # 15-bit samples in uint16 storage (bpc = 15)
        return orc.synth_u15(c["h"], c["w"], c["ch"], c["seed"], c.get("wide", False))
    return orc.synth_image(c["h"], c["w"], c["ch"], c["dtype"], c["seed"], c.get("wide", False))


def md5(a: np.ndarray) -> str:
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def fixture_image():
    """The reference's test photo, decoded as GIMP decodes it (see tests/golden/make_golden.py)."""
    if not os.path.exists(FIXTURE_RAW):
        return None
    h, w, c = golden()["fixture"]["shape"]
    return np.fromfile(FIXTURE_RAW, dtype=np.uint8).reshape(h, w, c)


def max_dim(w, h, lx, ly):
    xc, yc = int(lx), int(ly)
    return max(xc, yc, w - xc, h - yc)


def lsb_diff(a: np.ndarray, b: np.ndarray):
    """max |a-b| in LSB (ints) or absolute (floats), and the mismatching fraction."""
    if a.dtype.kind == "f":
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        d = np.where(np.isnan(a) & np.isnan(b), 0.0, d)
    else:
        d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    return d.max() if d.size else 0, float((d != 0).mean()) if d.size else 0.0
