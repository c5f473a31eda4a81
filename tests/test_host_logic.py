"""CPU: the C-ABI library loads, exports what include/fixca_cuda.h declares, and its host-side
logic (band halos, band split, lens reset, range check, format probe, argument errors) agrees
with the oracle.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as orc
from helpers import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fixca_cuda.h")).read()
    return sorted(set(re.findall(r"FIXCA_API[^;(]*?\b(fixca_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(fx):
    lib = fx.load()
    names = declared_symbols()
    assert len(names) >= 19 and set(names) == set(fx.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_params_layout_matches_reference(fx):
    # FixCaParams is 80 bytes on x86-64 (fix-ca.c:70-82); the plug-in passes its struct through a cast
    assert ctypes.sizeof(fx.FixCaParams) == 80
    assert fx.FixCaParams.interpolation.offset == 36 and fx.FixCaParams.saturation.offset == 40
    assert fx.FixCaParams.y_red.offset == 72
    p = fx.FixCaParams()
    fx.load().fixca_params_default(ctypes.byref(p))
    assert (p.blue, p.red, p.lens_x, p.lens_y, p.interpolation, p.saturation) == (0, 0, -1, -1, 1, 0)


def test_no_gpu_fails_loudly(fx):
    if fx.device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is for CPU-only hosts")
    img = np.zeros((8, 8, 3), np.uint8)
    with pytest.raises(fx.FixCaError) as e:
        fx.correct(img, fx.FixCaParams())
    assert e.value.code == fx.ERR_NO_DEVICE and "no CPU path" in str(e.value)


def test_argument_errors_before_any_gpu_work(fx):
    img = np.zeros((8, 8, 3), np.uint8)
    out = np.zeros_like(img)
    P = fx.FixCaParams

    def rc(*a, **k):
        with pytest.raises(fx.FixCaError) as e:
            fx.fix_ca_region(*a, **k)
        return e.value.code

    assert rc(img, out, 8, 8, 3, 1, P(), 1, 8, 0, 8) == fx.ERR_REGION           # x1 != 0
    assert rc(img, out, 8, 8, 3, 1, P(), 0, 7, 0, 8) == fx.ERR_REGION
    # ... and with the opt-in (FIXCA_COLUMN_SELECTION) only sane column ranges pass the argument check
    assert rc(img, out, 8, 8, 3, 1, P(), 3, 3, 0, 8, True, fx.COLUMN_SELECTION) == fx.ERR_REGION
    assert rc(img, out, 8, 8, 3, 1, P(), 0, 9, 0, 8, True, fx.COLUMN_SELECTION) == fx.ERR_REGION
    assert rc(img, out, 8, 8, 3, -2, P(), 0, 8, 0, 8) == fx.ERR_FORMAT          # half
    assert rc(img, out, 8, 8, 3, -99, P(), 0, 8, 0, 8) == fx.ERR_FORMAT
    assert rc(img, out, 8, 8, 5, 1, P(), 0, 8, 0, 8) == fx.ERR_FORMAT
    assert rc(img, out, 8, 8, 3, 1, P(interpolation=3), 0, 8, 0, 8) == fx.ERR_INTERP
    assert rc(img, out, 8, 8, 3, 1, P(), 0, 8, 0, 9) == fx.ERR_ARG
    assert rc(img, out, 0, 8, 3, 1, P(), 0, 0, 0, 8) == fx.ERR_ARG
    assert rc(0, out, 8, 8, 3, 1, P(), 0, 8, 0, 8) == fx.ERR_ARG
    # max_dim + amount == 0: the reference itself indexes out of bounds (scale = inf)
    assert rc(img, out, 8, 8, 3, 1, P(lens_x=4, lens_y=4, red=-4.0), 0, 8, 0, 8) == fx.ERR_DEGENERATE
    # u64 Linear / Cubic is computed (the x87 steps of fix-ca.c:728-733, :759-761 restated): no format error, only the missing GPU
    big = np.zeros((8, 8, 3), np.uint64)
    assert rc(big, np.zeros_like(big), 8, 8, 24, 8, P(interpolation=2), 0, 8, 0, 8) in (0, fx.ERR_NO_DEVICE)


def test_multi_gpu_entries_check_their_arguments_first(fx):
    """fixca_cuda_frames_multi / fixca_cuda_frame_*: argument errors come before any device work; without a GPU the
    allocation fails loudly (no CPU path)."""
    import ctypes

    L = fx.load()
    img = np.zeros((8, 8, 3), np.uint8)
    out = np.zeros_like(img)
    vp = ctypes.c_void_p
    srcs, dsts = (vp * 1)(img.ctypes.data), (vp * 1)(out.ctypes.data)
    P = fx.FixCaParams()
    assert L.fixca_cuda_frames_multi(srcs, dsts, 1, 8, 8, 3, 1, ctypes.byref(P), 0, None, 0) == fx.ERR_ARG      # ndev 0
    assert L.fixca_cuda_frames_multi(srcs, dsts, 1, 8, 8, 3, 1, ctypes.byref(P), 0, None, 17) == fx.ERR_ARG
    assert L.fixca_cuda_frames_multi(None, dsts, 1, 8, 8, 3, 1, ctypes.byref(P), 0, None, 1) == fx.ERR_ARG
    assert L.fixca_cuda_frames_multi(srcs, dsts, 0, 8, 8, 3, 1, ctypes.byref(P), 0, None, 1) == fx.OK          # empty stream
    ptr = vp()
    assert L.fixca_cuda_frame_open(None, ctypes.byref(ptr)) == fx.ERR_ARG
    assert L.fixca_cuda_frame_alloc(0, ctypes.byref(ptr), ctypes.create_string_buffer(64)) == fx.ERR_ARG
    assert L.fixca_cuda_frame_close(None) == fx.OK and L.fixca_cuda_frame_free(None) == fx.OK
    if fx.device_count() == 0:
        assert L.fixca_cuda_frame_alloc(1 << 20, ctypes.byref(ptr), ctypes.create_string_buffer(64)) == fx.ERR_NO_DEVICE
        assert L.fixca_cuda_frames_multi(srcs, dsts, 1, 8, 8, 3, 1, ctypes.byref(P), 0, None, 1) == fx.ERR_NO_DEVICE


def test_band_source_rows_match_oracle_tables(fx, restatement):
    rng = np.random.default_rng(5)
    for n in range(200):
        w, h = int(rng.integers(1, 400)), int(rng.integers(2, 400))
        interp = n % 3
        kw = dict(blue=float(rng.uniform(-30, 30)), red=float(rng.uniform(-30, 30)),
                  lens_x=float(rng.integers(-3, w + 3)), lens_y=float(rng.integers(-3, h + 3)),
                  x_blue=float(rng.uniform(-30, 30)), x_red=float(rng.uniform(-30, 30)),
                  y_blue=float(rng.uniform(-30, 30)), y_red=float(rng.uniform(-30, 30)), interpolation=interp)
        m = max(int(kw["lens_x"]), int(kw["lens_y"]), w - int(kw["lens_x"]), h - int(kw["lens_y"]))
        if m + kw["blue"] == 0 or m + kw["red"] == 0:
            continue
        y1 = int(rng.integers(0, h - 1))
        y2 = int(rng.integers(y1 + 1, h + 1))
        lo, hi = fx.band_source_rows(w, h, fx.FixCaParams(**kw), y1, y2)
        want_lo, want_hi = y1, y2 - 1
        for ch in (0, 1):
            idx, _ = restatement.axis(w, h, orc.Params(**kw), ch, 1)
            seg = idx[y1:y2]
            if interp == 0:
                a, b = seg.min(), seg.max()
            elif interp == 1:
                a, b = seg.min(), min(seg.max() + 1, h - 1)
            else:
                a, b = max(seg.min() - 1, 0), min(seg.max() + 2, h - 1)
            want_lo, want_hi = min(want_lo, a), max(want_hi, b)
        assert (lo, hi) == (want_lo, want_hi), (n, w, h, kw, y1, y2)


def test_halo_bound_from_survey(fx):
    # 6000x4000 centre lens, a = +-30, shift = +-30 -> ~50 rows of displacement
    p = fx.FixCaParams(blue=30, red=-30, y_blue=30, y_red=-30, lens_x=3000, lens_y=2000, interpolation=2)
    lo, hi = fx.band_source_rows(6000, 4000, p, 0, 500)
    assert lo == 0 and 499 + 10 <= hi <= 499 + 64
    p = fx.FixCaParams(blue=6, red=-2.4, y_blue=1, lens_x=6144, lens_y=4096, interpolation=2)
    lo, hi = fx.band_source_rows(12288, 8192, p, 4096, 6144)
    assert 4096 - 4 <= lo <= 4096 and 6143 <= hi <= 6143 + 8


def test_split_bands(fx):
    for y1, y2, n in [(0, 8192, 8), (3, 1000, 7), (0, 5, 8), (10, 10, 3)]:
        b = fx.split_bands(y1, y2, n)
        assert b[0][0] == y1 and b[-1][1] == y2
        assert all(b[i][1] == b[i + 1][0] for i in range(n - 1))
        sizes = [e - s for s, e in b]
        assert max(sizes) - min(sizes) <= 1


def test_resolve_lens_check_params_color_size(fx, restatement):
    for (w, h, lx, ly) in [(1441, 2561, -1, -1), (100, 50, 0, 0), (100, 50, 100, 50), (100, 50, 30.5, 20.25), (7, 9, 7.5, -3)]:
        assert fx.resolve_lens(w, h, lx, ly) == restatement.resolve_lens(w, h, lx, ly)
    P = fx.FixCaParams
    assert fx.check_params(P(blue=30, red=-30, x_blue=30, y_red=-30)) == fx.OK
    for k in ("blue", "red", "x_blue", "x_red", "y_blue", "y_red"):
        assert fx.check_params(P(**{k: 30.01})) == fx.ERR_RANGE
        assert fx.check_params(P(**{k: -30.01})) == fx.ERR_RANGE
    assert fx.check_params(P(interpolation=3)) == fx.ERR_INTERP
    assert fx.check_params(P(lens_x=1e9)) == fx.OK           # lens is not range-checked (fix-ca.c:279-292)
    for name, bpp, want in [("R'G'B' u8", 3, 1), ("R'G'B'A u8", 4, 1), ("RGB u16", 6, 2), ("RGBA u16", 8, 2),
                            ("RGB u32", 12, 4), ("RGBA u32", 16, 4), ("RGBA u64", 32, 8), ("RGB float", 12, -4),
                            ("RGBA double", 32, -8), ("RGB half", 6, -99), ("RGB u15", 6, -99), ("Y u8", 1, -99),
                            ("RGBA u64x", 40, -99)]:
        assert fx.color_size(name, bpp) == want


def test_color_size_matches_reference(fx, reference):
    for name in ("R'G'B' u8", "RGBA u16", "RGB u32", "RGB float", "RGBA double", "RGB half", "RGB u15", "CMYK u8", "RGB u8 double"):
        for bpp in (1, 2, 3, 4, 6, 8, 11, 12, 16, 23, 24, 32, 33):
            assert fx.color_size(name, bpp) == reference.color_size(name, bpp), (name, bpp)


def test_exact_decode_division_is_correctly_rounded_for_every_8_and_16_bit_sample():
    """ExactF64::div_by_max (csrc/fixca_kernels.cuh) replaces get_pixel's v / max (fix-ca.c:717-722) by
    q0 = RN(v * r), e = fma(-max, q0, v), q = fma(e, r, q0) with r = RN(1 / max).  Evaluated here in exact rational
    arithmetic (Fraction -> float rounds to nearest-even, like the device's _rn operations) for every sample
    value: the result equals IEEE division, so EXACT mode stays bit-identical to the reference."""
    from fractions import Fraction as Fr
    for m in (255, 65535):
        r = float(Fr(1, m))
        assert r == 1.0 / m
        for v in range(m + 1):
            q0 = float(Fr(v) * Fr(r))
            e = float(Fr(v) - Fr(m) * Fr(q0))
            assert Fr(e) == Fr(v) - Fr(m) * Fr(q0)        # the residual is exact
            assert float(Fr(q0) + Fr(e) * Fr(r)) == v / float(m), (m, v)


def test_color_size_ext_u15(fx):
    """fixca_color_size_ext: color_size() with both "TODO for another day" rows answered (fix-ca.c:692-695)."""
    assert fx.color_size("R'G'B' u15", 6) == -99 and fx.color_size_half("R'G'B' u15", 6) == -99
    assert fx.color_size_ext("R'G'B' u15", 6) == fx.BPC_U15 == 15 and fx.color_size_ext("R'G'B'A u15", 8) == 15
    assert fx.color_size_ext("Y u15", 2) == -99                     # neither RGB nor RGBA of 16-bit storage
    assert fx.color_size_ext("R'G'B' half", 6) == -2
    for name, bpp in (("R'G'B' u8", 3), ("RGBA u16", 8), ("RGB float", 12), ("RGBA double", 32), ("CMYK u8", 4)):
        assert fx.color_size_ext(name, bpp) == fx.color_size(name, bpp)


def test_host_alloc_without_gpu_returns_null(fx):
    """fixca_cuda_host_alloc: NULL (the caller keeps its own allocator) when no GPU is usable; freeing NULL is a no-op."""
    L = fx.load()
    L.fixca_cuda_host_free(None)
    if fx.device_count() > 0:
        pytest.skip("a GPU is present; the pinned path is covered by the gpu tests")
    assert not L.fixca_cuda_host_alloc(4096)
    with pytest.raises(fx.FixCaError):
        fx.PinnedBuffer(4096)


def test_non_finite_parameters_are_rejected(fx):
    """NaN / infinite amounts, shifts and lens coordinates (undefined behaviour in the reference: (int) NaN,
    floor(NaN) as a row index) and lens coordinates beyond +-2^30 fail with ERR_ARG before any planning."""
    P = fx.FixCaParams
    nan, inf = float("nan"), float("inf")
    for kw in (dict(y_red=nan), dict(x_blue=inf), dict(blue=nan), dict(red=-inf), dict(lens_x=nan), dict(lens_y=inf),
               dict(lens_x=3.0e9), dict(lens_y=-2.0e9)):
        p = P(interpolation=1, **{"lens_x": 50, "lens_y": 40, **kw})
        with pytest.raises(fx.FixCaError) as e:
            fx.band_source_rows(100, 80, p, 10, 20)
        assert e.value.code == fx.ERR_ARG, kw
    # the lens value the PDB quirk produces stays legal
    assert fx.band_source_rows(100, 80, P(lens_x=-858993459.0, lens_y=0.0, interpolation=2), 10, 20)[0] >= 0


def test_flags_are_validated_by_every_entry(fx):
    """Unknown flag bits, the preview overlay on the frame entries and column selections outside region_ex are
    argument errors everywhere (checked before any device work)."""
    import ctypes

    L = fx.load()
    img = np.zeros((8, 8, 3), np.uint8)
    out = np.zeros_like(img)
    p = fx.FixCaParams(lens_x=4, lens_y=4)
    vp = ctypes.c_void_p
    srcs, dsts = (vp * 1)(img.ctypes.data), (vp * 1)(out.ctypes.data)
    assert L.fixca_cuda_region_ex(img.ctypes.data, out.ctypes.data, 8, 8, 3, 1, ctypes.byref(p), 0, 8, 0, 8, 1, 0x4000, -1) == fx.ERR_ARG
    assert L.fixca_cuda_region_ex(img.ctypes.data, out.ctypes.data, 8, 8, 3, 1, ctypes.byref(p), 0, 8, 0, 8, 1, 0x2, -1) == fx.ERR_ARG
    assert L.fixca_cuda_frames(srcs, dsts, 1, 8, 8, 3, 1, ctypes.byref(p), fx.PREVIEW_OVERLAY, 0) == fx.ERR_UNSUPPORTED
    assert L.fixca_cuda_frames_dev(img.ctypes.data, 24, 192, out.ctypes.data, 24, 192, 1, 8, 8, 3, 1, ctypes.byref(p),
                                   fx.PREVIEW_OVERLAY, None) == fx.ERR_UNSUPPORTED
    assert L.fixca_cuda_region_multi(img.ctypes.data, out.ctypes.data, 8, 8, 3, 1, ctypes.byref(p), 0, 8,
                                     fx.COLUMN_SELECTION, None, 1) == fx.ERR_UNSUPPORTED
    big = np.zeros((8, 8, 3), np.uint64)
    ps = fx.FixCaParams(lens_x=4, lens_y=4, interpolation=0, saturation=20.0)
    # (the preview's saturation boost on u64 samples is computed since r02: past the flag checks, only the GPU is missing here)
    assert L.fixca_cuda_region_multi(big.ctypes.data, big.ctypes.data, 8, 8, 24, 8, ctypes.byref(ps), 0, 8,
                                     fx.PREVIEW_OVERLAY, None, 1) in (0, fx.ERR_NO_DEVICE)


def test_color_size_half_extension(fx):
    """fixca_color_size_half: the reference's color_size() with its commented-out half line enabled
    (fix-ca.c:692-693); fixca_color_size keeps answering like the shipped reference."""
    assert fx.color_size("R'G'B' half", 6) == -99 and fx.color_size_half("R'G'B' half", 6) == -2
    assert fx.color_size_half("R'G'B'A half", 8) == -2
    for name, bpp in (("R'G'B' u8", 3), ("R'G'B'A u16", 8), ("RGB float", 12), ("RGBA double", 32), ("R'G'B' u15", 6), ("Y u8", 1)):
        assert fx.color_size_half(name, bpp) == fx.color_size(name, bpp)


def test_wide_emit_fixed_point_rounding_and_boundary_test():
    """vertical_emit_wide (fixca_strip.cuh) rounds a 16-bit result and tests its distance from a rounding
    boundary with ONE FP64 addition: t = v + 0.5 + 1.5 * 2^22, the integer from a funnel shift of t's words, the test from
    the low 30 bits.  Restated here bit for bit on 8 M values (uniform, and planted within a few 1e-6 of boundaries, on
    them, and on integers): every value within kEps = 1e-6 LSB of a boundary is flagged, every value that is not flagged
    gets round-to-nearest, and the flag rate stays at ~2 kEps."""
    rng = np.random.default_rng(1)
    magic = np.float64(6291456.5)
    expo = (0x415 << 22) & 0xFFFFFFFF
    base = expo + (1 << 21)
    eps = int(1e-6 * 1073741824.0) + 26

    def emit(v, kmax):
        bits = (v + magic).view(np.uint64)
        lo, hi = bits & np.uint64(0xFFFFFFFF), bits >> np.uint64(32)
        word = (((hi << np.uint64(2)) | (lo >> np.uint64(30))) & np.uint64(0xFFFFFFFF)).astype(np.int64)
        n = np.clip(word, base, base + kmax) - base
        assert ((np.clip(word, base, base + kmax) & 0xFFFF) == (n & 0xFFFF)).all()     # what the 16-bit store keeps
        flag = (((lo + np.uint64(eps)) << np.uint64(2)) & np.uint64(0xFFFFFFFF)) <= np.uint64((2 * eps) << 2)
        return n, flag

    v = rng.uniform(-40000, 105000, 4_000_000)
    k = rng.integers(-39000, 104000, 1_000_000).astype(np.float64)
    d = rng.uniform(-3e-6, 3e-6, 1_000_000)
    v = np.concatenate([v, k + 0.5 + d, k + 0.5, k, k + 0.5 + np.sign(d) * 1.05e-6])
    dist = np.abs((v - np.floor(v)) - 0.5)
    for kmax in (65535, 32768):
        n, flag = emit(v, kmax)
        want = np.clip(np.floor(v + 0.5), 0, kmax).astype(np.int64)
        assert not ((dist <= 1e-6) & ~flag).any()
        assert not (~flag & (n != want)).any()
        assert flag[:4_000_000].mean() < 4e-6 and dist[flag].max() < 1.1e-6
