"""CPU: pin the oracle.  (a) the restatement (oracle/fixca_oracle.c) against the committed golden
digests, which were produced by the reference's own fix-ca.c; (b) against that code directly when
oracle/_ref is present; (c) the reference's only known-answer test, tests/test1.md5, through the
JPEG -> pass -> BMP chain; (d) the plug-in surface (run(), color_size, lens reset) on the fake GIMP."""
import hashlib
import itertools

import numpy as np
import pytest

import oracle as orc
from fixture_io import encode_gimp_bmp24
from helpers import case_image, fixture_image, golden, golden_half, golden_u15, golden_u64, max_dim, md5, oracle_params


def test_restatement_matches_golden_suite(restatement):
    bad = [c["name"] for c in golden()["suite"] if md5(restatement.region(case_image(c), oracle_params(c))) != c["md5"]]
    assert not bad, "%d of %d golden cases differ, first: %s" % (len(bad), len(golden()["suite"]), bad[:5])


def test_reference_matches_golden_suite(reference):
    # the digests must be reproducible from the reference build that travels with the repo
    sample = golden()["suite"][::7]
    bad = [c["name"] for c in sample if md5(reference.region(case_image(c), oracle_params(c))) != c["md5"]]
    assert not bad, bad[:5]


def test_restatement_equals_reference_random(reference, restatement):
    rng = np.random.default_rng(20240518)
    for n in range(300):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        dt = ["u1", "u2", "u4", "f4", "f8", "u8"][n % 6]
        ch = 3 + (n // 6) % 2
        p = orc.Params(blue=float(rng.uniform(-30, 30)), red=float(rng.uniform(-30, 30)),
                       lens_x=float(rng.integers(-5, w + 5)), lens_y=float(rng.integers(-5, h + 5)),
                       interpolation=int(rng.integers(0, 3)),
                       x_blue=float(rng.uniform(-30, 30)), x_red=float(rng.uniform(-30, 30)),
                       y_blue=float(rng.uniform(-30, 30)), y_red=float(rng.uniform(-30, 30)))
        m = max_dim(w, h, p.lens_x, p.lens_y)
        if m + p.blue == 0 or m + p.red == 0:
            continue
        img = orc.synth_image(h, w, ch, dt, seed=n, wide=bool(n % 2))
        assert reference.region(img, p).tobytes() == restatement.region(img, p).tobytes(), (n, h, w, dt, ch, p)


@pytest.mark.parametrize("which", ["restatement", "reference"])
def test_known_answer_test1_md5(which, request):
    """tests/Makefile.am:18 + tests/test1.md5: Linear, blue 6.0, red -2.4, effective lens (0,0)."""
    img = fixture_image()
    if img is None:
        pytest.skip("oracle/_ref/full-branches.rgb not generated (needs /root/reference at build time)")
    chk = request.getfixturevalue(which)
    fxg = golden()["fixture"]
    assert md5(img) == fxg["input_md5"]
    out = chk.region(img, orc.Params(blue=6.0, red=-2.4, lens_x=0, lens_y=0, interpolation=1))
    assert md5(out) == fxg["test1_raw_md5"]
    assert hashlib.md5(encode_gimp_bmp24(out)).hexdigest() == fxg["test1_bmp_md5"] == "c472550cda23c8cb717853ac0dd93e2b"


def test_fixture_variants(restatement):
    img = fixture_image()
    if img is None:
        pytest.skip("fixture not generated")
    for interp, lens in itertools.product((0, 1, 2), ((0, 0), (658, 1280))):
        p = orc.Params(blue=6.0, red=-2.4, lens_x=lens[0], lens_y=lens[1], interpolation=interp)
        assert md5(restatement.region(img, p)) == golden()["fixture"]["outputs"]["i%d-lens%d,%d" % (interp, *lens)]


def test_band_and_thread_invariance(restatement):
    img = orc.synth_image(211, 157, 4, "u2", 11)
    p = orc.Params(blue=4, red=-3, lens_x=60, lens_y=100, interpolation=2, x_blue=1.5, y_red=-2.25)
    full = restatement.region(img, p)
    part = np.full_like(img, 9)
    restatement.region(img, p, 40, 97, dst=part)
    assert (part[40:97] == full[40:97]).all() and (part[:40] == 9).all() and (part[97:] == 9).all()
    assert (restatement.region(img, p, threads=4) == full).all()


def test_zero_params_is_identity(restatement):
    for dt, interp in itertools.product(("u1", "u2", "u4", "f4", "f8"), (0, 1, 2)):
        img = orc.synth_image(33, 47, 3, dt, 3)
        assert (restatement.region(img, orc.Params(interpolation=interp, lens_x=20, lens_y=10)) == img).all()


# ---- plug-in surface through the reference's real run() on the fake GIMP ----
def test_run_reproduces_test1_through_pdb(reference):
    img = fixture_image()
    if img is None:
        pytest.skip("fixture not generated")
    px = img.copy()
    # the PDB call of tests/Makefile.am:18: lens 658,1280 arrive as 0,0 through d_int32 (fix-ca.c:254,258)
    st = reference.run(px, "R'G'B' u8", 1, 12, blue=6.0, red=-2.4, lens_x=658, lens_y=1280, interpolation=1)
    assert st == 3
    assert hashlib.md5(encode_gimp_bmp24(px)).hexdigest() == golden()["fixture"]["test1_bmp_md5"]
    assert reference.counter(0) == 1 and reference.counter(1) == 2560 // 8 + 1


def test_run_argument_handling(reference):
    img = orc.synth_image(20, 30, 3, "u1", 1)
    assert reference.run(img.copy(), "R'G'B' u8", 1, 4) == 1           # too few params
    assert reference.run(img.copy(), "R'G'B' u8", 1, 13) == 1          # too many
    assert reference.run(img.copy(), "R'G'B' u8", 1, 5, proc_name="Fix-CA") == 1   # wrong name for the test build
    assert reference.run(img.copy(), "R'G'B' u8", 1, 5, blue=30.5) == 1
    assert "out of range" in reference.last_message()
    assert reference.run(img.copy(), "R'G'B' u8", 1, 8, interpolation=3) == 1
    assert reference.run(img.copy(), "R'G'B' half", 1, 5, blue=1.0) == 1
    assert "Invalid color type" in reference.last_message()
    # 5 params: interpolation defaults to None, shifts 0 (fix-ca.c:251-278)
    a = img.copy()
    assert reference.run(a, "R'G'B' u8", 1, 5, blue=2.0, red=-1.0) == 3
    want = reference.region(img, orc.Params(blue=2.0, red=-1.0, lens_x=-1, lens_y=-1, interpolation=0))
    assert (a == want).all()


def test_color_size_and_lens_reset(reference, restatement):
    for name, bpp, want in [("R'G'B' u8", 3, 1), ("R'G'B'A u8", 4, 1), ("RGB u16", 6, 2), ("RGBA u16", 8, 2),
                            ("RGB u32", 12, 4), ("RGBA u32", 16, 4), ("RGB float", 12, -4), ("RGBA double", 32, -8),
                            ("RGB half", 6, -99), ("RGB u15", 6, -99), ("Y u8", 1, -99)]:
        assert reference.color_size(name, bpp) == want
    for (w, h, lx, ly) in [(1441, 2561, -1, -1), (100, 50, 0, 0), (100, 50, 100, 50), (100, 50, 30.5, 20.25), (7, 9, 7.5, -3)]:
        assert reference.dialog_lens(w, h, lx, ly) == restatement.resolve_lens(w, h, lx, ly)
    assert reference.lib.ref_sizeof_params() == 80


def test_preview_epilogue_restatement_matches_golden(restatement):
    """show_progress = FALSE (fix-ca.c:1322-1327): saturate() + centerline() on top of the pass.  The digests were
    produced by the reference's own compiled saturate()/centerline(); the HSV pair underneath is libgimpcolor's
    (not in the reference tree), restated identically on both sides -- parity unpinned for that pair only."""
    from helpers import case_image, golden, md5, oracle_params

    bad = []
    for c in golden()["preview"]:
        got = restatement.region(case_image(c), oracle_params(c), preview=True)
        if md5(got) != c["md5"]:
            bad.append(c["name"])
    assert not bad, "%d of %d preview cases differ, first: %s" % (len(bad), len(golden()["preview"]), bad[:5])


def test_preview_centerline_shape(restatement):
    """Spot-check the overlay itself: with zero shifts and no saturation boost the preview call changes exactly the
    pixels on the lens row / column / diagonals, to pure 0 or 1."""
    import oracle as orc

    img = orc.synth_image(41, 57, 3, "u1", 5)
    p = orc.Params(lens_x=20, lens_y=13, interpolation=1)
    plain = restatement.region(img, p)
    prev = restatement.region(img, p, preview=True)
    changed = (plain != prev).any(axis=2)
    ys, xs = np.nonzero(changed)
    for y, x in zip(ys, xs):
        dy = abs(y - 13)
        assert y == 13 or x in (20, 20 - dy, 20 + dy)
        assert prev[y, x, 0] == prev[y, x, 1] == prev[y, x, 2] and prev[y, x, 0] in (0, 255)
    assert prev[13].reshape(-1, 3).min() >= 0 and set(np.unique(prev[13])) <= {0, 255}


# ---------------------------------------------------------------------------------------------
# half precision (bpc = -2): the reference's commented-out lines enabled (oracle/patch_half.py)
# ---------------------------------------------------------------------------------------------
def test_half_restatement_matches_golden(restatement):
    g = golden_half()
    bad = [c["name"] for c in g["suite"] if md5(restatement.region(case_image(c), oracle_params(c))) != c["md5"]]
    bad += [c["name"] for c in g["preview"]
            if md5(restatement.region(case_image(c), oracle_params(c), preview=True)) != c["md5"]]
    assert not bad, "%d half cases differ, first: %s" % (len(bad), bad[:5])


def test_half_reference_build_matches_golden_and_leaves_other_formats_alone(reference, restatement):
    if not orc.ReferenceHalf.available():
        pytest.skip("oracle/_ref/libfixca_ref_half.so not built (needs /root/reference at build time)")
    half = orc.ReferenceHalf()
    sample = golden_half()["suite"][::5]
    bad = [c["name"] for c in sample if md5(half.region(case_image(c), oracle_params(c))) != c["md5"]]
    assert not bad, bad[:5]
    # the three enabled fragments are the only difference: every other format computes what the unpatched build does
    for c in golden()["suite"][::97]:
        assert md5(half.region(case_image(c), oracle_params(c))) == c["md5"], c["name"]
    # color_size: -2 for half names only (fix-ca.c:692-693 enabled), -99 in the unpatched reference
    assert half.color_size("R'G'B' half", 6) == -2 and half.color_size("R'G'B'A half", 8) == -2
    assert reference.color_size("R'G'B' half", 6) == -99
    assert half.color_size("R'G'B' u16", 6) == 2 and half.color_size("R'G'B' float", 12) == -4


# ---------------------------------------------------------------------------------------------
# u64 Linear / Cubic: long double get_pixel, the out-of-range conversion in set_pixel (fix-ca.c:728-733, :759-761)
# ---------------------------------------------------------------------------------------------
def test_u64_restatement_matches_golden(restatement):
    """The digests come from the compiled reference (oracle/_ref); the restatement takes the same x87 steps."""
    g = golden_u64()
    bad = [c["name"] for c in g["suite"] if md5(restatement.region(case_image(c), oracle_params(c))) != c["md5"]]
    bad += [c["name"] for c in g["preview"]
            if md5(restatement.region(case_image(c), oracle_params(c), preview=True)) != c["md5"]]
    assert not bad, "%d u64 cases differ, first: %s" % (len(bad), bad[:5])
    assert len(g["suite"]) >= 150 and sum(c["extremes"] for c in g["suite"]) >= 50


def test_u64_white_wraps_to_black_in_the_compiled_reference(reference):
    """set_pixel's `roundl(d * 18446744073709551615UL)` is 2^64 for d = 1.0 -- out of range for uint64_t, undefined in
    C; the reference as compiled (gcc, x86-64) yields 0.  Pinned here because the CUDA path reproduces it."""
    if not reference.available:
        pytest.skip("oracle/_ref not built")
    img = np.full((6, 8, 3), np.iinfo(np.uint64).max, dtype=np.uint64)
    img[2, 3, 0] = 1 << 63
    img[3, 3, 0] = (1 << 64) - (1 << 11)        # decodes to 1 - 2^-53: the largest value that survives
    img[3, 4, 0] = (1 << 64) - (1 << 10) - 1    # rounds to 1.0 in get_pixel's second rounding
    out = reference.region(img, orc.Params(interpolation=1, lens_x=4, lens_y=3))
    assert out[0, 0, 0] == 0 and out[2, 3, 0] == 1 << 63 and out[3, 3, 0] == (1 << 64) - (1 << 11) and out[3, 4, 0] == 0
    assert (out[..., 1] == img[..., 1]).all()   # green is copied


# ---------------------------------------------------------------------------------------------
# u15 (bpc = 15): the rows the reference leaves "TODO for another day", written in by oracle/patch_u15.py
# ---------------------------------------------------------------------------------------------
def test_u15_restatement_matches_golden(restatement):
    g = golden_u15()
    bad = [c["name"] for c in g["suite"]
           if md5(restatement.region(case_image(c), oracle_params(c), bpc=orc.BPC_U15)) != c["md5"]]
    bad += [c["name"] for c in g["preview"]
            if md5(restatement.region(case_image(c), oracle_params(c), preview=True, bpc=orc.BPC_U15)) != c["md5"]]
    assert not bad, "%d u15 cases differ, first: %s" % (len(bad), bad[:5])


def test_u15_reference_build_matches_golden_and_leaves_other_formats_alone(reference):
    if not orc.ReferenceU15.available():
        pytest.skip("oracle/_ref/libfixca_ref_u15.so not built (needs /root/reference at build time)")
    u15 = orc.ReferenceU15()
    sample = golden_u15()["suite"][::5]
    bad = [c["name"] for c in sample if md5(u15.region(case_image(c), oracle_params(c), bpc=orc.BPC_U15)) != c["md5"]]
    assert not bad, bad[:5]
    # the written-in rows are the only difference: every other format computes what the unpatched build does
    for c in golden()["suite"][::97]:
        assert md5(u15.region(case_image(c), oracle_params(c))) == c["md5"], c["name"]
    assert u15.color_size("R'G'B' u15", 6) == 15 and reference.color_size("R'G'B' u15", 6) == -99
    assert u15.color_size("R'G'B' u16", 6) == 2 and u15.color_size("R'G'B' float", 12) == -4


def test_u15_is_the_u16_pass_at_another_scale(restatement):
    """The spec of the extension in one property: decode v / 32768 and encode round(d * 32768) are the u16
    branches with another maximum, so interpolation None moves the same 16-bit codes u16 does, and a constant
    image stays constant under Linear / Cubic (weights sum to one; 12345 / 32768 is exact)."""
    img = orc.synth_u15(40, 61, 3, seed=5, wide=True)
    p = orc.Params(blue=3.0, red=-2.0, lens_x=30, lens_y=20, interpolation=0, x_blue=0.7, y_red=-0.9)
    assert restatement.region(img, p, bpc=orc.BPC_U15).tobytes() == restatement.region(img, p).tobytes()
    flat = np.full((33, 47, 4), 12345, dtype=np.uint16)
    for interp in (1, 2):
        p.interpolation = interp
        assert (restatement.region(flat, p, bpc=orc.BPC_U15) == 12345).all()
