"""GPU: the CUDA path, called through the C ABI, against the oracle.

Bars (BASELINE.json north_star): interpolation None and the EXACT arithmetic are bit-exact for
every format; FAST (FP32) Linear/Cubic is within +-1 LSB per channel for u8/u16 and within
FLOAT_ABS_TOL for float32 images (tolerances written here)."""
import ctypes
import hashlib
import itertools
import os

import numpy as np
import pytest

import oracle as orc
from fixture_io import encode_gimp_bmp24
from helpers import golden_u64, golden_half, golden_u15, case_image, fixture_image, fx_params, golden, lsb_diff, max_dim, md5, oracle_params

pytestmark = pytest.mark.gpu

FAST_LSB_TOL = 1            # u8 / u16, FIXCA_PRECISION_FAST
FLOAT_ABS_TOL = 1.0e-6      # float32 images, FIXCA_PRECISION_FAST (outputs are clipped to [0,1])

KW = dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(fx):
    assert fx.device_count() > 0, "these tests need a CUDA device; the product has no CPU path"


# ---------------------------------------------------------------------------------------------
# golden digests (produced by the reference's own code, tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("force", ["auto", "direct", "tiled", "inline"])
def test_golden_suite_bit_exact(fx, force, tuning):
    """EXACT arithmetic through each of its kernel families: auto = the streaming kernel with exact repair of
    near-tie samples for 8-bit (f32+f64) and 16-bit (f64+exact) integers and tiled_kernel<ExactF64> for the rest; tiled = the FP64 tile
    kernel for every format (FIXCA_EXACT_KERNEL=tiled); direct = the per-pixel kernel; inline = 8-bit samples repaired
    inside the streaming kernel (FIXCA_EXACT_KERNEL=inline) instead of by repair_patch_kernel behind it."""
    if force in ("tiled", "inline"):
        tuning("FIXCA_EXACT_KERNEL", force)
    flags = fx.PRECISION_EXACT | (fx.FORCE_DIRECT if force == "direct" else 0)
    bad, kernels = [], set()
    for c in golden()["suite"]:
        got = fx.correct(case_image(c), fx_params(fx, c), flags=flags)
        k = fx.last_kernel().split("/")
        kernels.add(k[0] + ("/" + k[2] if k[1] != "none" else "/none"))
        if md5(got) != c["md5"]:
            bad.append(c["name"])
    assert not bad, "%d of %d golden cases differ (%s), first: %s" % (len(bad), len(golden()["suite"]), force, bad[:8])
    if force == "auto":
        assert {"stream/f32+f64", "stream/f64+exact", "tiled/f64"} <= kernels, kernels
    elif force == "inline":
        assert {"stream/f32+f64inline", "stream/f64+exact", "tiled/f64"} <= kernels and "stream/f32+f64" not in kernels, kernels
    elif force == "tiled":
        assert "tiled/f64" in kernels and not {"stream/f32+f64", "stream/f64+exact"} & kernels, kernels
    else:
        assert {k.split("/")[0] for k in kernels} == {"direct"}, kernels


def test_golden_suite_fast_within_tolerance(fx, checker):
    worst = {"u1": 0, "u2": 0, "f4": 0.0}
    for c in golden()["suite"]:
        if c["dtype"] not in worst or c["interpolation"] == 0:
            continue
        img = case_image(c)
        want = checker.region(img, oracle_params(c))
        got = fx.correct(img, fx_params(fx, c), flags=fx.PRECISION_FAST)
        d, _ = lsb_diff(got, want)
        worst[c["dtype"]] = max(worst[c["dtype"]], d)
        tol = FLOAT_ABS_TOL if c["dtype"] == "f4" else FAST_LSB_TOL
        assert d <= tol, (c["name"], d, fx.last_kernel())
    print("fast-mode worst |diff|:", worst)


# ---------------------------------------------------------------------------------------------
# the reference's known-answer test (tests/test1.md5) and its variants
# ---------------------------------------------------------------------------------------------
def test_known_answer_test1_md5(fx):
    img = fixture_image()
    if img is None:
        pytest.skip("oracle/_ref/full-branches.rgb missing (generated where /root/reference is mounted)")
    g = golden()["fixture"]
    out = fx.correct(img, fx.FixCaParams(blue=6.0, red=-2.4, lens_x=0, lens_y=0, interpolation=1))
    assert "stream/linear/f32+f64" in fx.last_kernel()
    assert md5(out) == g["test1_raw_md5"]
    assert hashlib.md5(encode_gimp_bmp24(out)).hexdigest() == "c472550cda23c8cb717853ac0dd93e2b"


def test_fixture_variants_bit_exact_and_fast(fx, checker):
    img = fixture_image()
    if img is None:
        pytest.skip("fixture missing")
    for interp, lens in itertools.product((0, 1, 2), ((0, 0), (658, 1280))):
        p = fx.FixCaParams(blue=6.0, red=-2.4, lens_x=lens[0], lens_y=lens[1], interpolation=interp)
        key = "i%d-lens%d,%d" % (interp, *lens)
        out = fx.correct(img, p)
        assert md5(out) == golden()["fixture"]["outputs"][key], key
        if interp:
            fast = fx.correct(img, p, flags=fx.PRECISION_FAST)
            d, frac = lsb_diff(fast, out)
            assert d <= FAST_LSB_TOL and frac < 1e-3, (key, d, frac)


# ---------------------------------------------------------------------------------------------
# seeded matrix against the checker: formats x interpolation x kernels, awkward shapes
# ---------------------------------------------------------------------------------------------
SHAPES = [(1, 1), (2, 3), (7, 129), (129, 7), (64, 128), (65, 257), (301, 517), (97, 1000)]


@pytest.mark.parametrize("dtype,ch", [("u1", 3), ("u1", 4), ("u2", 3), ("u2", 4), ("u4", 3), ("u4", 4),
                                      ("f4", 3), ("f4", 4), ("f8", 3), ("f8", 4), ("u8", 3), ("u8", 4)])
def test_matrix_exact(fx, checker, dtype, ch):
    n = 0
    for (h, w), interp, lens in itertools.product(SHAPES, (0, 1, 2), ("c", (0, 0), (-1, -1))):
        n += 1
        lx, ly = (w // 2, h // 2) if lens == "c" else lens
        kw = dict(KW, lens_x=lx, lens_y=ly, interpolation=interp)
        m = max_dim(w, h, lx, ly)
        if m + kw["blue"] == 0 or m + kw["red"] == 0:
            continue        # scale = inf: the reference reads out of bounds, the ABI returns ERR_DEGENERATE
        img = orc.synth_image(h, w, ch, dtype, seed=1000 + n, wide=bool(n % 2))
        want = checker.region(img, orc.Params(**kw))
        for flags in (fx.PRECISION_EXACT, fx.PRECISION_EXACT | fx.FORCE_DIRECT):
            got = fx.correct(img, fx.FixCaParams(**kw), flags=flags)
            assert got.tobytes() == want.tobytes(), (h, w, dtype, ch, interp, lens, fx.last_kernel())


FAST_SHAPES = [(1, 1), (2, 3), (5, 40), (7, 129), (129, 7), (64, 128), (65, 257), (33, 513), (301, 517), (97, 1000), (40, 2051)]


@pytest.mark.parametrize("dtype,ch", [("u1", 3), ("u1", 4), ("u2", 3), ("u2", 4), ("u4", 3), ("u4", 4)])
def test_matrix_none_stream(fx, checker, dtype, ch):
    """interpolation = None on the streaming kernel: every sample size x awkward shapes (strips narrower than
    a warp's column group, widths that are not a multiple of the strip, bands that end inside a chunk) x lens
    positions x scales on both sides of 1: identical bytes (fix-ca.c:1100-1121)."""
    n, kernels = 0, set()
    rng = np.random.default_rng(77)
    for (h, w), lens, amounts in itertools.product(FAST_SHAPES, ("c", (0, 0), (-1, -1)),
                                                   ((3.0, -2.0), (-6.0, 2.4), (0.0, 0.0), (30.0, -30.0))):
        n += 1
        lx, ly = (w // 2, h // 2) if lens == "c" else lens
        kw = dict(KW, blue=amounts[0], red=amounts[1], lens_x=lx, lens_y=ly, interpolation=0)
        m = max_dim(w, h, lx, ly)
        if m + kw["blue"] <= 0 or m + kw["red"] <= 0:
            continue
        img = rng.integers(0, np.iinfo(dtype).max, size=(h, w, ch), dtype=dtype, endpoint=True)
        want = checker.region(img, orc.Params(**kw))
        got = fx.correct(img, fx.FixCaParams(**kw))
        kernels.add(fx.last_kernel().split("/")[0])
        assert got.tobytes() == want.tobytes(), (h, w, dtype, ch, lens, amounts, fx.last_kernel())
        # a band that starts and ends inside chunks, into a poisoned buffer: only its rows change
        if h >= 6:
            y1, y2 = h // 3 + 1, max(h // 3 + 2, 2 * h // 3 - 1)
            out = np.full_like(img, 0x5A)
            fx.correct(img, fx.FixCaParams(**kw), y1=y1, y2=y2, out=out)
            assert out[y1:y2].tobytes() == want[y1:y2].tobytes() and (out[:y1] == 0x5A).all() and (out[y2:] == 0x5A).all()
    assert "stream" in kernels, kernels


@pytest.mark.parametrize("variant", ["stream", "strip"])
@pytest.mark.parametrize("dtype,ch", [("u1", 3), ("u1", 4), ("u2", 3), ("u2", 4), ("f4", 3), ("f4", 4)])
def test_matrix_fast(fx, checker, dtype, ch, variant, tuning):
    """FAST (FP32) arithmetic, streaming kernel and its per-tile fallback (strip): every format x Linear/Cubic x awkward shapes
    (tiles narrower than a warp's column group, widths that are not a multiple of the tile) x
    lens positions x scales on both sides of 1, against the checker within the stated tolerance."""
    tuning("FIXCA_FAST_KERNEL", variant)
    tol = FLOAT_ABS_TOL if dtype == "f4" else FAST_LSB_TOL
    n, kernels = 0, set()
    for (h, w), interp, lens, amounts in itertools.product(
            FAST_SHAPES, (1, 2), ("c", (0, 0), (-1, -1)), ((3.0, -2.0), (-6.0, 2.4), (0.0, 0.0), (30.0, -30.0))):
        n += 1
        lx, ly = (w // 2, h // 2) if lens == "c" else lens
        kw = dict(KW, blue=amounts[0], red=amounts[1], lens_x=lx, lens_y=ly, interpolation=interp)
        m = max_dim(w, h, lx, ly)
        if m + kw["blue"] <= 0 or m + kw["red"] <= 0:
            continue        # infinite / negative scale: covered by the direct-kernel tests
        img = orc.synth_image(h, w, ch, dtype, seed=5000 + n)
        want = checker.region(img, orc.Params(**kw))
        got = fx.correct(img, fx.FixCaParams(**kw), flags=fx.PRECISION_FAST)
        kernels.add(fx.last_kernel().split("/")[0])
        d, frac = lsb_diff(got, want)
        assert d <= tol, (h, w, dtype, ch, interp, lens, amounts, d, frac, fx.last_kernel())
        # pass-through channels are byte copies in every mode (fix-ca.c:1094-1098)
        assert np.array_equal(got[..., 1], img[..., 1])
        if ch == 4:
            assert np.array_equal(got[..., 3], img[..., 3])
    assert variant in kernels, kernels


@pytest.mark.parametrize("variant", ["stream", "tiled"])
def test_none_is_bit_exact_for_every_sample_size_including_nan_payloads(fx, checker, variant, tuning):
    tuning("FIXCA_NONE_KERNEL", variant)
    rng = np.random.default_rng(3)
    for dt, ch in itertools.product(("u1", "u2", "u4", "u8"), (3, 4)):
        img = rng.integers(0, np.iinfo(dt).max, size=(131, 259, ch), dtype=dt, endpoint=True)
        kw = dict(KW, lens_x=100, lens_y=60, interpolation=0)
        want = checker.region(img, orc.Params(**kw))
        assert fx.correct(img, fx.FixCaParams(**kw)).tobytes() == want.tobytes()
        # 8-byte samples have no streaming kernel (a strip of them exceeds a TMA box)
        assert fx.last_kernel().startswith("tiled" if dt == "u8" else variant), fx.last_kernel()
        # the same bits viewed as floats (NaN payloads, infinities) must pass through untouched
        if dt in ("u4", "u8"):
            fimg = img.view("f4" if dt == "u4" else "f8")
            got = fx.correct(fimg, fx.FixCaParams(**kw))
            assert got.tobytes() == want.tobytes()


def test_large_shifts_and_fallback_kernel(fx, checker):
    """+-30 lateral and +-30 opposite directional shifts: the union source window of a tile grows by
    ~120 px; whichever kernel the planner picks must stay exact."""
    img = orc.synth_image(400, 600, 3, "u2", 77)
    kw = dict(blue=30.0, red=-30.0, x_blue=30.0, x_red=-30.0, y_blue=-30.0, y_red=30.0, lens_x=300, lens_y=200)
    for interp in (0, 1, 2):
        want = checker.region(img, orc.Params(interpolation=interp, **kw))
        got = fx.correct(img, fx.FixCaParams(interpolation=interp, **kw))
        assert got.tobytes() == want.tobytes(), (interp, fx.last_kernel())


def test_negative_scale_uses_direct_kernel(fx, checker):
    # max_dim + amount < 0: scale negative, the map decreases; the reference still produces output
    img = orc.synth_image(20, 24, 3, "u1", 5)
    kw = dict(blue=-30.0, red=-29.0, lens_x=12, lens_y=10)
    for interp in (0, 1, 2):
        want = checker.region(img, orc.Params(interpolation=interp, **kw))
        got = fx.correct(img, fx.FixCaParams(interpolation=interp, **kw))
        assert got.tobytes() == want.tobytes() and fx.last_kernel().startswith("direct")


def test_clip_of_float_images(fx, checker):
    # clip_d clamps float/double images to [0,1] in Linear/Cubic but not in None (fix-ca.c:873-880)
    for dt in ("f4", "f8"):
        img = orc.synth_image(90, 140, 3, dt, 9, wide=True)
        for interp in (0, 1, 2):
            kw = dict(KW, lens_x=70, lens_y=45, interpolation=interp)
            want = checker.region(img, orc.Params(**kw))
            got = fx.correct(img, fx.FixCaParams(**kw))
            assert got.tobytes() == want.tobytes()
            if interp:
                assert got[..., 0].min() >= 0.0 and got[..., 0].max() <= 1.0 and got[..., 1].max() > 1.0


def test_zero_params_identity(fx):
    for dt, interp, flags in itertools.product(("u1", "u2", "u4", "f4", "f8"), (0, 1, 2), (0, 1)):
        img = orc.synth_image(150, 300, 3, dt, 21)
        got = fx.correct(img, fx.FixCaParams(interpolation=interp, lens_x=150, lens_y=75), flags=flags)
        assert (got == img).all(), (dt, interp, flags)


# ---------------------------------------------------------------------------------------------
# drop-in semantics of the region call
# ---------------------------------------------------------------------------------------------
def test_band_call_writes_only_its_rows(fx, checker):
    img = orc.synth_image(333, 411, 4, "u2", 31)
    kw = dict(KW, lens_x=200, lens_y=100, interpolation=2)
    full = checker.region(img, orc.Params(**kw))
    dst = np.full_like(img, 0xABCD)
    fx.fix_ca_region(img, dst, 411, 333, 8, 2, fx.FixCaParams(**kw), 0, 411, 100, 177, True)
    assert (dst[100:177] == full[100:177]).all()
    assert (dst[:100] == 0xABCD).all() and (dst[177:] == 0xABCD).all()
    # single row, first row, last row
    for y in (0, 166, 332):
        d2 = np.zeros_like(img)
        fx.fix_ca_region(img, d2, 411, 333, 8, 2, fx.FixCaParams(**kw), 0, 411, y, y + 1, True)
        assert (d2[y] == full[y]).all()


def test_column_selection_is_the_crop_of_the_full_width_pass(fx, checker):
    """FIXCA_COLUMN_SELECTION (extension): columns [x1,x2) of rows [y1,y2) get what the
    full-width pass computes for them -- the reference's per-pixel arithmetic (fix-ca.c:1105-1320 never uses
    x1 / x2 in an index), checked against the reference run full width -- and nothing else is written.
    Without the flag the call is refused like before (the reference's own x1 != 0 path is broken)."""
    import torch

    for (h, w, ch, dt, interp, flags) in ((333, 411, 4, "u2", 2, fx.PRECISION_EXACT), (257, 300, 3, "u1", 1, fx.PRECISION_EXACT),
                                          (129, 200, 3, "f4", 0, fx.PRECISION_EXACT), (1400, 2100, 3, "u1", 2, fx.PRECISION_EXACT)):
        img = orc.synth_image(h, w, ch, dt, 77)
        kw = dict(KW, lens_x=w // 2 - 7, lens_y=h // 3, interpolation=interp)
        full = checker.region(img, orc.Params(**kw))
        bpp = ch * img.dtype.itemsize
        for (x1, x2, y1, y2) in ((5, w - 9, 10, h - 20), (0, 17, 0, h), (w - 1, w, 3, 4), (100, 101, 0, h), (0, w, 7, 90)):
            # pageable caller (numpy; the 8.8 MB case goes through the staging rings) and pinned caller
            for pinned in (False, True):
                if pinned:
                    t = torch.empty(img.nbytes, dtype=torch.uint8, pin_memory=True)
                    dst = t.numpy().view(img.dtype).reshape(img.shape)
                else:
                    dst = np.empty_like(img)
                dst.view(np.uint8)[:] = 0xA5
                before = dst.copy()
                fx.fix_ca_region(img, dst, w, h, bpp, fx.bpc_of(img.dtype), fx.FixCaParams(**kw), x1, x2, y1, y2, True,
                                 flags | fx.COLUMN_SELECTION)
                assert dst[y1:y2, x1:x2].tobytes() == full[y1:y2, x1:x2].tobytes(), (h, w, x1, x2, y1, y2, pinned)
                dst[y1:y2, x1:x2] = before[y1:y2, x1:x2]
                assert dst.tobytes() == before.tobytes(), "bytes outside the selection were written"
    img = orc.synth_image(16, 16, 3, "u1", 1)
    for (x1, x2, flags) in ((1, 16, fx.PRECISION_EXACT), (4, 4, fx.COLUMN_SELECTION), (-1, 8, fx.COLUMN_SELECTION),
                            (0, 17, fx.COLUMN_SELECTION)):
        with pytest.raises(fx.FixCaError) as e:
            fx.fix_ca_region(img, np.zeros_like(img), 16, 16, 3, 1, fx.FixCaParams(), x1, x2, 0, 16, True, flags)
        assert e.value.code == fx.ERR_REGION


def test_progress_protocol_matches_reference(fx):
    # fix-ca.c:1022-1023, 1331-1332, 1335-1336
    img = orc.synth_image(50, 64, 3, "u1", 2)
    calls = []
    fx.set_progress(lambda kind, frac: calls.append((kind, frac)))
    try:
        fx.correct(img, fx.FixCaParams(blue=1.0, lens_x=32, lens_y=25), y1=3, y2=45)
    finally:
        fx.set_progress(None)
    assert calls[0] == (0, 0.0)
    want = [(1, (y - 3) / 42.0) for y in range(3, 45) if (y - 3) % 8 == 0] + [(1, 0.0)]
    assert calls[1:] == want


def test_multi_device_bands_and_frames(fx, checker):
    img = orc.synth_image(500, 700, 3, "u2", 41)
    kw = dict(KW, lens_x=350, lens_y=250, interpolation=2)
    want = checker.region(img, orc.Params(**kw))
    ndev = fx.device_count()
    # several bands on whatever devices exist (the same device may serve more than one band)
    for bands in (1, 2, 3):
        devs = [i % ndev for i in range(bands)]
        got = fx.correct(img, fx.FixCaParams(**kw), devices=devs)
        assert got.tobytes() == want.tobytes(), bands
    frames = [orc.synth_image(120, 200, 3, "u1", 100 + k) for k in range(7)]
    outs = fx.correct_frames(frames, fx.FixCaParams(**kw), flags=fx.PRECISION_EXACT)
    for f, o in zip(frames, outs):
        assert o.tobytes() == checker.region(f, orc.Params(**kw)).tobytes()
    # the same stream sharded by index over devices from one process (a device may serve several shards)
    for nd in (1, 2, 3):
        outs = fx.correct_frames(frames, fx.FixCaParams(**kw), flags=fx.PRECISION_EXACT, devices=[i % ndev for i in range(nd)])
        for f, o in zip(frames, outs):
            assert o.tobytes() == checker.region(f, orc.Params(**kw)).tobytes(), nd


def test_fanout_stores_every_destination(fx, checker):
    """fixca_cuda_region_dev_fanout: one launch stores the band into several frames (the all-gather form; here all
    frames live on this GPU).  Streaming kernels (FAST, None, EXACT on 8-/16-bit samples) fan out in ONE launch; the FP64 tile kernel goes launch by launch.
    Every frame holds the single-destination result, and nothing outside the band's rows is written."""
    import torch

    h, w, ch = 403, 640, 3
    st = torch.cuda.current_stream().cuda_stream
    for dt, interp, flags, streaming in (("u2", 2, fx.PRECISION_FAST, True), ("u1", 0, fx.PRECISION_EXACT, True),
                                         ("f4", 1, fx.PRECISION_FAST, True), ("u2", 2, fx.PRECISION_EXACT, True),
                                         ("u1", 2, fx.PRECISION_EXACT, True), ("f4", 2, fx.PRECISION_EXACT, False)):
        img = orc.synth_image(h, w, ch, dt, 91)
        kw = dict(KW, lens_x=300, lens_y=128, interpolation=interp)
        p = fx.FixCaParams(**kw)
        want = fx.correct(img, p, flags=flags)
        bpp = ch * img.dtype.itemsize
        pitch = (w * bpp + 127) // 128 * 128
        src = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
        src[:, :w * bpp] = torch.from_numpy(img.view(np.uint8).reshape(h, w * bpp)).cuda()
        y1, y2 = 96, 301
        for ndst in (1, 3, 8):
            frames = [torch.full((h, pitch), 0x5A, dtype=torch.uint8, device="cuda") for _ in range(ndst)]
            n0 = fx.launch_count()
            fx.fix_ca_region_dev_fanout(src.data_ptr(), pitch, 0, h, [f.data_ptr() for f in frames], pitch, 0, w, h, bpp,
                                        fx.bpc_of(img.dtype), p, y1, y2, flags, st)
            torch.cuda.synchronize()
            # (8-bit EXACT: the streaming kernel + repair_patch_kernel behind it, which patches every destination)
            per_call = 2 if fx.last_kernel().split("/")[2] == "f32+f64" else 1
            assert fx.launch_count() - n0 == (per_call if streaming else ndst), (dt, interp, ndst, fx.last_kernel())
            for f in frames:
                got = f[:, :w * bpp].cpu().numpy()
                assert got[y1:y2].tobytes() == want[y1:y2].tobytes(), (dt, interp, ndst)
                assert (got[:y1] == 0x5A).all() and (got[y2:] == 0x5A).all()
    with pytest.raises(fx.FixCaError):
        fx.fix_ca_region_dev_fanout(src.data_ptr(), pitch, 0, h, [src.data_ptr()] * 9, pitch, 0, w, h, bpp, -4, p, 0, h, flags, st)


def test_exact8_concurrent_streams_and_many_buffers(fx, checker):
    """8-bit EXACT is two launches joined by a queue in device memory (stream kernel -> repair_patch_kernel).  The queue is
    kept per calling thread, device and stream: launches of the same geometry on two streams at once, launches that
    alternate between a large and a small image on one stream (the queue grows), and one geometry over many buffer pairs
    all return the reference's bytes."""
    import torch

    kw = dict(KW, lens_x=300, lens_y=128, interpolation=2)
    p = fx.FixCaParams(**kw)
    sizes = ((257, 640, 3), (97, 320, 4))
    imgs = {sz: [orc.synth_image(sz[0], sz[1], sz[2], "u1", 500 + 10 * i + k) for k in range(6)] for i, sz in enumerate(sizes)}
    want = {sz: [checker.region(im, orc.Params(**kw)) for im in imgs[sz]] for sz in sizes}
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    jobs = []
    for rnd in range(3):
        for k in range(6):
            for si, sz in enumerate(sizes):      # the two sizes alternate on both streams
                h, w, ch = sz
                st = streams[(k + si + rnd) % 2]
                with torch.cuda.stream(st):
                    src = torch.from_numpy(imgs[sz][k]).cuda()
                    dst = torch.full_like(src, 0x5A)
                    fx.fix_ca_region_dev(src.data_ptr(), w * ch, 0, h, dst.data_ptr(), w * ch, 0, w, h, ch, 1, p, 0, h,
                                         fx.PRECISION_EXACT, st.cuda_stream)
                assert fx.last_kernel().split("/")[2] == "f32+f64", fx.last_kernel()
                jobs.append((sz, k, src, dst))
    torch.cuda.synchronize()
    for sz, k, _src, dst in jobs:
        assert dst.cpu().numpy().tobytes() == want[sz][k].tobytes(), (sz, k)


def test_device_resident_entry_with_torch_buffers(fx, checker):
    import torch

    h, w = 257, 640                       # 640*6 bytes: pitch is a multiple of 16 -> tiled kernel
    img = orc.synth_image(h, w, 3, "u2", 51)
    kw = dict(KW, lens_x=300, lens_y=128, interpolation=2)
    want = checker.region(img, orc.Params(**kw))
    src = torch.from_numpy(img.view(np.int16)).cuda()
    dst = torch.zeros_like(src)
    stream = torch.cuda.current_stream().cuda_stream
    fx.fix_ca_region_dev(src.data_ptr(), w * 6, 0, h, dst.data_ptr(), w * 6, 0, w, h, 6, 2, fx.FixCaParams(**kw), 0, h,
                         fx.PRECISION_EXACT, stream)
    torch.cuda.synchronize()
    assert fx.last_kernel().startswith("stream/cubic/f64+exact")
    assert dst.cpu().numpy().view(np.uint16).tobytes() == want.tobytes()
    # a band whose source rows are a sub-range of the image (what one rank of a multi-GPU run holds)
    lo, hi = fx.band_source_rows(w, h, fx.FixCaParams(**kw), 100, 180)
    sub = src[lo:hi + 1].contiguous()
    out = torch.zeros((80, w, 3), dtype=torch.int16, device="cuda")
    fx.fix_ca_region_dev(sub.data_ptr(), w * 6, lo, hi - lo + 1, out.data_ptr(), w * 6, 100, w, h, 6, 2,
                         fx.FixCaParams(**kw), 100, 180, fx.PRECISION_EXACT, stream)
    torch.cuda.synchronize()
    assert out.cpu().numpy().view(np.uint16).tobytes() == want[100:180].tobytes()
    # missing halo rows are refused, not read out of bounds
    with pytest.raises(fx.FixCaError):
        fx.fix_ca_region_dev(sub.data_ptr(), w * 6, lo + 1, hi - lo, out.data_ptr(), w * 6, 100, w, h, 6, 2,
                             fx.FixCaParams(**kw), 100, 180, fx.PRECISION_EXACT, stream)
    # odd width: rows are not 16-byte multiples -> the direct kernel takes over, still exact
    h2, w2 = 100, 333
    img2 = orc.synth_image(h2, w2, 3, "u1", 52)
    want2 = checker.region(img2, orc.Params(**kw))
    s2 = torch.from_numpy(img2).cuda()
    d2 = torch.zeros_like(s2)
    fx.fix_ca_region_dev(s2.data_ptr(), w2 * 3, 0, h2, d2.data_ptr(), w2 * 3, 0, w2, h2, 3, 1, fx.FixCaParams(**kw), 0, h2,
                         fx.PRECISION_EXACT, stream)
    torch.cuda.synchronize()
    assert fx.last_kernel().startswith("direct") and d2.cpu().numpy().tobytes() == want2.tobytes()


def test_pinned_caller_chunk_ramp_changes_no_byte(fx, checker, tuning):
    """Pinned callers (the patched plug-in's buffers, bench.py's e2e): the staging chunks ramp up at the start of a
    band and down at its end (region_host_band_locked); the progress protocol and every byte are those of the
    uniform chunking, and both are the reference's."""
    h, w = 4000, 3000
    src_p, dst_p = fx.PinnedBuffer(h * w * 6), fx.PinnedBuffer(h * w * 6)
    try:
        img = src_p.array(np.uint16, (h, w, 3))
        img[...] = orc.synth_image(h, w, 3, "u2", 4242)
        out = dst_p.array(np.uint16, (h, w, 3))
        kw = dict(KW, lens_x=w // 2, lens_y=h // 2, interpolation=2)
        digests = {}
        for ramp in ("1", "0"):
            tuning("FIXCA_CHUNK_MB", "8")
            tuning("FIXCA_CHUNK_RAMP", ramp)
            out[...] = 0
            n0 = fx.launch_count()
            fx.fix_ca_region(img, out, w, h, 6, 2, fx.FixCaParams(**kw), 0, w, 0, h, True, fx.PRECISION_EXACT)
            digests[ramp] = (md5(out), fx.launch_count() - n0)
        assert digests["1"][0] == digests["0"][0]
        assert digests["1"][1] > digests["0"][1]            # more (smaller) chunks at the two ends
        for y1, y2 in ((0, 40), (1990, 2030), (3960, 4000)):
            want = checker.region(img, orc.Params(**kw), y1=y1, y2=y2)
            assert (out[y1:y2] == want[y1:y2]).all(), (y1, y2)
    finally:
        src_p.free()
        dst_p.free()


def test_plan_tables_survive_cache_churn_and_threads(fx, checker):
    """The streaming kernels read per-plan chunk / column tables from device memory (stream_meta_kernel,
    stream_cols_kernel); plans and tables sit in per-thread caches of 96 slots.  More distinct bands, widths and
    modes than there are slots: evicted tables are freed only with the last plan that holds them, every result
    stays identical to the reference's; then the same from four threads at once (per-thread caches, one device)."""
    import threading

    import torch

    h, w = 301, 512
    img = orc.synth_image(h, w, 3, "u1", 77)
    kwc = dict(KW, lens_x=250, lens_y=140, interpolation=2)
    want = {i: checker.region(img, orc.Params(**dict(kwc, interpolation=i))) for i in (0, 1, 2)}
    src = torch.from_numpy(img).cuda()
    stream = torch.cuda.current_stream().cuda_stream

    def bands(seed, n):
        rng = np.random.default_rng(seed)
        out = torch.empty_like(src)
        for k in range(n):
            interp = int(rng.integers(0, 3))
            y1 = int(rng.integers(0, h - 1))
            y2 = int(rng.integers(y1 + 1, h + 1))
            out.fill_(0x5A)
            fx.fix_ca_region_dev(src.data_ptr(), w * 3, 0, h, out.data_ptr(), w * 3, 0, w, h, 3, 1,
                                 fx.FixCaParams(**dict(kwc, interpolation=interp)), y1, y2, fx.PRECISION_EXACT, stream)
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            assert fx.last_kernel().startswith("stream"), fx.last_kernel()
            assert (got[y1:y2] == want[interp][y1:y2]).all(), (seed, k, interp, y1, y2)
            assert (got[:y1] == 0x5A).all() and (got[y2:] == 0x5A).all()

    bands(1, 260)           # > 2 x 96 distinct plans: every slot of both caches is recycled
    # widths: one column table each
    for w2 in range(256, 256 + 16 * 110, 16):
        img2 = orc.synth_image(9, w2, 3, "u1", w2)
        s2 = torch.from_numpy(img2).cuda()
        d2 = torch.empty_like(s2)
        kw2 = dict(KW, lens_x=w2 // 2, lens_y=4, interpolation=1)
        fx.fix_ca_region_dev(s2.data_ptr(), w2 * 3, 0, 9, d2.data_ptr(), w2 * 3, 0, w2, 9, 3, 1, fx.FixCaParams(**kw2), 0, 9,
                             fx.PRECISION_EXACT, stream)
        torch.cuda.synchronize()
        if w2 % 256 == 0:
            assert d2.cpu().numpy().tobytes() == checker.region(img2, orc.Params(**kw2)).tobytes(), w2
    errors = []

    def worker(seed):
        try:
            torch.cuda.set_device(0)
            bands(seed, 40)
        except BaseException as e:      # noqa: BLE001 -- reported on the main thread
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=worker, args=(100 + i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_rows_that_are_views_into_a_wider_buffer_are_not_overrun(fx, checker):
    """Device-resident destination rows that are a sub-rectangle of a wider buffer (pitch > row, width * bytes not a
    multiple of 16): by default nothing past width * bytes of a row is written (the per-pixel kernel takes the call);
    with FIXCA_PADDING_SCRATCH the TMA kernels run and may use the bytes up to the next 16-byte boundary."""
    import torch
    stream = torch.cuda.current_stream().cuda_stream
    h, w, ch = 90, 333, 3                  # 999 bytes per row
    img = orc.synth_image(h, w, ch, "u1", 61)
    kw = dict(KW, lens_x=160, lens_y=40, interpolation=2)
    want = checker.region(img, orc.Params(**kw))
    pitch = 1024 + 512                     # the image sits inside a wider buffer
    src = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
    src[:, :w * ch] = torch.from_numpy(img.reshape(h, w * ch)).cuda()
    for flags, kernel in ((fx.PRECISION_FAST, "direct"), (fx.PRECISION_FAST | fx.PADDING_SCRATCH, "stream"),
                          (fx.PRECISION_EXACT, "direct"), (fx.PRECISION_EXACT | fx.PADDING_SCRATCH, "stream/cubic/f32+f64")):
        dst = torch.full((h, pitch), 0x5A, dtype=torch.uint8, device="cuda")
        fx.fix_ca_region_dev(src.data_ptr(), pitch, 0, h, dst.data_ptr(), pitch, 0, w, h, ch, 1, fx.FixCaParams(**kw), 0, h, flags, stream)
        torch.cuda.synchronize()
        assert fx.last_kernel().startswith(kernel), (flags, fx.last_kernel())
        got = dst.cpu().numpy()
        d = np.abs(got[:, :w * ch].reshape(h, w, ch).astype(np.int64) - want.astype(np.int64)).max()
        assert d <= (FAST_LSB_TOL if flags & fx.PRECISION_FAST else 0)
        keep = w * ch if kernel == "direct" else (w * ch + 15) // 16 * 16
        assert (got[:, keep:] == 0x5A).all(), (flags, kernel)


def test_device_batch_of_frames_matches_single_frames(fx, checker):
    """fixca_cuda_frames_dev: a batch in one launch (grid layer per frame, long segments) gives every frame the
    bytes of its own single-frame call; formats without a streaming kernel are looped per frame."""
    import torch
    stream = torch.cuda.current_stream().cuda_stream
    launches = fx.load().fixca_cuda_launch_count
    for (h, w, ch, dt, interp, flags, nf) in ((270, 517, 3, "u1", 2, fx.PRECISION_FAST, 5), (64, 300, 4, "u2", 1, fx.PRECISION_FAST, 3),
                                              (130, 259, 3, "u2", 0, fx.PRECISION_EXACT, 4), (33, 100, 3, "f4", 2, fx.PRECISION_FAST, 2),
                                              (40, 64, 3, "u2", 2, fx.PRECISION_EXACT, 3), (270, 517, 3, "u1", 2, fx.PRECISION_EXACT, 5),
                                              (97, 640, 4, "u1", 1, fx.PRECISION_EXACT, 3)):
        kw = dict(KW, lens_x=w // 2 - 3, lens_y=h // 2 + 5, interpolation=interp)
        p = fx.FixCaParams(**kw)
        frames = [orc.synth_image(h, w, ch, dt, seed=300 + k) for k in range(nf)]
        es = np.dtype(dt).itemsize
        bpp, bpc = ch * es, (-es if dt.startswith("f") else es)
        pitch = (w * bpp + 127) // 128 * 128
        fstride = pitch * h + 256       # frames need not be back to back
        src = torch.zeros(nf * fstride, dtype=torch.uint8, device="cuda")
        dst = torch.full((nf * fstride,), 0x5A, dtype=torch.uint8, device="cuda")
        for k, fr in enumerate(frames):
            v = src[k * fstride:k * fstride + pitch * h].view(h, pitch)
            v[:, :w * bpp] = torch.from_numpy(fr.view(np.uint8).reshape(h, w * bpp)).cuda()
        n0 = launches()
        fx.fix_ca_frames_dev(src.data_ptr(), pitch, fstride, dst.data_ptr(), pitch, fstride, nf, w, h, bpp, bpc, p,
                             flags | fx.PADDING_SCRATCH, stream)     # rows are pitched: their padding is scratch
        torch.cuda.synchronize()
        streaming = fx.last_kernel().startswith("stream")
        per_call = 2 if fx.last_kernel().split("/")[2] == "f32+f64" else 1     # (+ repair_patch_kernel)
        assert launches() - n0 == (per_call if streaming else nf), (fx.last_kernel(), launches() - n0)
        for k, fr in enumerate(frames):
            got = dst[k * fstride:k * fstride + pitch * h].view(h, pitch)[:, :w * bpp].cpu().numpy().view(fr.dtype).reshape(h, w, ch)
            single = fx.correct(fr, p, flags=flags)
            assert got.tobytes() == single.tobytes(), (dt, ch, interp, k, fx.last_kernel())
        # the gaps between frames are not written
        gap = dst[pitch * h:fstride].cpu().numpy()
        assert (gap == 0x5A).all()


# FIXCA_FUZZ_SEEDS="100-160" runs more seeds than the four of the regular suite (a soak run on a GPU box)
def _fuzz_seeds():
    spec = os.environ.get("FIXCA_FUZZ_SEEDS", "")
    if "-" in spec:
        lo, hi = spec.split("-")
        return list(range(int(lo), int(hi)))
    return [11, 12, 13, 14]


@pytest.mark.parametrize("seed", _fuzz_seeds())
def test_fuzz_small_images_all_modes(fx, checker, seed):
    """Seeded random shapes (1..700 px, strips and chunks cut anywhere), formats, lens positions (inside, on the
    border, outside, the -1 reset value), lateral amounts and directional shifts over the plug-in's full +-30
    range, random row bands: None and EXACT must equal the reference's bytes, FAST must stay within tolerance,
    the pass-through channels must be copies, and rows outside the band must stay untouched."""
    rng = np.random.default_rng(seed)
    kernels = set()
    half = orc.half_checker()       # float16 images: the reference with its commented-out half lines enabled
    for it in range(70):
        h = int(rng.choice([1, 2, 3, 5, 9, 17, 40, 97, 130, 260, 517]))
        w = int(rng.choice([1, 2, 4, 7, 31, 64, 129, 255, 256, 257, 400, 700]))
        dt = str(rng.choice(["u1", "u2", "f4", "f2", "u1", "u2", "f4", "f2", "u4", "u8", "f8"]))       # (u32 / u64 / double: FAST computes EXACT)
        ch = int(rng.choice([3, 4]))
        interp = int(rng.integers(0, 3))
        lens = [(w // 2, h // 2), (0, 0), (-1, -1), (w - 1, h - 1), (w + 13, -7), (3, h + 40)][int(rng.integers(0, 6))]
        amt = [float(rng.uniform(-30, 30)) if rng.random() < 0.5 else float(rng.uniform(-4, 4)) for _ in range(2)]
        sh = [float(rng.uniform(-30, 30)) if rng.random() < 0.3 else float(rng.uniform(-2, 2)) for _ in range(4)]
        kw = dict(blue=amt[0], red=amt[1], x_blue=sh[0], x_red=sh[1], y_blue=sh[2], y_red=sh[3],
                  lens_x=lens[0], lens_y=lens[1], interpolation=interp)
        m = max_dim(w, h, lens[0], lens[1])
        if m + amt[0] <= 0.5 or m + amt[1] <= 0.5:
            continue        # degenerate / negative scale: the direct-kernel tests
        img = orc.synth_image(h, w, ch, dt, seed=int(rng.integers(1 << 30)))
        want = (half if dt == "f2" else checker).region(img, orc.Params(**kw))
        y1 = int(rng.integers(0, h))
        y2 = int(rng.integers(y1 + 1, h + 1))
        for flags in ((fx.PRECISION_EXACT,) if interp == 0 else (fx.PRECISION_EXACT, fx.PRECISION_FAST)):
            out = np.full_like(img, 0x5A) if dt[0] != "f" else np.full_like(img, 7.0)
            fx.correct(img, fx.FixCaParams(**kw), y1=y1, y2=y2, out=out, flags=flags)
            kernels.add(fx.last_kernel().split("/")[0] + "/" + fx.last_kernel().split("/")[1])
            ctx = (seed, it, h, w, dt, ch, kw, y1, y2, flags, fx.last_kernel())
            if flags == fx.PRECISION_EXACT:
                assert out[y1:y2].tobytes() == want[y1:y2].tobytes(), ctx
            else:
                d, _ = lsb_diff(out[y1:y2], want[y1:y2])
                assert d <= (FLOAT_ABS_TOL if dt in ("f4", "f8") else HALF_ABS_TOL if dt == "f2" else FAST_LSB_TOL), ctx + (d,)
                assert np.array_equal(out[y1:y2, :, 1], img[y1:y2, :, 1]), ctx
            fill = 7.0 if dt[0] == "f" else 0x5A
            assert (out[:y1] == fill).all() and (out[y2:] == fill).all(), ctx
    assert {"stream/none", "stream/linear", "stream/cubic"} <= kernels, kernels


# ---------------------------------------------------------------------------------------------
# half precision (bpc = -2), SURVEY.md 8(f) #4: checker = the reference with its commented-out half lines enabled
# ---------------------------------------------------------------------------------------------
HALF_ABS_TOL = 2.0 ** -11       # one unit in the last place of a half in [0.5, 1): FAST rounds an FP32 result once


def test_half_golden_suite_bit_exact(fx):
    """None and EXACT Linear / Cubic on float16 images: identical bytes (266 digests + 36 preview digests)."""
    g = golden_half()
    bad, kernels = [], set()
    for c in g["suite"]:
        got = fx.correct(case_image(c), fx_params(fx, c), flags=fx.PRECISION_EXACT)
        kernels.add(fx.last_kernel().rsplit("/", 1)[0])
        if md5(got) != c["md5"]:
            bad.append(c["name"])
    for c in g["preview"]:
        got = fx.correct(case_image(c), fx_params(fx, c), flags=fx.PRECISION_EXACT | fx.PREVIEW_OVERLAY)
        if md5(got) != c["md5"]:
            bad.append(c["name"])
    assert not bad, "%d half cases differ, first: %s" % (len(bad), bad[:8])
    assert {"stream/none/copy", "tiled/cubic/f64", "tiled/linear/f64"} <= kernels, kernels


def test_half_fast_within_one_ulp(fx):
    """FAST (FP32) Linear / Cubic on float16: the streaming kernel, at most one half-ulp step from the checker."""
    chk = orc.half_checker()
    worst, n, nbad, kernels = 0.0, 0, 0, set()
    for (h, w), ch, interp, wide in itertools.product(((65, 257), (301, 517), (40, 2051), (7, 129)), (3, 4), (1, 2), (False, True)):
        kw = dict(KW, lens_x=w // 2, lens_y=h // 2, interpolation=interp)
        img = orc.synth_image(h, w, ch, "f2", seed=h + w + ch + interp, wide=wide)
        want = chk.region(img, orc.Params(**kw))
        got = fx.correct(img, fx.FixCaParams(**kw), flags=fx.PRECISION_FAST)
        kernels.add(fx.last_kernel().split("/")[0])
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        worst = max(worst, float(d.max()))
        n += d.size
        nbad += int((d != 0).sum())
        assert np.array_equal(got[..., 1], img[..., 1])          # pass-through is a copy, clipped or not
    assert worst <= HALF_ABS_TOL, worst
    assert nbad / n < 2e-3, nbad / n                              # FP32 vs FP64 before one rounding to 11 bits
    assert kernels == {"stream"}, kernels


def test_float_pitch_padding_is_never_sampled(fx, checker):
    """FAST float kernels weigh out-of-image samples with 0, and 0 * NaN is NaN: the bytes between width * bpp
    and the 16-byte row end (caller's pitch padding, not zero-filled by the TMA unit) must not be read.
    Device-resident call with the padding of every row set to NaN bit patterns, widths around strip ends."""
    import torch
    stream = torch.cuda.current_stream().cuda_stream
    for (h, w, ch), interp in itertools.product(((129, 7, 3), (40, 131, 3), (33, 517, 4), (70, 257, 3)), (1, 2)):
        kw = dict(KW, lens_x=w // 2, lens_y=h // 2, interpolation=interp)
        img = orc.synth_image(h, w, ch, "f4", seed=900 + w)
        want = checker.region(img, orc.Params(**kw))
        bpp, pitch = ch * 4, (w * ch * 4 + 127) // 128 * 128
        src = torch.full((h, pitch), 0xFF, dtype=torch.uint8, device="cuda")          # 0xFFFFFFFF = NaN
        src[:, :w * bpp] = torch.from_numpy(img.view(np.uint8).reshape(h, w * bpp)).cuda()
        dst = torch.zeros_like(src)
        fx.fix_ca_region_dev(src.data_ptr(), pitch, 0, h, dst.data_ptr(), pitch, 0, w, h, bpp, -4,
                             fx.FixCaParams(**kw), 0, h, fx.PRECISION_FAST | fx.PADDING_SCRATCH, stream)
        torch.cuda.synchronize()
        got = dst[:, :w * bpp].cpu().numpy().view(np.float32).reshape(h, w, ch)
        assert fx.last_kernel().startswith("stream")
        assert np.isfinite(got).all()
        assert np.abs(got.astype(np.float64) - want).max() <= FLOAT_ABS_TOL, (h, w, ch, interp)


# ---------------------------------------------------------------------------------------------
# u64 Linear / Cubic (fix-ca.c:728-733, :759-761): the x87 long double steps restated in integer arithmetic
# ---------------------------------------------------------------------------------------------
def test_u64_golden_suite_bit_exact(fx):
    """EXACT Linear / Cubic and the preview overlay (saturation boost included) on u64 images, half of the samples at
    the rounding boundaries of get_pixel's two roundings and at the values set_pixel wraps: identical bytes to the
    compiled reference (198 digests + 36 preview digests), through the tiled and the direct kernel."""
    g = golden_u64()
    bad, kernels = [], set()
    for c in g["suite"]:
        for flags in (fx.PRECISION_EXACT, fx.PRECISION_EXACT | fx.FORCE_DIRECT, fx.PRECISION_FAST):    # FAST: u64 computes EXACT
            got = fx.correct(case_image(c), fx_params(fx, c), flags=flags)
            kernels.add(fx.last_kernel().rsplit("/", 1)[0])
            if md5(got) != c["md5"]:
                bad.append((c["name"], flags))
    for c in g["preview"]:
        got = fx.correct(case_image(c), fx_params(fx, c), flags=fx.PRECISION_EXACT | fx.PREVIEW_OVERLAY)
        if md5(got) != c["md5"]:
            bad.append((c["name"], "preview"))
    assert not bad, "%d u64 cases differ, first: %s" % (len(bad), bad[:8])
    assert {"tiled/cubic/f64", "tiled/linear/f64", "direct/cubic/f64", "direct/linear/f64"} <= kernels, kernels


# ---------------------------------------------------------------------------------------------
# u15 (bpc = 15), SURVEY.md 8(f) #4: checker = the reference with the u15 rows written in (oracle/patch_u15.py)
# ---------------------------------------------------------------------------------------------
def test_u15_golden_suite_bit_exact(fx):
    """None and EXACT Linear / Cubic on u15 images: identical bytes (266 digests + 36 preview digests)."""
    g = golden_u15()
    bad, kernels = [], set()
    for c in g["suite"]:
        got = fx.correct(case_image(c), fx_params(fx, c), flags=fx.PRECISION_EXACT, bpc=fx.BPC_U15)
        kernels.add(fx.last_kernel().rsplit("/", 1)[0])
        if md5(got) != c["md5"]:
            bad.append(c["name"])
    for c in g["preview"]:
        got = fx.correct(case_image(c), fx_params(fx, c), flags=fx.PRECISION_EXACT | fx.PREVIEW_OVERLAY, bpc=fx.BPC_U15)
        if md5(got) != c["md5"]:
            bad.append(c["name"])
    assert not bad, "%d u15 cases differ, first: %s" % (len(bad), bad[:8])
    assert {"stream/none/copy", "stream/cubic/f64+exact", "stream/linear/f64+exact"} <= kernels, kernels


@pytest.mark.parametrize("variant", ["stream", "strip"])
def test_u15_fast_within_one_lsb(fx, variant, tuning):
    """FAST (FP32) Linear / Cubic on u15 (streaming kernel and its per-tile fallback): within one code of the checker,
    in-range and out-of-range (> 32768, clipped) inputs; pass-through channels are copies."""
    tuning("FIXCA_FAST_KERNEL", variant)
    chk = orc.u15_checker()
    worst, n, nbad, kernels = 0, 0, 0, set()
    for (h, w), ch, interp, wide in itertools.product(((65, 257), (301, 517), (40, 2051), (7, 129)), (3, 4), (1, 2), (False, True)):
        kw = dict(KW, lens_x=w // 2, lens_y=h // 2, interpolation=interp)
        img = orc.synth_u15(h, w, ch, seed=h + w + ch + interp, wide=wide)
        want = chk.region(img, orc.Params(**kw), bpc=orc.BPC_U15)
        got = fx.correct(img, fx.FixCaParams(**kw), flags=fx.PRECISION_FAST, bpc=fx.BPC_U15)
        kernels.add(fx.last_kernel().split("/")[0])
        d = np.abs(got.astype(np.int64) - want.astype(np.int64))
        worst = max(worst, int(d.max()))
        n += d.size
        nbad += int((d != 0).sum())
        assert np.array_equal(got[..., 1], img[..., 1])
    assert worst <= FAST_LSB_TOL, worst
    assert nbad / n < 5e-3, nbad / n
    assert kernels == {variant}, kernels


@pytest.mark.parametrize("form", ["default", "inline"])
def test_exact_repair_kernels_on_exact_ties(fx, checker, tuning, form):
    """The exact-repair stream kernels (8-bit: FP32 + FP64 repair; 16-bit / u15: FP64 separable + reference-order
    repair) decide every sample that is NOT near a rounding boundary in their fast arithmetic; here most samples ARE on
    one: pure half- and quarter-pixel directional shifts make the Linear / Cubic weights dyadic (1/2, 9/16, 1/16 ...),
    so a large share of the results are exact .5 ties, whose rounding depends on the reference's own operation order.
    Identical bytes to the checker, and to the FP64 tile kernel.  8-bit samples, default form: far more samples are
    flagged than the launch's queue holds, so repair_patch_kernel takes its whole-launch path; form = inline: the
    per-warp queues inside the streaming kernel (FIXCA_EXACT_KERNEL=inline)."""
    chk15 = orc.u15_checker()
    n_ties = 0
    if form == "inline":
        tuning("FIXCA_EXACT_KERNEL", "inline")
    for dtype, ch, interp, shifts in itertools.product(("u1", "u2", "u15") if form == "default" else ("u1",), (3, 4), (1, 2),
                                                       ((0.5, -0.5, 0.5, 0.5), (0.25, 0.5, -0.75, 1.5), (0.5, 0.0, 0.0, -0.5))):
        h, w = 203, 1031
        kw = dict(blue=0.0, red=0.0, x_blue=shifts[0], x_red=shifts[1], y_blue=shifts[2], y_red=shifts[3],
                  lens_x=w // 2, lens_y=h // 2, interpolation=interp)
        if dtype == "u15":
            img = orc.synth_u15(h, w, ch, seed=h + ch + interp, wide=False)
            want = chk15.region(img, orc.Params(**kw), bpc=orc.BPC_U15)
            bpc = dict(bpc=fx.BPC_U15)
        else:
            img = orc.synth_image(h, w, ch, dtype, seed=77 + ch + interp)
            want = checker.region(img, orc.Params(**kw))
            bpc = {}
        got = fx.correct(img, fx.FixCaParams(**kw), flags=fx.PRECISION_EXACT, **bpc)
        assert "f32+f64" in fx.last_kernel() or "f64+exact" in fx.last_kernel(), fx.last_kernel()
        assert ("inline" in fx.last_kernel()) == (form == "inline"), fx.last_kernel()
        assert got.tobytes() == want.tobytes(), (dtype, ch, interp, shifts, fx.last_kernel())
        # how many samples sat on an exact tie (Linear, half-pixel shift in x only: (a + b) / 2 with a + b odd)
        if interp == 1 and shifts == (0.5, 0.0, 0.0, -0.5) and dtype != "u15":
            a = img[:, :, 2].astype(np.int64)
            n_ties += int(((a[:, :-1] + a[:, 1:]) & 1).sum())
    assert n_ties > 100000, n_ties


# ---------------------------------------------------------------------------------------------
# BASELINE.json's full sizes: size-independent properties + oracle on sampled bands
# ---------------------------------------------------------------------------------------------
FULL = [  # (name, h, w, ch, dtype, params)  -- SURVEY.md 8(d): every BASELINE.json config at its full size
    ("cfg2-24MP-rgb8-linear", 4000, 6000, 3, "u1", dict(blue=1.0, red=-1.5, lens_x=3000, lens_y=2000, interpolation=1)),
    ("cfg3-8K-rgba16-cubic", 4320, 7680, 4, "u2", dict(blue=6.0, red=-2.4, lens_x=658, lens_y=1280, interpolation=2)),
    ("cfg4-50MP-rgbf32-cubic", 6144, 8192, 3, "f4", dict(KW, lens_x=4096, lens_y=3072, interpolation=2)),
    ("cfg5-4K-rgb8-cubic", 2160, 3840, 3, "u1", dict(KW, blue=1.0, red=-1.5, lens_x=1920, lens_y=1080, interpolation=2)),
    ("target-100MP-rgb16-cubic", 8192, 12288, 3, "u2", dict(KW, lens_x=6144, lens_y=4096, interpolation=2)),
]


def _full_image(seed, h, w, ch, dtype):
    rng = np.random.default_rng(seed)
    if dtype == "f4":       # uniform [0, 1) with a sprinkling outside it, so that clip_d (fix-ca.c:873-880) works at full size
        img = rng.random((h, w, ch), dtype=np.float32)
        img[::97, ::89] = img[::97, ::89] * 2.0 - 0.5
        return img
    return rng.integers(0, np.iinfo(dtype).max, size=(h, w, ch), dtype=dtype, endpoint=True)


def _absdiff(a, b):
    if a.dtype.kind == "f":
        return np.abs(a.astype(np.float64) - b.astype(np.float64))
    return np.abs(a.astype(np.int64) - b.astype(np.int64))


@pytest.mark.parametrize("name,h,w,ch,dtype,kw", FULL, ids=[f[0] for f in FULL])
def test_full_size_properties(fx, checker, name, h, w, ch, dtype, kw, tuning):
    img = _full_image(4, h, w, ch, dtype)
    tol = FLOAT_ABS_TOL if dtype == "f4" else FAST_LSB_TOL
    p = fx.FixCaParams(**kw)
    full = fx.correct(img, p)
    interp_name = "linear" if kw["interpolation"] == 1 else "cubic"
    assert fx.last_kernel().startswith({"u1": "stream/%s/f32+f64" % interp_name, "u2": "stream/%s/f64+exact" % interp_name}.get(dtype, "tiled"))
    # green / alpha untouched everywhere
    assert (full[..., 1] == img[..., 1]).all() and (ch == 3 or (full[..., 3] == img[..., 3]).all())
    # oracle on sampled bands (top edge, an interior band straddling chunk borders, bottom edge)
    for y1, y2 in ((0, 96), (h // 2 - 40, h // 2 + 56), (h - 96, h)):
        want = np.zeros_like(img)
        checker.region(img, orc.Params(**kw), y1, y2, dst=want)
        assert full[y1:y2].tobytes() == want[y1:y2].tobytes(), (name, y1, y2)
    # band-split invariance: two partial calls write the same bytes as the full call
    halves = np.zeros_like(img)
    fx.correct(img, p, y1=0, y2=h // 3, out=halves)
    fx.correct(img, p, y1=h // 3, y2=h, out=halves)
    assert md5(halves) == md5(full)
    del halves
    # the FP64 tile kernel (the exact path of the float formats) computes the same bytes as the repair kernels
    if dtype in ("u1", "u2"):
        tuning("FIXCA_EXACT_KERNEL", "tiled")
        other = fx.correct(img, p)
        assert fx.last_kernel().startswith("tiled") and md5(other) == md5(full)
        del other
    # fast arithmetic: within tolerance of the exact result on the whole image
    fast = fx.correct(img, p, flags=fx.PRECISION_FAST)
    bad = 0
    for y in range(0, h, 512):      # in slabs: a 50 MP float image as float64 would not fit comfortably
        d = _absdiff(fast[y:y + 512], full[y:y + 512])
        assert d.max() <= tol, (name, y, float(d.max()))
        bad += int((d != 0).sum())
    print("%s: fast-vs-exact mismatch fraction %.2e" % (name, bad / fast.size))
    del fast
    # identity with zero amounts at full size (float images: the samples outside [0,1] are clipped, fix-ca.c:873-880)
    ident = fx.correct(img, fx.FixCaParams(lens_x=kw["lens_x"], lens_y=kw["lens_y"], interpolation=kw["interpolation"]),
                       flags=fx.PRECISION_FAST)
    if dtype == "f4":
        want = img.copy()
        want[..., 0::2] = np.clip(img[..., 0::2], 0.0, 1.0)
        assert np.array_equal(ident, want)
    else:
        assert md5(ident) == md5(img)


@pytest.mark.parametrize("name,h,w,ch,dtype,kw", FULL, ids=[f[0] for f in FULL])
def test_full_size_fast_within_one_lsb(fx, checker, name, h, w, ch, dtype, kw, tuning):
    """The bench configuration itself (FAST arithmetic, streaming kernel) at BASELINE.json's sizes:
    +-1 LSB (1e-6 for float) on sampled bands, pass-through channels identical, and band-split invariance."""
    img = _full_image(5, h, w, ch, dtype)
    tol = FLOAT_ABS_TOL if dtype == "f4" else FAST_LSB_TOL
    p = fx.FixCaParams(**kw)
    full = fx.correct(img, p, flags=fx.PRECISION_FAST)
    assert fx.last_kernel().startswith("stream")
    assert (full[..., 1] == img[..., 1]).all() and (ch == 3 or (full[..., 3] == img[..., 3]).all())
    worst, nbad, ntot = 0, 0, 0
    for y1, y2 in ((0, 96), (h // 2 - 40, h // 2 + 56), (h - 96, h)):
        want = np.zeros_like(img)
        checker.region(img, orc.Params(**kw), y1, y2, dst=want)
        d = _absdiff(full[y1:y2], want[y1:y2])
        worst = max(worst, float(d.max()))
        nbad += int((d != 0).sum())
        ntot += d.size
    assert worst <= tol, (name, worst)
    if dtype != "f4":
        assert nbad / ntot < 5e-3, (name, nbad / ntot)      # ~2e-3 of u16 samples
    # a band computed on its own equals the same rows of the full call (row-band independence)
    y1, y2 = h // 3, h // 3 + 77
    band = np.zeros_like(img)
    fx.correct(img, p, y1=y1, y2=y2, out=band, flags=fx.PRECISION_FAST)
    assert (band[y1:y2] == full[y1:y2]).all()
    # the per-tile fallback kernel shares the arithmetic (same weights, same FMA order): same bytes for integer
    # samples; float samples agree within the tolerance (columns whose taps are clamped to an image edge add the
    # merged weights in a different order)
    tuning("FIXCA_FAST_KERNEL", "strip")
    fx.correct(img, p, y1=y1, y2=y2, out=band, flags=fx.PRECISION_FAST)
    assert fx.last_kernel().startswith("strip")
    if dtype == "f4":
        assert _absdiff(band[y1:y2], full[y1:y2]).max() <= FLOAT_ABS_TOL
    else:
        assert (band[y1:y2] == full[y1:y2]).all()


def test_cfg5_frame_batch_full_size(fx, checker):
    """BASELINE configs[4] at its frame size: 8 device-resident 3840 x 2160 RGB8 frames in ONE fixca_cuda_frames_dev
    launch (Cubic, lateral + directional, FAST): every frame within 1 LSB of the reference on sampled bands, identical to
    its own single-frame call, pass-through channel copied; EXACT on the same batch is bit-identical to the reference."""
    import torch
    h, w, ch, nf = 2160, 3840, 3, 8
    kw = dict(KW, blue=1.0, red=-1.5, lens_x=1920, lens_y=1080, interpolation=2)
    p = fx.FixCaParams(**kw)
    rng = np.random.default_rng(55)
    frames = rng.integers(0, 255, size=(nf, h, w, ch), dtype=np.uint8, endpoint=True)
    bpp, pitch = 3, (w * 3 + 127) // 128 * 128
    src = torch.zeros((nf, h, pitch), dtype=torch.uint8, device="cuda")
    src[:, :, :w * bpp] = torch.from_numpy(frames.reshape(nf, h, w * bpp)).cuda()
    dst = torch.zeros_like(src)
    st = torch.cuda.current_stream().cuda_stream
    for flags, tol in ((fx.PRECISION_FAST, FAST_LSB_TOL), (fx.PRECISION_EXACT, 0)):
        dst.zero_()
        n0 = fx.launch_count()
        fx.fix_ca_frames_dev(src.data_ptr(), pitch, pitch * h, dst.data_ptr(), pitch, pitch * h, nf, w, h, bpp, 1, p, flags, st)
        torch.cuda.synchronize()
        if flags == fx.PRECISION_FAST:
            assert fx.launch_count() - n0 == 1 and fx.last_kernel().startswith("stream")     # the whole batch in one launch
        got = dst[:, :, :w * bpp].cpu().numpy().reshape(nf, h, w, ch)
        assert np.array_equal(got[..., 1], frames[..., 1])
        for k in (0, 3, nf - 1):
            for y1, y2 in ((0, 48), (h // 2 - 20, h // 2 + 28), (h - 48, h)):
                want = np.zeros((h, w, ch), np.uint8)
                checker.region(frames[k], orc.Params(**kw), y1, y2, dst=want)
                d = _absdiff(got[k, y1:y2], want[y1:y2])
                assert d.max() <= tol, (k, y1, y2, flags)
        if flags == fx.PRECISION_FAST:
            single = fx.correct(frames[nf - 2], p, flags=flags)
            assert single.tobytes() == got[nf - 2].tobytes()


# ---------------------------------------------------------------------------------------------
# the preview call: show_progress = FALSE (fix-ca.c:656-657, :1322-1327)
# ---------------------------------------------------------------------------------------------
def test_preview_call_matches_golden(fx):
    """fixca_cuda_region(..., show_progress = 0): correction + saturate() + centerline(), bit-exact against digests
    made by the reference's own code (HSV pair restated, see oracle/ref_harness.c)."""
    bad = []
    for c in golden()["preview"]:
        img = case_image(c)
        out = np.zeros_like(img)
        fx.fix_ca_region(img, out, c["w"], c["h"], img.shape[2] * img.dtype.itemsize, fx.bpc_of(img.dtype),
                         fx_params(fx, c), 0, c["w"], 0, c["h"], False)
        if md5(out) != c["md5"]:
            bad.append(c["name"])
    assert not bad, "%d of %d preview cases differ, first: %s" % (len(bad), len(golden()["preview"]), bad[:6])


def test_preview_band_and_device_flag(fx, checker):
    """The dialog previews a row band of the whole drawable (fix-ca.c:656-657); the device-resident entry takes
    FIXCA_PREVIEW_OVERLAY for the same result."""
    import torch

    img = orc.synth_image(300, 640, 3, "u2", 77)
    kw = dict(KW, lens_x=300, lens_y=140, interpolation=1, saturation=40.0)
    want = np.zeros_like(img)
    checker.region(img, orc.Params(**kw), 120, 190, dst=want, preview=True)
    got = np.zeros_like(img)
    fx.fix_ca_region(img, got, 640, 300, 6, 2, fx.FixCaParams(**kw), 0, 640, 120, 190, False)
    assert got.tobytes() == want.tobytes()
    src = torch.from_numpy(img.view(np.int16)).cuda()
    dst = torch.zeros_like(src)
    fx.fix_ca_region_dev(src.data_ptr(), 640 * 6, 0, 300, dst.data_ptr(), 640 * 6, 0, 640, 300, 6, 2, fx.FixCaParams(**kw),
                         120, 190, fx.PRECISION_EXACT | fx.PREVIEW_OVERLAY, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert dst.cpu().numpy().view(np.uint16)[120:190].tobytes() == want[120:190].tobytes()
