"""The drop-in at the real call site: the reference plug-in's own run() -> fix_ca(), behind the fake
GIMP backend of oracle/ref_harness.c, with INTEGRATION.md's patch applied (oracle/patch_plugin.py) so
that fix-ca.c:373-374 calls fixca_cuda_region().  Compared with the unpatched plug-in on the same
drawable and PDB arguments: pixels, PDB status, progress protocol and messages must be identical.

CPU part: the patched plug-in builds, loads and (no GPU) keeps its own row loop.
GPU part (-m gpu): the same calls, now served by the CUDA library (its launch counter moves)."""
import numpy as np
import pytest

import oracle as orc

NONINTERACTIVE, INTERACTIVE, WITH_LAST_VALS = 1, 0, 2
PDB_SUCCESS, PDB_CALLING_ERROR = 3, 1

DRAWABLES = [
    # shape, dtype, babl format name
    ((120, 200, 3), "u1", "R'G'B' u8"),
    ((97, 131, 4), "u2", "R'G'B'A u16"),
    ((64, 160, 3), "f4", "RGB float"),
    ((33, 48, 4), "u4", "R'G'B'A u32"),
    ((40, 56, 3), "f8", "RGB double"),
]

CALLS = [
    dict(run_mode=NONINTERACTIVE, nparams=12, blue=6.0, red=-2.4, lens_x=658.0, lens_y=1280.0, interpolation=1),
    dict(run_mode=NONINTERACTIVE, nparams=12, blue=3.0, red=-2.0, lens_x=50.0, lens_y=30.0, interpolation=2,
         x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9),
    dict(run_mode=NONINTERACTIVE, nparams=12, blue=-4.0, red=5.0, interpolation=0, x_blue=2.0, y_red=-3.0),
    dict(run_mode=NONINTERACTIVE, nparams=5, blue=2.0, red=-1.0),                 # missing args take defaults
    dict(run_mode=NONINTERACTIVE, nparams=12, blue=31.0, red=0.0, interpolation=1),   # out of range -> calling error
    dict(run_mode=INTERACTIVE, nparams=3, dialog_ok=True),                          # dialog path, lens reset
]


@pytest.fixture(scope="module")
def plugins():
    if not orc.Reference.available() or not orc.PatchedPlugin.available():
        pytest.skip("oracle/_ref plug-in builds missing (they need /root/reference at build time)")
    return orc.Reference(), orc.PatchedPlugin()


def _drive(plugin, pixels, fmt, call):
    px = pixels.copy()
    kw = {k: v for k, v in call.items()}
    status = plugin.run(px, fmt, kw.pop("run_mode"), kw.pop("nparams"), **kw)
    # the fake backend zeroes its counters when the drawable is installed (start of run())
    return px, status, [plugin.counter(i) for i in range(5)], plugin.last_message()


def _compare(plugins, expect_gpu):
    import fixca

    ref, patched = plugins
    n = 0
    for (shape, dt, fmt), call in [(d, c) for d in DRAWABLES for c in CALLS]:
        n += 1
        img = orc.synth_image(shape[0], shape[1], shape[2], dt, seed=7000 + n)
        want_px, want_status, want_counts, _ = _drive(ref, img, fmt, call)
        launches = fixca.launch_count()
        got_px, got_status, got_counts, msg = _drive(patched, img, fmt, call)
        assert got_status == want_status, (fmt, call, msg)
        assert got_px.tobytes() == want_px.tobytes(), (fmt, call)
        assert got_counts == want_counts, (fmt, call, got_counts, want_counts)    # progress init/update, messages, merge, flush
        if expect_gpu and want_status == PDB_SUCCESS:
            assert fixca.launch_count() > launches, "the patched plug-in did not reach the CUDA library"
        if not expect_gpu:
            assert fixca.launch_count() == launches


PREVIEWS = [
    # window (x, y, w, h) of the drawable the dialog shows, saturation, interpolation, lens
    ((0, 0, 64, 48), 0.0, 1, (30.0, 20.0)),
    ((17, 23, 80, 33), 35.0, 2, (60.0, 40.0)),
    ((5, 60, 100, 30), -60.0, 0, (-1.0, -1.0)),
]


def _compare_previews(plugins, expect_gpu):
    """preview_update() (fix-ca.c:617-679): whole-drawable fetch, the pass on the visible row band with
    show_progress = FALSE (overlay + saturation), 8-bit down-conversion, gimp_preview_draw_buffer()."""
    import fixca

    ref, patched = plugins
    n = 0
    for (shape, dt, fmt), (win, sat, interp, lens) in [(d, c) for d in DRAWABLES for c in PREVIEWS]:
        n += 1
        h, w = shape[0], shape[1]
        x, y, ww, hh = win
        ww, hh = min(ww, w - x), min(hh, h - y)
        if ww <= 0 or hh <= 0:
            continue        # the window lies outside this (small) drawable
        # float drawables: also samples outside [0,1] (pass-through channels are not clipped; set_pixel(.., 1) wraps them)
        img = orc.synth_image(h, w, shape[2], dt, seed=8000 + n, wide=bool(n % 2))
        P = orc.Params(blue=3.0, red=-2.0, lens_x=lens[0], lens_y=lens[1], interpolation=interp, saturation=sat,
                       x_blue=0.7, y_red=-0.9)
        want = ref.preview_update(img.copy(), fmt, x, y, ww, hh, P)
        launches = fixca.launch_count()
        got = patched.preview_update(img.copy(), fmt, x, y, ww, hh, P)
        assert got.tobytes() == want.tobytes(), (fmt, win, sat, interp)
        assert (fixca.launch_count() > launches) == expect_gpu


def test_patched_plugin_without_gpu_keeps_its_cpu_loop(plugins, fx):
    if fx.device_count() > 0:
        pytest.skip("a GPU is present; covered by the gpu test")
    _compare(plugins, expect_gpu=False)
    _compare_previews(plugins, expect_gpu=False)


@pytest.mark.gpu
def test_patched_plugin_runs_on_the_gpu_bit_identically(plugins, fx):
    assert fx.device_count() > 0
    _compare(plugins, expect_gpu=True)
    _compare_previews(plugins, expect_gpu=True)


# ---------------------------------------------------------------------------------------------
# half-precision drawables: the plug-in with its own commented-out half lines enabled (patch_half.py),
# unpatched vs with INTEGRATION.md's patch on top
# ---------------------------------------------------------------------------------------------
HALF_DRAWABLES = [((90, 140, 3), "f2", "R'G'B' half"), ((61, 97, 4), "f2", "R'G'B'A half")]


@pytest.mark.gpu
def test_half_drawables_through_the_patched_plugin(fx):
    import fixca

    if not orc.ReferenceHalf.available() or not orc.PatchedPluginHalf.available():
        pytest.skip("oracle/_ref half plug-in builds missing (they need /root/reference at build time)")
    ref, patched = orc.ReferenceHalf(), orc.PatchedPluginHalf()
    n = 0
    for (shape, dt, fmt), call in [(d, c) for d in HALF_DRAWABLES for c in CALLS[:4]]:
        n += 1
        img = orc.synth_image(shape[0], shape[1], shape[2], dt, seed=9000 + n, wide=bool(n % 2))
        want_px, want_status, want_counts, _ = _drive(ref, img, fmt, call)
        launches = fixca.launch_count()
        got_px, got_status, got_counts, msg = _drive(patched, img, fmt, call)
        assert got_status == want_status == PDB_SUCCESS, (fmt, call, msg)
        assert got_px.tobytes() == want_px.tobytes(), (fmt, call)
        assert got_counts == want_counts
        assert fixca.launch_count() > launches
    # the shipped reference refuses the same drawable (color_size -> -99, fix-ca.c:692-697)
    plain = orc.Reference()
    img = orc.synth_image(20, 30, 3, "f2", seed=1)
    _, status, _, _ = _drive(plain, img, "R'G'B' half", CALLS[0])
    assert status != PDB_SUCCESS
