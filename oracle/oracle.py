"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for the two CPU checkers:

  * ``Restatement``  -> oracle/libfixca_oracle.so  (our plain-C restatement,
    oracle/fixca_oracle.c; built anywhere by oracle/Makefile)
  * ``Reference``    -> oracle/_ref/libfixca_ref.so (the reference's own fix-ca.c
    compiled unmodified behind shim headers; built only where /root/reference
    exists, travels prebuilt to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs import this module.  The product package
(gimp-fix-ca_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RESTATEMENT_SO = os.path.join(HERE, "libfixca_oracle.so")
REFERENCE_SO = os.path.join(HERE, "_ref", "libfixca_ref.so")
REFERENCE_HALF_SO = os.path.join(HERE, "_ref", "libfixca_ref_half.so")
REFERENCE_U15_SO = os.path.join(HERE, "_ref", "libfixca_ref_u15.so")
BPC_U15 = 15    # 15-bit unsigned samples held in uint16 (extension; include/fixca_cuda.h FIXCA_BPC_U15)

_c_int = ctypes.c_int
_c_dp = ctypes.POINTER(ctypes.c_double)
_c_vp = ctypes.c_void_p


@dataclass
class Params:
    """FixCaParams (fix-ca.c:70-82) without the unused update_preview field."""

    blue: float = 0.0
    red: float = 0.0
    lens_x: float = -1.0
    lens_y: float = -1.0
    interpolation: int = 1
    saturation: float = 0.0
    x_blue: float = 0.0
    x_red: float = 0.0
    y_blue: float = 0.0
    y_red: float = 0.0

    def as_array(self):
        return (ctypes.c_double * 10)(
            self.blue, self.red, self.lens_x, self.lens_y, float(self.interpolation),
            self.saturation, self.x_blue, self.x_red, self.y_blue, self.y_red)


def build(force: bool = False) -> None:
    """Run oracle/Makefile (restatement always; _ref only where the reference is mounted)."""
    if force or not os.path.exists(RESTATEMENT_SO) or (
            os.path.exists("/root/reference/fix-ca.c") and not (os.path.exists(REFERENCE_SO) and os.path.exists(
                REFERENCE_HALF_SO) and os.path.exists(REFERENCE_U15_SO) and os.path.exists(os.path.join(HERE, "_ref", "libfixca_plugin_cuda.so")) and
                os.path.exists(os.path.join(HERE, "_ref", "libfixca_plugin_cuda_half.so")))):
        subprocess.run(["make", "-C", HERE, "-s"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _img_args(src: np.ndarray):
    assert src.ndim == 3 and src.flags["C_CONTIGUOUS"]
    h, w, ch = src.shape
    b = src.dtype.itemsize
    if src.dtype.kind == "f":
        bpc = -b
    else:
        bpc = b
    return h, w, ch * b, bpc


class Restatement:
    def __init__(self):
        build()
        self.lib = ctypes.CDLL(RESTATEMENT_SO)
        self.lib.fixca_oracle_region.argtypes = [_c_vp, _c_vp] + [_c_int] * 4 + [_c_dp] + [_c_int] * 4
        self.lib.fixca_oracle_region.restype = _c_int
        self.lib.fixca_oracle_region_mt.argtypes = [_c_vp, _c_vp] + [_c_int] * 4 + [_c_dp] + [_c_int] * 3
        self.lib.fixca_oracle_region_mt.restype = _c_int
        self.lib.fixca_oracle_region_preview.argtypes = [_c_vp, _c_vp] + [_c_int] * 4 + [_c_dp] + [_c_int] * 2
        self.lib.fixca_oracle_region_preview.restype = _c_int
        self.lib.fixca_oracle_axis.argtypes = [_c_int, _c_int, _c_dp, _c_int, _c_int, _c_vp, _c_vp]
        self.lib.fixca_oracle_axis.restype = _c_int
        self.lib.fixca_oracle_resolve_lens.argtypes = [_c_int, _c_int, _c_dp, _c_dp]

    kind = "port"

    def region(self, src: np.ndarray, p: Params, y1=None, y2=None, dst=None, threads: int = 1, preview: bool = False,
               bpc=None):
        """preview=True is the call with show_progress == FALSE: saturation boost + centre lines on top.
        ``bpc`` overrides the code derived from the dtype (BPC_U15 for 15-bit samples held in uint16)."""
        h, w, bytes_, bpc0 = _img_args(src)
        bpc = bpc0 if bpc is None else bpc
        y1 = 0 if y1 is None else y1
        y2 = h if y2 is None else y2
        if dst is None:
            dst = np.zeros_like(src)
        if preview:
            rc = self.lib.fixca_oracle_region_preview(src.ctypes.data, dst.ctypes.data, w, h, bytes_, bpc,
                                                      p.as_array(), y1, y2)
        elif threads > 1:
            rc = self.lib.fixca_oracle_region_mt(src.ctypes.data, dst.ctypes.data, w, h, bytes_, bpc,
                                                 p.as_array(), y1, y2, threads)
        else:
            rc = self.lib.fixca_oracle_region(src.ctypes.data, dst.ctypes.data, w, h, bytes_, bpc,
                                              p.as_array(), 0, w, y1, y2)
        if rc:
            raise ValueError("fixca_oracle_region rc=%d" % rc)
        return dst

    def axis(self, w: int, h: int, p: Params, channel: int, axis: int):
        n = h if axis else w
        idx = np.zeros(n, dtype=np.int32)
        frac = np.zeros(n, dtype=np.float64)
        rc = self.lib.fixca_oracle_axis(w, h, p.as_array(), channel, axis, idx.ctypes.data, frac.ctypes.data)
        if rc:
            raise ValueError("fixca_oracle_axis rc=%d" % rc)
        return idx, frac

    def resolve_lens(self, w: int, h: int, lx: float, ly: float):
        a, b = ctypes.c_double(lx), ctypes.c_double(ly)
        self.lib.fixca_oracle_resolve_lens(w, h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value


class Reference:
    """The reference's own code.  ``available()`` is False where it was never built."""

    kind = "reference"

    @staticmethod
    def available() -> bool:
        build()
        return os.path.exists(REFERENCE_SO)

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(REFERENCE_SO)
        L = self.lib
        L.ref_fix_ca_region.argtypes = [_c_vp, _c_vp] + [_c_int] * 4 + [_c_dp] + [_c_int] * 5
        L.ref_fix_ca_region.restype = None
        L.ref_fix_ca_region_mt.argtypes = [_c_vp, _c_vp] + [_c_int] * 4 + [_c_dp] + [_c_int] * 3
        L.ref_fix_ca_region_mt.restype = None
        L.ref_sizeof_params.restype = _c_int
        L.ref_color_size.argtypes = [ctypes.c_char_p, _c_int]
        L.ref_color_size.restype = _c_int
        L.ref_fake_set_drawable.argtypes = [_c_int] * 3 + [ctypes.c_char_p, _c_vp]
        L.ref_fake_set_selection.argtypes = [_c_int] * 4
        L.ref_fake_set_dialog_response.argtypes = [_c_int]
        L.ref_fake_set_saved_params.argtypes = [_c_dp]
        L.ref_fake_get_saved_params.argtypes = [_c_dp]
        L.ref_fake_get_saved_params.restype = _c_int
        L.ref_fake_last_message.restype = ctypes.c_char_p
        L.ref_fake_counter.argtypes = [_c_int]
        L.ref_fake_counter.restype = _c_int
        L.ref_run.argtypes = [ctypes.c_char_p, _c_int, _c_int, _c_dp, _c_int]
        L.ref_run.restype = _c_int
        L.ref_dialog_lens.argtypes = [_c_int, _c_int, _c_dp, _c_dp]
        L.ref_preview_update.argtypes = [_c_int] * 4 + [_c_dp, _c_vp]
        L.ref_preview_update.restype = _c_int

    def region(self, src: np.ndarray, p: Params, y1=None, y2=None, dst=None, threads: int = 1, preview: bool = False,
               bpc=None):
        h, w, bytes_, bpc0 = _img_args(src)
        bpc = bpc0 if bpc is None else bpc
        y1 = 0 if y1 is None else y1
        y2 = h if y2 is None else y2
        if dst is None:
            dst = np.zeros_like(src)
        if threads > 1 and not preview:
            self.lib.ref_fix_ca_region_mt(src.ctypes.data, dst.ctypes.data, w, h, bytes_, bpc,
                                          p.as_array(), y1, y2, threads)
        else:
            self.lib.ref_fix_ca_region(src.ctypes.data, dst.ctypes.data, w, h, bytes_, bpc,
                                       p.as_array(), 0, w, y1, y2, 0 if preview else 1)
        return dst

    def color_size(self, name: str, bpp: int) -> int:
        return self.lib.ref_color_size(name.encode(), bpp)

    def run(self, pixels: np.ndarray, fmt: str, run_mode: int, nparams: int, blue=0.0, red=0.0,
            lens_x=-1.0, lens_y=-1.0, interpolation=0, x_blue=0.0, x_red=0.0, y_blue=0.0, y_red=0.0,
            proc_name="Test-Fix-CA", selection=None, saved=None, dialog_ok=True):
        """Drive the reference's real run() on an in-memory drawable (modified in place)."""
        h, w, ch = pixels.shape
        self.lib.ref_fake_set_drawable(w, h, ch * pixels.dtype.itemsize, fmt.encode(), pixels.ctypes.data)
        if selection is not None:
            self.lib.ref_fake_set_selection(*selection)
        if saved is not None:
            self.lib.ref_fake_set_saved_params(saved.as_array())
        self.lib.ref_fake_set_dialog_response(1 if dialog_ok else 0)
        f = (ctypes.c_double * 8)(blue, red, lens_x, lens_y, x_blue, x_red, y_blue, y_red)
        return self.lib.ref_run(proc_name.encode(), run_mode, nparams, f, interpolation)

    def preview_update(self, pixels: np.ndarray, fmt: str, x: int, y: int, w: int, h: int, p: Params):
        """The dialog's preview refresh (fix-ca.c:617-679) of window (x, y, w, h): returns the 8-bit buffer
        the plug-in draws, shape (h, w, channels)."""
        hh, ww, ch = pixels.shape
        self.lib.ref_fake_set_drawable(ww, hh, ch * pixels.dtype.itemsize, fmt.encode(), pixels.ctypes.data)
        out = np.zeros((h, w, ch), dtype=np.uint8)
        stride = self.lib.ref_preview_update(x, y, w, h, p.as_array(), out.ctypes.data)
        assert stride == w * ch, (stride, w, ch)
        return out

    def last_message(self) -> str:
        return self.lib.ref_fake_last_message().decode()

    def counter(self, which: int) -> int:
        return self.lib.ref_fake_counter(which)

    def dialog_lens(self, w: int, h: int, lx: float, ly: float):
        a, b = ctypes.c_double(lx), ctypes.c_double(ly)
        self.lib.ref_dialog_lens(w, h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value


PLUGIN_CUDA_SO = os.path.join(HERE, "_ref", "libfixca_plugin_cuda.so")


class PatchedPlugin(Reference):
    """The reference plug-in with INTEGRATION.md's patch applied (oracle/patch_plugin.py): the same
    run() / fix_ca() code, behind the same fake GIMP, but its final-render call goes through
    fixca_cuda_region() when a GPU is present.  Used to prove the drop-in at the real call site."""

    kind = "reference+cuda"

    @staticmethod
    def available() -> bool:
        build()
        return os.path.exists(PLUGIN_CUDA_SO)

    def __init__(self):
        build()
        so = REFERENCE_SO
        try:
            globals()["REFERENCE_SO"] = PLUGIN_CUDA_SO
            Reference.__init__(self)
        finally:
            globals()["REFERENCE_SO"] = so


class ReferenceHalf(Reference):
    """The reference with its own commented-out half-precision lines enabled (oracle/patch_half.py,
    fix-ca.c:692-693, :740-742, :768-770): the checker for float16 images (bpc = -2).  For every other
    format it computes what ``Reference`` computes."""

    kind = "reference+half"

    @staticmethod
    def available() -> bool:
        build()
        return os.path.exists(REFERENCE_HALF_SO)

    def __init__(self):
        build()
        so = REFERENCE_SO
        try:
            globals()["REFERENCE_SO"] = REFERENCE_HALF_SO
            Reference.__init__(self)
        finally:
            globals()["REFERENCE_SO"] = so


class ReferenceU15(Reference):
    """The reference with u15 samples written into color_size() / get_pixel() / set_pixel() in the pattern of
    its other unsigned types (oracle/patch_u15.py; fix-ca.c:694-695 only rejects them): the checker for
    bpc = 15.  For every other format it computes what ``Reference`` computes."""

    kind = "reference+u15"

    @staticmethod
    def available() -> bool:
        build()
        return os.path.exists(REFERENCE_U15_SO)

    def __init__(self):
        build()
        so = REFERENCE_SO
        try:
            globals()["REFERENCE_SO"] = REFERENCE_U15_SO
            Reference.__init__(self)
        finally:
            globals()["REFERENCE_SO"] = so


PLUGIN_CUDA_HALF_SO = os.path.join(HERE, "_ref", "libfixca_plugin_cuda_half.so")


class PatchedPluginHalf(Reference):
    """INTEGRATION.md's patch on top of patch_half.py: the plug-in with its half lines enabled, calling
    fixca_cuda_region() -- drives `R'G'B' half` drawables through run() on the GPU."""

    kind = "reference+half+cuda"

    @staticmethod
    def available() -> bool:
        build()
        return os.path.exists(PLUGIN_CUDA_HALF_SO)

    def __init__(self):
        build()
        so = REFERENCE_SO
        try:
            globals()["REFERENCE_SO"] = PLUGIN_CUDA_HALF_SO
            Reference.__init__(self)
        finally:
            globals()["REFERENCE_SO"] = so


def best_checker():
    """The strongest CPU checker present: the reference's own code, else the restatement."""
    return Reference() if Reference.available() else Restatement()


def half_checker():
    """Checker for float16 images: the reference with its half lines enabled, else the restatement."""
    return ReferenceHalf() if ReferenceHalf.available() else Restatement()


def u15_checker():
    """Checker for u15 images (bpc = 15): the reference with the u15 rows written in, else the restatement."""
    return ReferenceU15() if ReferenceU15.available() else Restatement()


def synth_u15(h: int, w: int, ch: int, seed: int, wide: bool = False) -> np.ndarray:
    """Seeded u15 image in uint16 storage: values 0 .. 32768; ``wide`` also draws the out-of-range codes
    32769 .. 65535 (they decode to > 1.0 and are clipped by Linear / Cubic, copied by None)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 65535 if wide else 32768, size=(h, w, ch), dtype=np.uint16, endpoint=True)


U64_EXTREMES = np.array([0, 1, 2, (1 << 53) - 1, 1 << 53, (1 << 53) + 1, (1 << 63) - 1, 1 << 63, (1 << 63) + 1,
                         (1 << 64) - (1 << 11) - 1, (1 << 64) - (1 << 11), (1 << 64) - (1 << 10) - 1, (1 << 64) - (1 << 10),
                         (1 << 64) - (1 << 10) + 1, (1 << 64) - 2, (1 << 64) - 1, 0x8000000000000400, 0x80000000000003FF,
                         0x8000000000000C00, 0x0000000100000000, 0x00000000FFFFFFFF], dtype=np.uint64)


def synth_u64(h: int, w: int, ch: int, seed: int, extremes: bool = False) -> np.ndarray:
    """Seeded full-range 64-bit samples; ``extremes``: half of them drawn from U64_EXTREMES -- the boundaries of the two
    roundings in get_pixel's long double division (fix-ca.c:728-733) and the values that decode to 1.0, which
    set_pixel (:759-761) wraps to 0 in the reference as compiled."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, np.iinfo(np.uint64).max, size=(h, w, ch), dtype=np.uint64, endpoint=True)
    if extremes:
        pick = rng.integers(0, len(U64_EXTREMES), size=(h, w, ch))
        img = np.where(rng.random((h, w, ch)) < 0.5, U64_EXTREMES[pick], img)
    return np.ascontiguousarray(img)


def synth_image(h: int, w: int, ch: int, dtype: str, seed: int, wide: bool = False) -> np.ndarray:
    """Seeded synthetic image (PCG64), uniform noise: worst case for caches.
    ``wide`` draws floats from [-0.5, 1.5) to exercise clip_d (fix-ca.c:873-880)."""
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    if dt.kind == "f":
        a = rng.random((h, w, ch), dtype=np.float64)
        if wide:
            a = a * 2.0 - 0.5
        return np.ascontiguousarray(a.astype(dt))
    return rng.integers(0, np.iinfo(dt).max, size=(h, w, ch), dtype=dt, endpoint=True)
