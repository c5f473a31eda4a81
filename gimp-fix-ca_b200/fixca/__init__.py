"""fixca -- thin ctypes binding over the C ABI of include/fixca_cuda.h.

Mirrors the reference plug-in's interface for the correction pass:

* ``FixCaParams``            <- ``FixCaParams`` (fix-ca.c:70-82), same field names
* ``fix_ca_region(...)``     <- ``fix_ca_region()`` (fix-ca.c:998-1001), same
                                argument order and meaning, host (numpy) buffers
* ``fix_ca_region_dev(...)`` device-resident variant (raw CUDA pointers)
* ``resolve_lens``, ``check_params``, ``color_size``, ``band_source_rows`` ...

This module does no computing of its own: every call goes through
``lib/libfixca_cuda.so`` (hand-written sm_100a kernels).  If the library is
missing, importing fails; if no GPU is usable the compute calls raise
``FixCaError`` -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FIXCA_LIB: developer override for A/B runs of two builds of the same library (scripts/quick_bench.py)
LIB_PATH = os.environ.get("FIXCA_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libfixca_cuda.so")

INTERP_NONE, INTERP_LINEAR, INTERP_CUBIC = 0, 1, 2
PRECISION_EXACT, PRECISION_FAST = 0x0, 0x1
FORCE_DIRECT, FORCE_TILED = 0x10, 0x20
PREVIEW_OVERLAY = 0x40      # saturate() + centerline() on the rows written (the show_progress=False call)
COLUMN_SELECTION = 0x80     # accept x1 != 0 / x2 != width: columns [x1, x2) of the full-width result (extension)
PADDING_SCRATCH = 0x100     # device entries: the bytes up to the next 16-byte boundary of each dst row are scratch
INPUT_MAX = 30.0

OK = 0
ERR_ARG, ERR_FORMAT, ERR_INTERP, ERR_REGION, ERR_DEGENERATE = -1, -2, -3, -4, -5
ERR_RANGE, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = -6, -7, -8, -9, -10

# Every symbol include/fixca_cuda.h declares.
EXPORTS = (
    "fixca_cuda_region", "fixca_cuda_region_ex", "fixca_cuda_region_multi", "fixca_cuda_region_dev",
    "fixca_cuda_frames", "fixca_cuda_frames_dev", "fixca_band_source_rows", "fixca_split_bands", "fixca_resolve_lens",
    "fixca_check_params", "fixca_color_size", "fixca_color_size_half", "fixca_params_default", "fixca_cuda_set_progress",
    "fixca_cuda_last_error", "fixca_strerror", "fixca_cuda_device_count", "fixca_cuda_last_kernel",
    "fixca_cuda_launch_count", "fixca_cuda_release", "fixca_version",
    "fixca_cuda_frames_multi", "fixca_cuda_frame_alloc", "fixca_cuda_frame_open", "fixca_cuda_frame_close", "fixca_cuda_frame_free",
    "fixca_cuda_host_alloc", "fixca_cuda_host_free", "fixca_cuda_reload_tuning", "fixca_color_size_ext", "fixca_cuda_last_call_ms", "fixca_cuda_preview", "fixca_cuda_region_dev_fanout",
)
BPC_HALF, BPC_U15 = -2, 15
IPC_HANDLE_BYTES = 64


class FixCaParams(ctypes.Structure):
    """Layout-identical to the reference's FixCaParams (fix-ca.c:70-82), 80 bytes."""

    _fields_ = [
        ("blue", ctypes.c_double),
        ("red", ctypes.c_double),
        ("lens_x", ctypes.c_double),
        ("lens_y", ctypes.c_double),
        ("update_preview", ctypes.c_int),
        ("interpolation", ctypes.c_int),
        ("saturation", ctypes.c_double),
        ("x_blue", ctypes.c_double),
        ("x_red", ctypes.c_double),
        ("y_blue", ctypes.c_double),
        ("y_red", ctypes.c_double),
    ]

    def __init__(self, blue=0.0, red=0.0, lens_x=-1.0, lens_y=-1.0, interpolation=INTERP_LINEAR,
                 saturation=0.0, x_blue=0.0, x_red=0.0, y_blue=0.0, y_red=0.0, update_preview=1):
        # defaults of fix_ca_params_default (fix-ca.c:85-97)
        super().__init__(blue, red, lens_x, lens_y, update_preview, interpolation, saturation,
                         x_blue, x_red, y_blue, y_red)


class FixCaError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__("fixca error %d: %s" % (code, text))
        self.code = code


PROGRESS_FN = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.c_double, ctypes.c_void_p)

_lib = None


def load() -> ctypes.CDLL:
    """Load lib/libfixca_cuda.so (raises OSError when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("%s not found: build it with `make -C gimp-fix-ca_b200` "
                      "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    i, vp, pp = ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(FixCaParams)
    L.fixca_cuda_region.argtypes = [vp, vp, i, i, i, i, pp, i, i, i, i, i]
    L.fixca_cuda_preview.argtypes = [vp, vp, i, i, i, i, pp, i, i, i, i]
    L.fixca_cuda_region_ex.argtypes = [vp, vp, i, i, i, i, pp, i, i, i, i, i, ctypes.c_uint, i]
    L.fixca_cuda_region_multi.argtypes = [vp, vp, i, i, i, i, pp, i, i, ctypes.c_uint, ctypes.POINTER(i), i]
    L.fixca_cuda_region_dev.argtypes = [vp, ctypes.c_size_t, i, i, vp, ctypes.c_size_t, i, i, i, i, i, pp, i, i,
                                        ctypes.c_uint, vp]
    L.fixca_cuda_region_dev_fanout.argtypes = [vp, ctypes.c_size_t, i, i, ctypes.POINTER(vp), i, ctypes.c_size_t, i, i, i, i, i, pp,
                                               i, i, ctypes.c_uint, vp]
    L.fixca_cuda_frames.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), i, i, i, i, i, pp, ctypes.c_uint, i]
    L.fixca_cuda_frames_multi.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), i, i, i, i, i, pp, ctypes.c_uint,
                                          ctypes.POINTER(i), i]
    L.fixca_cuda_frames_dev.argtypes = [vp, ctypes.c_size_t, ctypes.c_size_t, vp, ctypes.c_size_t, ctypes.c_size_t, i,
                                        i, i, i, i, pp, ctypes.c_uint, vp]
    L.fixca_cuda_frame_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp), ctypes.c_char_p]
    L.fixca_cuda_frame_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    L.fixca_cuda_frame_close.argtypes = [vp]
    L.fixca_cuda_frame_free.argtypes = [vp]
    L.fixca_band_source_rows.argtypes = [i, i, pp, i, i, ctypes.POINTER(i), ctypes.POINTER(i)]
    L.fixca_split_bands.argtypes = [i, i, i, ctypes.POINTER(i), ctypes.POINTER(i)]
    L.fixca_resolve_lens.argtypes = [i, i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    L.fixca_resolve_lens.restype = None
    L.fixca_check_params.argtypes = [pp]
    L.fixca_color_size.argtypes = [ctypes.c_char_p, i]
    L.fixca_color_size_half.argtypes = [ctypes.c_char_p, i]
    L.fixca_color_size_ext.argtypes = [ctypes.c_char_p, i]
    L.fixca_params_default.argtypes = [pp]
    L.fixca_params_default.restype = None
    L.fixca_cuda_set_progress.argtypes = [PROGRESS_FN, vp]
    L.fixca_cuda_set_progress.restype = None
    L.fixca_cuda_last_error.restype = ctypes.c_char_p
    L.fixca_strerror.argtypes = [i]
    L.fixca_strerror.restype = ctypes.c_char_p
    L.fixca_cuda_last_kernel.restype = ctypes.c_char_p
    L.fixca_cuda_launch_count.restype = ctypes.c_long
    L.fixca_cuda_release.restype = None
    L.fixca_version.restype = ctypes.c_char_p
    L.fixca_cuda_host_alloc.argtypes = [ctypes.c_size_t]
    L.fixca_cuda_host_alloc.restype = vp
    L.fixca_cuda_host_free.argtypes = [vp]
    L.fixca_cuda_host_free.restype = None
    L.fixca_cuda_reload_tuning.restype = None
    L.fixca_cuda_last_call_ms.restype = ctypes.c_double
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != OK:
        raise FixCaError(rc, load().fixca_cuda_last_error().decode() or load().fixca_strerror(rc).decode())


def bpc_of(dtype) -> int:
    """The reference's bpc code (fix-ca.c:688-707) for a numpy dtype."""
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return -dt.itemsize
    if dt.kind == "u":
        return dt.itemsize
    raise ValueError("unsupported dtype %s" % dt)


def fix_ca_region(src, dst, orig_width, orig_height, bytes, bpc, params, x1, x2, y1, y2, show_progress=True,
                  flags=None, device=-1):
    """fix_ca_region() of the reference (fix-ca.c:998-1001) on the GPU.

    ``src`` / ``dst`` are host buffers (numpy arrays or integer addresses) of
    ``orig_width * orig_height * bytes`` bytes; rows ``[y1, y2)`` of dst are written.
    ``flags`` None means the library default (exact FP64 arithmetic)."""
    L = load()
    s = src.ctypes.data if isinstance(src, np.ndarray) else int(src)
    d = dst.ctypes.data if isinstance(dst, np.ndarray) else int(dst)
    if flags is None:
        rc = L.fixca_cuda_region(s, d, orig_width, orig_height, bytes, bpc, ctypes.byref(params),
                                 x1, x2, y1, y2, 1 if show_progress else 0)
    else:
        rc = L.fixca_cuda_region_ex(s, d, orig_width, orig_height, bytes, bpc, ctypes.byref(params),
                                    x1, x2, y1, y2, 1 if show_progress else 0, flags, device)
    _check(rc)


def preview(image: np.ndarray, params: FixCaParams, x: int, y: int, pw: int, ph: int, bpc=None) -> np.ndarray:
    """preview_update()'s middle (fix-ca.c:656-671): the 8-bit preview buffer of window (x, y, pw, ph), shape
    (ph, pw, channels): correction + saturation boost + centre lines + get_pixel / set_pixel(.., 1)."""
    h, w, ch = image.shape
    out = np.zeros((ph, pw, ch), dtype=np.uint8)
    _check(load().fixca_cuda_preview(image.ctypes.data, out.ctypes.data, w, h, ch * image.dtype.itemsize,
                                     bpc_of(image.dtype) if bpc is None else bpc, ctypes.byref(params), x, y, pw, ph))
    return out


def correct(image: np.ndarray, params: FixCaParams, y1=None, y2=None, out=None, flags=PRECISION_EXACT,
            device=-1, devices=None, bpc=None) -> np.ndarray:
    """Array-level convenience: (H, W, C) uint8/16/32/64/float16/32/64 image in, corrected image out.
    ``bpc`` overrides the code derived from the dtype (BPC_U15 for 15-bit samples held in uint16)."""
    assert image.ndim == 3 and image.flags["C_CONTIGUOUS"]
    h, w, ch = image.shape
    y1 = 0 if y1 is None else y1
    y2 = h if y2 is None else y2
    if out is None:
        out = np.zeros_like(image)
    bytes_ = ch * image.dtype.itemsize
    bpc = bpc_of(image.dtype) if bpc is None else bpc
    if devices is not None:
        arr = (ctypes.c_int * len(devices))(*devices)
        _check(load().fixca_cuda_region_multi(image.ctypes.data, out.ctypes.data, w, h, bytes_, bpc,
                                              ctypes.byref(params), y1, y2, flags, arr, len(devices)))
    else:
        fix_ca_region(image, out, w, h, bytes_, bpc, params, 0, w, y1, y2, True, flags, device)
    return out


def fix_ca_region_dev(d_src: int, src_pitch: int, src_row0: int, src_rows: int, d_dst: int, dst_pitch: int,
                      dst_row0: int, width: int, height: int, bytes: int, bpc: int, params: FixCaParams,
                      y1: int, y2: int, flags: int = PRECISION_EXACT, stream: int = 0) -> None:
    """Device-resident pass (raw CUDA pointers, asynchronous on ``stream``)."""
    _check(load().fixca_cuda_region_dev(d_src, src_pitch, src_row0, src_rows, d_dst, dst_pitch, dst_row0,
                                        width, height, bytes, bpc, ctypes.byref(params), y1, y2, flags, stream))


def fix_ca_region_dev_fanout(d_src: int, src_pitch: int, src_row0: int, src_rows: int, d_dsts, dst_pitch: int,
                             dst_row0: int, width: int, height: int, bytes: int, bpc: int, params: FixCaParams,
                             y1: int, y2: int, flags: int = PRECISION_EXACT, stream: int = 0) -> None:
    """The device-resident pass with several destination frames (raw CUDA pointers, same pitch and first row): every
    finished chunk is stored into each of them by the same launch (all-gather form of the reassembly)."""
    arr = (ctypes.c_void_p * len(d_dsts))(*d_dsts)
    _check(load().fixca_cuda_region_dev_fanout(d_src, src_pitch, src_row0, src_rows, arr, len(d_dsts), dst_pitch, dst_row0,
                                               width, height, bytes, bpc, ctypes.byref(params), y1, y2, flags, stream))


def fix_ca_frames_dev(d_src: int, src_pitch: int, src_frame_stride: int, d_dst: int, dst_pitch: int,
                      dst_frame_stride: int, nframes: int, width: int, height: int, bytes: int, bpc: int,
                      params: FixCaParams, flags: int = PRECISION_EXACT, stream: int = 0) -> None:
    """A batch of equal device-resident frames (raw CUDA pointers, frame i at base + i * frame_stride), one launch
    for the whole batch on the streaming kernels; asynchronous on ``stream``."""
    _check(load().fixca_cuda_frames_dev(d_src, src_pitch, src_frame_stride, d_dst, dst_pitch, dst_frame_stride, nframes,
                                        width, height, bytes, bpc, ctypes.byref(params), flags, stream))


def correct_frames(frames, params: FixCaParams, flags=PRECISION_EXACT, device=-1, devices=None, outs=None):
    """A stream of equal-shaped host frames through the pinned H2D / kernel / D2H ring of one device, or
    (``devices``) sharded by index over several GPUs from this process."""
    if not frames:
        return []
    h, w, ch = frames[0].shape
    if outs is None:
        outs = [np.zeros_like(f) for f in frames]
    n = len(frames)
    src = (ctypes.c_void_p * n)(*[f.ctypes.data for f in frames])
    dst = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
    if devices is not None:
        arr = (ctypes.c_int * len(devices))(*devices)
        _check(load().fixca_cuda_frames_multi(src, dst, n, w, h, ch * frames[0].dtype.itemsize, bpc_of(frames[0].dtype),
                                              ctypes.byref(params), flags, arr, len(devices)))
    else:
        _check(load().fixca_cuda_frames(src, dst, n, w, h, ch * frames[0].dtype.itemsize, bpc_of(frames[0].dtype),
                                        ctypes.byref(params), flags, device))
    return outs


def frame_alloc(nbytes: int):
    """A destination frame on the current GPU that other ranks can map (fixca_cuda_frame_alloc):
    returns (device pointer, 64-byte CUDA IPC handle)."""
    ptr = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
    _check(load().fixca_cuda_frame_alloc(nbytes, ctypes.byref(ptr), handle))
    return int(ptr.value), handle.raw


def frame_open(handle: bytes) -> int:
    """Map another rank's frame into this process (peer access over NVLink); returns the device pointer."""
    ptr = ctypes.c_void_p()
    _check(load().fixca_cuda_frame_open(ctypes.create_string_buffer(bytes(handle), IPC_HANDLE_BYTES), ctypes.byref(ptr)))
    return int(ptr.value)


def frame_close(ptr: int) -> None:
    _check(load().fixca_cuda_frame_close(ptr))


def frame_free(ptr: int) -> None:
    _check(load().fixca_cuda_frame_free(ptr))


def band_source_rows(width: int, height: int, params: FixCaParams, y1: int, y2: int):
    lo, hi = ctypes.c_int(), ctypes.c_int()
    _check(load().fixca_band_source_rows(width, height, ctypes.byref(params), y1, y2, ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


def split_bands(y1: int, y2: int, nbands: int):
    a, b = (ctypes.c_int * nbands)(), (ctypes.c_int * nbands)()
    _check(load().fixca_split_bands(y1, y2, nbands, a, b))
    return list(zip(list(a), list(b)))


def resolve_lens(width: int, height: int, lens_x: float, lens_y: float):
    a, b = ctypes.c_double(lens_x), ctypes.c_double(lens_y)
    load().fixca_resolve_lens(width, height, ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def check_params(params: FixCaParams) -> int:
    return load().fixca_check_params(ctypes.byref(params))


def color_size(format_name: str, bytes_per_pixel: int) -> int:
    return load().fixca_color_size(format_name.encode(), bytes_per_pixel)


def color_size_half(format_name: str, bytes_per_pixel: int) -> int:
    """color_size() with the reference's commented-out half-precision line enabled (bpc -2 for float16)."""
    return load().fixca_color_size_half(format_name.encode(), bytes_per_pixel)


def color_size_ext(format_name: str, bytes_per_pixel: int) -> int:
    """color_size() with both of the reference's "TODO for another day" formats answered (half -2, u15 15)."""
    return load().fixca_color_size_ext(format_name.encode(), bytes_per_pixel)


def device_count() -> int:
    return load().fixca_cuda_device_count()


def last_kernel() -> str:
    return load().fixca_cuda_last_kernel().decode()


def last_call_ms() -> float:
    """Wall-clock milliseconds of the last fixca_cuda_region*() host call on this thread."""
    return float(load().fixca_cuda_last_call_ms())


def launch_count() -> int:
    return load().fixca_cuda_launch_count()


def reload_tuning() -> None:
    """Re-read the FIXCA_* tuning variables (they are read once per process)."""
    load().fixca_cuda_reload_tuning()


class PinnedBuffer:
    """Pinned host memory from fixca_cuda_host_alloc (what the patched plug-in uses in place of g_new,
    fix-ca.c:366-367): ``array(dtype, shape)`` views it as numpy; ``free()`` releases it."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = load().fixca_cuda_host_alloc(self.nbytes)
        if not self.ptr:
            raise FixCaError(ERR_NOMEM, load().fixca_cuda_last_error().decode() or "fixca_cuda_host_alloc failed")

    def array(self, dtype=np.uint8, shape=None) -> np.ndarray:
        buf = (ctypes.c_ubyte * self.nbytes).from_address(self.ptr)
        a = np.frombuffer(buf, dtype=dtype)
        return a if shape is None else a.reshape(shape)

    def free(self) -> None:
        if self.ptr:
            load().fixca_cuda_host_free(self.ptr)
            self.ptr = None


_progress_keepalive = None


def set_progress(fn) -> None:
    """Install fn(kind, fraction) as the progress callback (None removes it)."""
    global _progress_keepalive
    if fn is None:
        _progress_keepalive = ctypes.cast(None, PROGRESS_FN)
    else:
        _progress_keepalive = PROGRESS_FN(lambda kind, frac, user: fn(kind, frac))
    load().fixca_cuda_set_progress(_progress_keepalive, None)
