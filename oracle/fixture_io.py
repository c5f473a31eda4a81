"""oracle/fixture_io.py -- TEST INFRASTRUCTURE ONLY.

File-format ends of the reference's one known-answer test
(/root/reference/tests/Makefile.am:13-22, tests/test-fix-ca.scm:1-6): GIMP loads
img-fix-ca/full-branches.jpg with libjpeg's *float* DCT, runs the plug-in, and
saves a 24-bit BMP whose md5 is tests/test1.md5.  Neither codec lives in the
reference tree; both are restated here and are verified
only by the md5 itself.
"""
from __future__ import annotations

import ctypes
import glob
import os
import struct

import numpy as np

JDCT_FLOAT = 2
_JPEG_LIB_VERSION = 62
_SIZEOF_DECOMPRESS = 632          # libjpeg-turbo, v62 ABI, x86-64
_OFF_DCT_METHOD = 96
_OFF_OUTPUT_WIDTH = 136
_OFF_OUTPUT_HEIGHT = 140
_OFF_OUTPUT_COMPONENTS = 148


def _find_libjpeg() -> str:
    import PIL

    pat = os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libjpeg-*.so*")
    hits = sorted(glob.glob(pat))
    if not hits:
        raise FileNotFoundError("Pillow's bundled libjpeg not found: " + pat)
    return hits[0]


def decode_jpeg_float_dct(path: str) -> np.ndarray:
    """Decode a JPEG the way GIMP 2.10's file-jpeg does (dct_method = JDCT_FLOAT).

    Returns an (H, W, 3) uint8 array.  PIL/OpenCV use JDCT_ISLOW and differ in
    ~1.5 % of samples, which does not reproduce tests/test1.md5.
    """
    lib = ctypes.CDLL(_find_libjpeg())
    data = open(path, "rb").read()
    cinfo = ctypes.create_string_buffer(_SIZEOF_DECOMPRESS)
    jerr = ctypes.create_string_buffer(512)

    lib.jpeg_std_error.restype = ctypes.c_void_p
    lib.jpeg_std_error.argtypes = [ctypes.c_void_p]
    lib.jpeg_CreateDecompress.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
    lib.jpeg_mem_src.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_ulong]
    lib.jpeg_read_header.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.jpeg_start_decompress.argtypes = [ctypes.c_void_p]
    lib.jpeg_read_scanlines.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint]
    lib.jpeg_read_scanlines.restype = ctypes.c_uint
    lib.jpeg_finish_decompress.argtypes = [ctypes.c_void_p]
    lib.jpeg_destroy_decompress.argtypes = [ctypes.c_void_p]

    err_ptr = lib.jpeg_std_error(ctypes.addressof(jerr))
    ctypes.c_void_p.from_address(ctypes.addressof(cinfo)).value = err_ptr   # cinfo.err
    lib.jpeg_CreateDecompress(ctypes.addressof(cinfo), _JPEG_LIB_VERSION, _SIZEOF_DECOMPRESS)
    lib.jpeg_mem_src(ctypes.addressof(cinfo), data, len(data))
    lib.jpeg_read_header(ctypes.addressof(cinfo), 1)
    ctypes.c_int.from_address(ctypes.addressof(cinfo) + _OFF_DCT_METHOD).value = JDCT_FLOAT
    lib.jpeg_start_decompress(ctypes.addressof(cinfo))
    w = ctypes.c_uint.from_address(ctypes.addressof(cinfo) + _OFF_OUTPUT_WIDTH).value
    h = ctypes.c_uint.from_address(ctypes.addressof(cinfo) + _OFF_OUTPUT_HEIGHT).value
    c = ctypes.c_int.from_address(ctypes.addressof(cinfo) + _OFF_OUTPUT_COMPONENTS).value
    out = np.empty((h, w, c), dtype=np.uint8)
    row_ptr = (ctypes.c_void_p * 1)()
    y = 0
    while y < h:
        row_ptr[0] = out.ctypes.data + y * w * c
        n = lib.jpeg_read_scanlines(ctypes.addressof(cinfo), row_ptr, 1)
        if n != 1:
            raise RuntimeError("jpeg_read_scanlines failed at row %d" % y)
        y += 1
    lib.jpeg_finish_decompress(ctypes.addressof(cinfo))
    lib.jpeg_destroy_decompress(ctypes.addressof(cinfo))
    return out


def encode_gimp_bmp24(rgb: np.ndarray) -> bytes:
    """Serialise (H, W, 3) uint8 RGB as GIMP 2.10 file-bmp writes it
    non-interactively: 14-byte file header + 124-byte BITMAPV5HEADER, BI_RGB,
    72 dpi (2835 px/m), sRGB colour space, rows bottom-up in BGR order."""
    h, w, c = rgb.shape
    assert c == 3 and rgb.dtype == np.uint8
    stride = (w * 3 + 3) & ~3
    image_size = stride * h
    offset = 14 + 124
    hdr = struct.pack("<2sIHHI", b"BM", offset + image_size, 0, 0, offset)
    v5 = struct.pack(
        "<IiiHHIIiiII" "IIII" "I" "36s" "III" "IIII",
        124, w, h, 1, 24, 0, image_size, 2835, 2835, 0, 0,
        0x00FF0000, 0x0000FF00, 0x000000FF, 0x00000000,
        0x73524742,                    # 'sRGB'
        b"\0" * 36,                    # endpoints
        0, 0, 0,                       # gamma r/g/b
        2, 0, 0, 0,                    # intent, profile data, profile size, reserved
    )
    assert len(v5) == 124
    rows = np.zeros((h, stride), dtype=np.uint8)
    rows[:, : w * 3] = rgb[::-1, :, ::-1].reshape(h, w * 3)
    return hdr + v5 + rows.tobytes()
