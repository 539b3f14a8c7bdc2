"""basisu_rs_b200 -- host-side mirror of the reference crate's public API over the C ABI.

The product is ``libb2bu.so`` (hand-written sm_100a CUDA kernels behind ``include/b2bu.h``).  This
module is the thin Python host layer used by the tests and ``bench.py``; it mirrors the names,
argument meaning and error behaviour of the reference's ``src/lib.rs:20-79``:

    read_to_rgba / read_to_etc1 / read_to_etc2 / read_to_uastc / read_to_astc / read_to_bc7
    unpack_uastc_block_to_rgba / transcode_uastc_block_to_{astc,bc7,etc1,etc2}
    Image(w, h, stride, data), Header(...26 fields...), errors as BasisuError(message)

There is no CPU fallback: if the shared library is missing or CUDA is unusable the calls raise.
"""
from __future__ import annotations

import ctypes
import dataclasses
import pathlib
from typing import List, Tuple

__all__ = [
    "BasisuError", "Header", "Image", "lib", "library_path",
    "read_to_rgba", "read_to_etc1", "read_to_etc2", "read_to_uastc", "read_to_astc", "read_to_bc7",
    "unpack_uastc_block_to_rgba", "transcode_uastc_block_to_astc", "transcode_uastc_block_to_bc7",
    "transcode_uastc_block_to_etc1", "transcode_uastc_block_to_etc2",
    "uastc_transcode", "uastc_decode_rgba", "Etc1sDecoder",
    "RGBA", "ASTC", "BC7", "ETC1", "ETC2", "UASTC",
]

RGBA, ASTC, BC7, ETC1, ETC2, UASTC, BC1 = 0, 1, 2, 3, 4, 5, 6
BLOCK_BYTES = {RGBA: 64, ASTC: 16, BC7: 16, ETC1: 8, ETC2: 16, UASTC: 16}

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None


class BasisuError(Exception):
    """Mirrors the reference's ``Error = String`` (src/lib.rs:26): str(e) is the message."""

    def __init__(self, status: int, message: str, first_bad_block: int | None = None):
        super().__init__(message)
        self.status = status
        self.first_bad_block = first_bad_block


def library_path() -> pathlib.Path:
    import os
    override = os.environ.get("B2BU_LIBRARY")          # tuning variants built by basisu_rs_b200.build --out=...
    return pathlib.Path(override) if override else _HERE / "libb2bu.so"


def lib() -> ctypes.CDLL:
    """Loads libb2bu.so (built in-tree by ``python -m basisu_rs_b200.build``).  Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists():
        raise ImportError(
            f"{path} is missing: build the CUDA extension first (python -m basisu_rs_b200.build). "
            "basisu_rs_b200 has no CPU fallback.")
    L = ctypes.CDLL(str(path))
    c = ctypes
    u8p, sz, u64p = c.c_void_p, c.c_size_t, c.POINTER(c.c_uint64)
    L.b2bu_error_string.restype = c.c_char_p
    L.b2bu_error_string.argtypes = [c.c_int]
    L.b2bu_last_cuda_error.restype = c.c_char_p
    L.b2bu_init.argtypes = [c.c_int]
    L.b2bu_device_count.argtypes = [c.POINTER(c.c_int)]
    L.b2bu_block_bytes.restype = sz
    L.b2bu_block_bytes.argtypes = [c.c_int]
    L.b2bu_host_alloc.restype = c.c_void_p
    L.b2bu_host_alloc.argtypes = [sz]
    L.b2bu_host_free.argtypes = [c.c_void_p]
    for name in ("b2bu_unpack_uastc_block_to_rgba", "b2bu_transcode_uastc_block_to_astc", "b2bu_transcode_uastc_block_to_bc7",
                 "b2bu_transcode_uastc_block_to_etc1", "b2bu_transcode_uastc_block_to_etc2"):
        getattr(L, name).argtypes = [u8p, u8p]
    L.b2bu_uastc_transcode.argtypes = [c.c_int, u8p, sz, u8p, sz, u64p]
    L.b2bu_uastc_decode_rgba.argtypes = [u8p, sz, sz, u8p, sz, u64p]
    L.b2bu_uastc_transcode_dev.argtypes = [c.c_int, c.c_void_p, sz, sz, c.c_void_p, sz, c.c_void_p, c.c_void_p]
    L.b2bu_uastc_transcode_slices_dev.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p, c.c_void_p]
    L.b2bu_crc16_dev.argtypes = [c.c_void_p, sz, c.c_uint16, c.POINTER(c.c_uint16), c.c_void_p]
    L.b2bu_etc1s_table_info.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
    L.b2bu_status_reset_dev.argtypes = [c.c_void_p, c.c_void_p]
    L.b2bu_status_read_dev.argtypes = [c.c_void_p, c.c_void_p, u64p]
    L.b2bu_probe_int_peak.argtypes = [c.POINTER(c.c_double), c.POINTER(c.c_double)]
    L.b2bu_launch_count.restype = c.c_uint64
    L.b2bu_etc1s_open.argtypes = [c.c_uint32, c.c_uint32, u8p, sz, u8p, sz, u8p, sz, c.c_int, c.POINTER(c.c_void_p)]
    L.b2bu_etc1s_close.argtypes = [c.c_void_p]
    L.b2bu_etc1s_transcode_to_etc1.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, u8p, sz, u8p, sz]
    L.b2bu_etc1s_transcode_to_bc1.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, u8p, sz, u8p, sz]
    L.b2bu_etc1s_decode_to_rgba.argtypes = [c.c_void_p, c.c_uint32, c.c_uint32, u8p, sz, u8p, sz, u8p, sz]
    L.b2bu_etc1s_transcode_slices.argtypes = [c.c_void_p, c.c_int, c.c_uint32, c.c_uint32, u8p, sz, u64p, u64p, c.c_uint32, u8p, sz]
    L.b2bu_etc1s_last_timing.argtypes = [c.c_void_p, c.POINTER(c.c_float), c.POINTER(c.c_float), c.POINTER(c.c_float), u64p]
    L.b2bu_read_header.argtypes = [u8p, sz, c.c_void_p]
    L.b2bu_crc16.restype = c.c_uint16
    L.b2bu_crc16.argtypes = [u8p, sz, c.c_uint16]
    L.b2bu_read_to.argtypes = [c.c_int, u8p, sz, c.c_void_p, c.c_void_p, c.c_uint32, c.POINTER(c.c_uint32), u8p, c.c_uint64, u64p]
    L.b2bu_read_to_flags.argtypes = L.b2bu_read_to.argtypes + [c.c_uint32]
    _LIB = L
    return L


def _check(status: int, first_bad: int | None = None) -> None:
    if status == 0:
        return
    L = lib()
    msg = L.b2bu_error_string(status).decode()
    if status == 18:
        msg += ": " + L.b2bu_last_cuda_error().decode()
    raise BasisuError(status, msg, first_bad)


def _buf(data) -> Tuple[ctypes.Array, int]:
    b = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    arr = (ctypes.c_uint8 * max(len(b), 1)).from_buffer_copy(b if len(b) else b"\0")
    return arr, len(b)


# ---- single-block API: src/lib.rs:29-53 ---------------------------------------------------------

def _block(fn_name: str, data, out_bytes: int) -> bytes:
    if len(data) != 16:
        raise ValueError("a UASTC block is 16 bytes")     # the reference takes [u8; 16]
    src, _ = _buf(data)
    out = (ctypes.c_uint8 * out_bytes)()
    _check(getattr(lib(), fn_name)(src, out))
    return bytes(out)


def unpack_uastc_block_to_rgba(data) -> List[int]:
    """lib.rs:29 -- 16 pixels as 0xAABBGGRR u32, raster order."""
    raw = _block("b2bu_unpack_uastc_block_to_rgba", data, 64)
    return [int.from_bytes(raw[4 * i:4 * i + 4], "little") for i in range(16)]


def transcode_uastc_block_to_astc(data) -> bytes:
    return _block("b2bu_transcode_uastc_block_to_astc", data, 16)


def transcode_uastc_block_to_bc7(data) -> bytes:
    return _block("b2bu_transcode_uastc_block_to_bc7", data, 16)


def transcode_uastc_block_to_etc1(data) -> bytes:
    return _block("b2bu_transcode_uastc_block_to_etc1", data, 8)


def transcode_uastc_block_to_etc2(data) -> bytes:
    return _block("b2bu_transcode_uastc_block_to_etc2", data, 16)


# ---- slice level: uastc::Decoder (src/uastc.rs:77-165) ------------------------------------------

def uastc_transcode(target: int, data) -> bytes:
    """Decoder::transcode (uastc.rs:112): data = whole UASTC slice, returns the transcoded slice."""
    src, n = _buf(data)
    out_len = (n // 16) * BLOCK_BYTES[target]
    out = (ctypes.c_uint8 * max(out_len, 1))()
    bad = ctypes.c_uint64(0)
    st = lib().b2bu_uastc_transcode(target, src, n, out, out_len, ctypes.byref(bad))
    _check(st, bad.value)
    return bytes(out)[:out_len]


def uastc_decode_rgba(data, blocks_per_row: int) -> bytes:
    """Decoder::decode_to_rgba (uastc.rs:89) + Color32::into_rgba_bytes: RGBA bytes, row-major."""
    src, n = _buf(data)
    px = (n // 16) * 16
    out = (ctypes.c_uint8 * max(px * 4, 1))()
    bad = ctypes.c_uint64(0)
    st = lib().b2bu_uastc_decode_rgba(src, n, blocks_per_row, out, px, ctypes.byref(bad))
    _check(st, bad.value)
    return bytes(out)[:px * 4]


# ---- ETC1S: basis_lz::Decoder (src/basis_lz/mod.rs:50-186) --------------------------------------

class SliceDev(ctypes.Structure):
    """b2bu_slice_dev: one slice of a device buffer for b2bu_uastc_transcode_slices_dev."""
    _fields_ = [("in_ofs", ctypes.c_uint64), ("out_ofs", ctypes.c_uint64), ("nblocks", ctypes.c_uint64),
                ("blocks_per_row", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class Etc1sDecoder:
    """basis_lz::Decoder: ``new`` decodes the codebooks / Huffman models once per file."""

    def __init__(self, endpoint_count, selector_count, endpoints_data, selector_data, tables_data, extended_data=b"", is_video=False):
        ep, epn = _buf(endpoints_data)
        se, sen = _buf(selector_data)
        tb, tbn = _buf(tables_data)
        self._h = ctypes.c_void_p()
        _check(lib().b2bu_etc1s_open(endpoint_count, selector_count, ep, epn, se, sen, tb, tbn, int(bool(is_video)), ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().b2bu_etc1s_close(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def transcode_to_etc1(self, num_blocks_x, num_blocks_y, block_data) -> bytes:
        src, n = _buf(block_data)
        out_len = num_blocks_x * num_blocks_y * 8
        out = (ctypes.c_uint8 * max(out_len, 1))()
        _check(lib().b2bu_etc1s_transcode_to_etc1(self._h, num_blocks_x, num_blocks_y, src, n, out, out_len))
        return bytes(out)[:out_len]

    def transcode_to_bc1(self, num_blocks_x, num_blocks_y, block_data) -> bytes:
        """EXTENSION (the reference has no BC1): see b2bu_etc1s_transcode_to_bc1 in include/b2bu.h."""
        src, n = _buf(block_data)
        out_len = num_blocks_x * num_blocks_y * 8
        out = (ctypes.c_uint8 * max(out_len, 1))()
        _check(lib().b2bu_etc1s_transcode_to_bc1(self._h, num_blocks_x, num_blocks_y, src, n, out, out_len))
        return bytes(out)[:out_len]

    def decode_to_rgba(self, num_blocks_x, num_blocks_y, rgb_data, alpha_data=None) -> bytes:
        src, n = _buf(rgb_data)
        if alpha_data is not None:
            al, an = _buf(alpha_data)
        else:
            al, an = None, 0
        out_len = num_blocks_x * num_blocks_y * 64
        out = (ctypes.c_uint8 * max(out_len, 1))()
        _check(lib().b2bu_etc1s_decode_to_rgba(self._h, num_blocks_x, num_blocks_y, src, n, al, an, out, out_len))
        return bytes(out)[:out_len]


# ---- file level: src/basis.rs, src/lib.rs:63-79 --------------------------------------------------

_HEADER_FIELDS = ["sig", "ver", "header_size", "header_crc16", "data_size", "data_crc16", "total_slices", "total_images",
                  "tex_format", "flags", "tex_type", "us_per_frame", "reserved", "userdata0", "userdata1", "total_endpoints",
                  "endpoint_cb_file_ofs", "endpoint_cb_file_size", "total_selectors", "selector_cb_file_ofs",
                  "selector_cb_file_size", "tables_file_ofs", "tables_file_size", "slice_desc_file_ofs", "extended_file_ofs",
                  "extended_file_size"]


class _CHeader(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in _HEADER_FIELDS]


class _CImage(ctypes.Structure):
    _fields_ = [("w", ctypes.c_uint32), ("h", ctypes.c_uint32), ("stride", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
                ("offset", ctypes.c_uint64), ("nbytes", ctypes.c_uint64)]


@dataclasses.dataclass
class Header:
    """basis.rs:419-454 (all public fields) + the three helper methods :459-473."""
    sig: int = 0
    ver: int = 0
    header_size: int = 0
    header_crc16: int = 0
    data_size: int = 0
    data_crc16: int = 0
    total_slices: int = 0
    total_images: int = 0
    tex_format: int = 0
    flags: int = 0
    tex_type: int = 0
    us_per_frame: int = 0
    reserved: int = 0
    userdata0: int = 0
    userdata1: int = 0
    total_endpoints: int = 0
    endpoint_cb_file_ofs: int = 0
    endpoint_cb_file_size: int = 0
    total_selectors: int = 0
    selector_cb_file_ofs: int = 0
    selector_cb_file_size: int = 0
    tables_file_ofs: int = 0
    tables_file_size: int = 0
    slice_desc_file_ofs: int = 0
    extended_file_ofs: int = 0
    extended_file_size: int = 0

    def has_alpha(self) -> bool:
        return (self.flags & 4) != 0

    def has_y_flipped(self) -> bool:
        return (self.flags & 2) != 0

    def texture_format(self) -> str:
        if self.tex_format == 0:
            return "ETC1S"
        if self.tex_format == 1:
            return "UASTC4x4"
        raise BasisuError(13, "Unknown texture format")


@dataclasses.dataclass
class Image:
    """lib.rs:63-68 Image<u8>."""
    w: int
    h: int
    stride: int
    data: bytes


def read_header(buf) -> Header:
    src, n = _buf(buf)
    ch = _CHeader()
    _check(lib().b2bu_read_header(src, n, ctypes.byref(ch)))
    return Header(**{f: getattr(ch, f) for f in _HEADER_FIELDS})


def _read_to(target: int, buf, apply_y_flip: bool = False) -> Tuple[Header, List[Image]]:
    L = lib()
    src, n = _buf(buf)
    ch = _CHeader()
    count = ctypes.c_uint32(0)
    need = ctypes.c_uint64(0)
    flags = 1 if apply_y_flip else 0                      # B2BU_READ_APPLY_Y_FLIP
    _check(L.b2bu_read_to_flags(target, src, n, ctypes.byref(ch), None, 0, ctypes.byref(count), None, 0, ctypes.byref(need), flags))
    imgs = (_CImage * max(count.value, 1))()
    out = (ctypes.c_uint8 * max(need.value, 1))()
    _check(L.b2bu_read_to_flags(target, src, n, ctypes.byref(ch), imgs, count.value, ctypes.byref(count), out, need.value, ctypes.byref(need), flags))
    raw = memoryview(out)
    images = [Image(im.w, im.h, im.stride, bytes(raw[im.offset:im.offset + im.nbytes])) for im in imgs[:count.value]]
    return Header(**{f: getattr(ch, f) for f in _HEADER_FIELDS}), images


def read_to_rgba(buf, apply_y_flip: bool = False) -> Tuple[Header, List[Image]]:
    """basis.rs:8 -- returns (Header, images) like the reference.  apply_y_flip (opt-in, not in the reference's signature):
    honour Header::has_y_flipped the way the reference's own tests do (tests/common.rs:284-301 rgba_rows)."""
    return _read_to(RGBA, buf, apply_y_flip)


def rgba_rows(image: "Image", y_flipped: bool) -> List[bytes]:
    """tests/common.rs:284-301 of the reference: the first h rows of `stride` bytes, trimmed to w pixels, reversed when flipped"""
    rows = [bytes(image.data[y * image.stride: y * image.stride + 4 * image.w]) for y in range(image.h)]
    return rows[::-1] if y_flipped else rows


def read_to_etc1(buf) -> List[Image]:
    return _read_to(ETC1, buf)[1]


def read_to_etc2(buf) -> List[Image]:
    return _read_to(ETC2, buf)[1]


def read_to_uastc(buf) -> List[Image]:
    return _read_to(UASTC, buf)[1]


def read_to_astc(buf) -> List[Image]:
    return _read_to(ASTC, buf)[1]


def read_to_bc7(buf) -> List[Image]:
    return _read_to(BC7, buf)[1]
