"""The C-ABI library must load and export every symbol include/b2bu.h declares (no compute here)."""
import ctypes
import pathlib
import re

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "b2bu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2bu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import basisu_rs_b200 as b
    path = b.library_path()
    assert path.exists(), "run `python -m basisu_rs_b200.build` (or __graft_entry__.build()) first"
    L = ctypes.CDLL(str(path))
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b2bu.h but not exported"


def test_rust_shim_binds_exactly_the_declared_symbols():
    """rust/src/ffi.rs is the binding a maintainer of the reference crate would add; it cannot be compiled here
    (no rustc), so at least its extern block must name exactly the functions include/b2bu.h declares, with as many arguments."""
    hdr = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "b2bu.h").read_text(), flags=re.S)
    rs = re.sub(r"//.*", "", (ROOT / "rust" / "src" / "ffi.rs").read_text())
    c_args = {m.group(1): len([a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"])
              for m in re.finditer(r"\b(b2bu_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)}
    rs_args = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
               for m in re.finditer(r"pub fn (b2bu_[a-z0-9_]+)\s*\(([^)]*)\)", rs)}
    assert set(c_args) == set(declared_symbols())
    assert set(rs_args) == set(c_args), (sorted(set(c_args) - set(rs_args)), sorted(set(rs_args) - set(c_args)))
    assert rs_args == c_args


def test_host_only_entry_points_work_without_gpu():
    import basisu_rs_b200 as b
    from basis_writer import uastc_file, crc16
    L = b.lib()
    assert L.b2bu_block_bytes(0) == 64 and L.b2bu_block_bytes(3) == 8 and L.b2bu_block_bytes(2) == 16
    assert L.b2bu_error_string(2) == b"invalid mode index"
    assert L.b2bu_error_string(3) == b"block pattern is not valid"
    assert L.b2bu_error_string(1) == b"data length is not divisible by UASTC block size (16)"
    d = bytes(range(200))
    assert L.b2bu_crc16(d, len(d), 0) == crc16(d)
    blocks = np.zeros((6, 16), dtype=np.uint8).tobytes()
    f = uastc_file(blocks, 3, 2)
    h = b.read_header(f)
    assert (h.sig, h.header_size, h.total_slices, h.tex_format) == (0x4273, 77, 1, 1)
    assert h.texture_format() == "UASTC4x4" and not h.has_alpha()
    # sizing call (out == NULL) does not touch the GPU
    cnt = ctypes.c_uint32(0)
    need = ctypes.c_uint64(0)
    imgs = (b._CImage * 4)()
    st = L.b2bu_read_to(b.BC7, f, len(f), None, imgs, 4, ctypes.byref(cnt), None, 0, ctypes.byref(need))
    assert (st, cnt.value, need.value) == (0, 1, 6 * 16)
    assert (imgs[0].w, imgs[0].h, imgs[0].stride) == (12, 8, 48)
    for corrupt, code in (("sig", 9), ("header", 11), ("data", 12)):
        g = uastc_file(blocks, 3, 2, corrupt=corrupt)
        assert L.b2bu_read_to(b.BC7, g, len(g), None, None, 0, ctypes.byref(cnt), None, 0, ctypes.byref(need)) == code
    assert L.b2bu_read_to(b.BC7, f, 40, None, None, 0, ctypes.byref(cnt), None, 0, ctypes.byref(need)) == 10


def test_sizing_call_of_a_large_file_leaves_the_data_crc_to_the_decoding_call():
    """Files of 256 KiB and more are CRC-checked by the GPU inside the decoding call; the host-only sizing call (out == NULL) must
    not spend a single-core CRC pass on them (and so cannot report a data CRC error), while small files keep the host check."""
    import basisu_rs_b200 as b
    from basis_writer import uastc_file
    L = b.lib()
    cnt, need = ctypes.c_uint32(0), ctypes.c_uint64(0)
    big = np.zeros((160 * 128, 16), dtype=np.uint8).tobytes()          # 320 KiB of (valid mode 11) zero blocks
    g = uastc_file(big, 160, 128, corrupt="data")
    assert len(g) >= 256 * 1024
    assert L.b2bu_read_to(b.BC7, g, len(g), None, None, 0, ctypes.byref(cnt), None, 0, ctypes.byref(need)) == 0
    assert (cnt.value, need.value) == (1, 160 * 128 * 16)
    # read_to_uastc (a plain copy, never on the device) still verifies on the host, sizing call aside
    out = (ctypes.c_uint8 * need.value)()
    imgs = (b._CImage * 1)()
    assert L.b2bu_read_to(b.UASTC, g, len(g), None, imgs, 1, ctypes.byref(cnt), out, need.value, ctypes.byref(need)) == 12


def _patch_header(f: bytes, **fields) -> bytes:
    """rewrites header fields of a .basis file and re-signs the header CRC (basis.rs:419-454 layout)"""
    import struct
    from basis_writer import _crc16_fast
    g = bytearray(f)
    if "total_slices" in fields:
        g[14:17] = struct.pack("<I", fields["total_slices"])[:3]
    if "slice_desc_file_ofs" in fields:
        g[65:69] = struct.pack("<I", fields["slice_desc_file_ofs"])
    g[6:8] = struct.pack("<H", _crc16_fast(bytes(g[8:77])))
    return bytes(g)


def test_sizing_call_of_a_large_file_defers_body_errors_behind_the_crc():
    """basis.rs:9-13 checks the data CRC before the slice table is looked at.  The sizing call of a large file does not read the
    payload, so it must not report a body error either (it would take precedence over "Data CRC16 failed"): it returns OK
    with no images, and the transcoding call delivers the verdict.  Small files report everything from the sizing call."""
    import basisu_rs_b200 as b
    from basis_writer import uastc_file
    L = b.lib()
    cnt, need = ctypes.c_uint32(7), ctypes.c_uint64(7)
    big = np.zeros((160 * 128, 16), dtype=np.uint8).tobytes()
    g = _patch_header(uastc_file(big, 160, 128, corrupt="data"), slice_desc_file_ofs=len(big) + 4096)     # bad CRC AND bad slice table
    assert L.b2bu_read_to(b.BC7, g, len(g), None, None, 0, ctypes.byref(cnt), None, 0, ctypes.byref(need)) == 0
    assert (cnt.value, need.value) == (0, 0)
    small = np.zeros((6, 16), dtype=np.uint8).tobytes()
    g = _patch_header(uastc_file(small, 3, 2), slice_desc_file_ofs=4096)
    # the data CRC covers everything after the header, so patching the header keeps it valid: the body error shows
    assert L.b2bu_read_to(b.BC7, g, len(g), None, None, 0, ctypes.byref(cnt), None, 0, ctypes.byref(need)) == 8


def test_crafted_slice_count_does_not_allocate_or_unwind():
    """total_slices is a 24-bit field of an unverified header: a 100-byte file claiming 16.7 M slices must fail with the
    reference's slice-descriptor error, not with a 670 MB allocation or an exception through the C ABI."""
    import resource
    import basisu_rs_b200 as b
    from basis_writer import uastc_file
    L = b.lib()
    cnt, need = ctypes.c_uint32(0), ctypes.c_uint64(0)
    g = _patch_header(uastc_file(np.zeros((1, 16), dtype=np.uint8).tobytes(), 1, 1), total_slices=0xFFFFFF)
    before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    st = L.b2bu_read_to(b.BC7, g, len(g), None, None, 0, ctypes.byref(cnt), None, 0, ctypes.byref(need))
    assert st in (8, 16)                                               # B2BU_ERR_RANGE / B2BU_ERR_SLICE_DESC
    assert resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - before < 64 * 1024      # KiB
    assert L.b2bu_error_string(19) == b"out of host memory"


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    import basisu_rs_b200 as b
    monkeypatch.setattr(b, "_LIB", None)
    monkeypatch.setattr(b, "_HERE", tmp_path)
    with pytest.raises(ImportError):
        b.lib()
