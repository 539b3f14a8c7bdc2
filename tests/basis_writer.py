"""Minimal .basis container WRITER for tests and bench inputs (the reference only reads).

Layout follows reference src/basis.rs:419-572 (Header 77 B, SliceDesc 23 B, CRC-16 over bytes
8..77 and 77..end).  The CRC here is an independent bit-serial CRC-16/GENIBUS-style
implementation (poly 0x1021, init 0xFFFF, xorout 0xFFFF, MSB first), not the reference's nibble form."""
import struct


def crc16(data: bytes, crc: int = 0) -> int:
    crc = (~crc) & 0xFFFF
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return (~crc) & 0xFFFF


def _crc16_fast(data: bytes, crc: int = 0) -> int:
    tab = _crc16_fast.tab
    crc = (~crc) & 0xFFFF
    for b in data:
        crc = ((crc << 8) & 0xFFFF) ^ tab[(crc >> 8) ^ b]
    return (~crc) & 0xFFFF


def _make_tab():
    tab = []
    for i in range(256):
        c = i << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x1021) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
        tab.append(c)
    return tab


_crc16_fast.tab = _make_tab()


def u24(v):
    return struct.pack("<I", v)[:3]


def build_basis(slices, tex_format=1, flags=0, tex_type=0, total_images=None, etc1s=None, corrupt=None):
    """slices: list of dicts {data, orig_width, orig_height, num_blocks_x, num_blocks_y, flags?, image_index?, level_index?}.
    etc1s: dict {endpoints, selectors, tables, total_endpoints, total_selectors} for ETC1S files.
    Returns the file bytes."""
    n = len(slices)
    header_size = 77
    desc_ofs = header_size
    pos = desc_ofs + 23 * n
    sections = b""
    ep_ofs = ep_size = sel_ofs = sel_size = tab_ofs = tab_size = 0
    total_endpoints = total_selectors = 0
    if etc1s is not None:
        ep_ofs, ep_size = pos, len(etc1s["endpoints"]); pos += ep_size
        sel_ofs, sel_size = pos, len(etc1s["selectors"]); pos += sel_size
        tab_ofs, tab_size = pos, len(etc1s["tables"]); pos += tab_size
        sections = etc1s["endpoints"] + etc1s["selectors"] + etc1s["tables"]
        total_endpoints, total_selectors = etc1s["total_endpoints"], etc1s["total_selectors"]
    descs = b""
    payload = b""
    for i, s in enumerate(slices):
        d = bytes(s["data"])
        descs += u24(s.get("image_index", i)) + struct.pack("<BB", s.get("level_index", 0), s.get("flags", 0))
        descs += struct.pack("<HHHH", s["orig_width"], s["orig_height"], s["num_blocks_x"], s["num_blocks_y"])
        descs += struct.pack("<II", pos, len(d)) + struct.pack("<H", _crc16_fast(d))
        payload += d
        pos += len(d)
    body = descs + sections + payload
    rest = struct.pack("<I", len(body)) + struct.pack("<H", _crc16_fast(body))
    rest += u24(n) + u24(total_images if total_images is not None else n)
    rest += struct.pack("<B", tex_format) + struct.pack("<H", flags) + struct.pack("<B", tex_type) + u24(0)
    rest += struct.pack("<III", 0, 0, 0)
    rest += struct.pack("<H", total_endpoints) + struct.pack("<I", ep_ofs) + u24(ep_size)
    rest += struct.pack("<H", total_selectors) + struct.pack("<I", sel_ofs) + u24(sel_size)
    rest += struct.pack("<II", tab_ofs, tab_size)
    rest += struct.pack("<I", desc_ofs) + struct.pack("<II", 0, 0)
    assert len(rest) == 77 - 8
    head = struct.pack("<HHH", 0x4273, 0x13, header_size) + struct.pack("<H", _crc16_fast(rest))
    out = bytearray(head + rest + body)
    if corrupt == "data":
        out[-1] ^= 0x5A
    elif corrupt == "header":
        out[20] ^= 0x01
    elif corrupt == "sig":
        out[0] ^= 0xFF
    return bytes(out)


def uastc_file(blocks_bytes: bytes, num_blocks_x: int, num_blocks_y: int, **kw):
    return build_basis([dict(data=blocks_bytes, orig_width=4 * num_blocks_x, orig_height=4 * num_blocks_y,
                             num_blocks_x=num_blocks_x, num_blocks_y=num_blocks_y)], tex_format=1, **kw)
