"""CPU tests of the oracle (oracle/basisu_oracle.c): it must reproduce every golden vector the
reference's own tests hold for this path before it is trusted as the checker.
Mirrors reference tests/transcode_uastc_block.rs:35-78, src/bitreader.rs:63-100,
src/bitwriter.rs:118-225, src/basis.rs:574-621."""
import ctypes

import numpy as np
import pytest

from conftest import OUT_BYTES, TARGETS, oracle_transcode, rgba_image_to_blocks
from basis_writer import crc16, _crc16_fast


@pytest.mark.parametrize("name", list(TARGETS))
def test_oracle_reproduces_reference_kat_vectors(oracle, kat, name):
    t = TARGETS[name]
    fn = {0: "orc_unpack_uastc_block_to_rgba", 1: "orc_transcode_uastc_block_to_astc", 2: "orc_transcode_uastc_block_to_bc7",
          3: "orc_transcode_uastc_block_to_etc1", 4: "orc_transcode_uastc_block_to_etc2"}[t]
    f = getattr(oracle, fn)
    for i in range(kat.n):
        out = ctypes.create_string_buffer(OUT_BYTES[t])
        assert f(kat.inputs[i].tobytes(), out) == 0
        assert out.raw == kat.expected[t][i].tobytes(), f"mode {kat.modes[i]} block {i}"


def test_oracle_slice_api_matches_block_api(oracle, kat):
    for t in range(5):
        e, bad, out = oracle_transcode(oracle, t, kat.inputs, blocks_per_row=19, threads=3)
        assert e == 0
        got = rgba_image_to_blocks(out, kat.n, 19) if t == 0 else out.reshape(kat.n, OUT_BYTES[t])
        assert (got == kat.expected[t]).all()


def test_oracle_error_paths(oracle):
    # uastc.rs:55-56 / :336 / :364
    blk = np.zeros((3, 16), dtype=np.uint8)
    blk[:, 0] = 0x01                       # mode 0
    blk[1, 0] = 69                         # invalid mode code
    e, bad, _ = oracle_transcode(oracle, 1, blk)
    assert (e, bad) == (2, 1)
    blk[1, 0] = 0x1D                       # mode 2, partition field = bits 20..24
    blk[1, 2] = 0xF0; blk[1, 3] = 0x01     # partition 31 >= 30
    e, bad, _ = oracle_transcode(oracle, 2, blk)
    assert (e, bad) == (3, 1)
    out = np.zeros(64, dtype=np.uint8)
    assert oracle.orc_uastc_transcode_slice(1, blk.ctypes.data, 17, 1, out.ctypes.data, 1, None) == 1


def test_bitreader_sweep(oracle):
    """src/bitreader.rs:63-100: 16 patterns x all (offset, len) in 0..32."""
    pattern = 0x5555_5555_5555_5555
    for i in range(16):
        seg = 0xFFFF
        xor = (seg * ((i >> 3) & 1)) << 48 | (seg * ((i >> 2) & 1)) << 32 | (seg * ((i >> 1) & 1)) << 16 | (seg * (i & 1))
        data = pattern ^ xor
        raw = data.to_bytes(8, "little")
        for ln in range(32):
            for off in range(32):
                assert oracle.orc_bitreader_read_at(raw, 8, 0, off) == data & ((1 << off) - 1)
                assert oracle.orc_bitreader_read_at(raw, 8, off, ln) == (data >> off) & ((1 << ln) - 1)
    # reads past the end yield zeros (bitreader.rs:44,55)
    assert oracle.orc_bitreader_read_at(b"\xff", 1, 4, 32) == 0xF


def test_bitwriter_sweeps(oracle):
    """src/bitwriter.rs:118-225: LSB writer and MSB-from-the-end writer with and without bit reversal."""
    rng = np.random.default_rng(7)
    for _ in range(400):
        off, ln = int(rng.integers(0, 32)), int(rng.integers(0, 33))
        v = int(rng.integers(0, 1 << 32))
        buf = ctypes.create_string_buffer(8)
        oracle.orc_bitwriter_lsb(buf, 8, off, ln, v)
        want = (v & ((1 << ln) - 1)) << off
        assert int.from_bytes(buf.raw, "little") == want & ((1 << 64) - 1)
        buf = ctypes.create_string_buffer(8)
        oracle.orc_bitwriter_msb(buf, 8, off, ln, v, 0)
        want = (v & ((1 << ln) - 1)) << (64 - off - ln)
        assert int.from_bytes(buf.raw, "little") == want
        buf = ctypes.create_string_buffer(8)
        oracle.orc_bitwriter_msb(buf, 8, off, ln, v, 1)
        rev = int(format(v & 0xFFFFFFFF, "032b")[::-1], 2) >> ((32 - ln) & 31)
        want = (rev & ((1 << ln) - 1)) << (64 - off - ln)
        assert int.from_bytes(buf.raw, "little") == want


def test_header_field_layout(oracle):
    """src/basis.rs:574-621: header bytes 0..77 map to the 26 fields in order."""
    raw = bytes(range(77))
    fields = (ctypes.c_uint32 * 26)()
    oracle.orc_parse_header.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    assert oracle.orc_parse_header(raw, 77, fields) == 0
    widths = [2, 2, 2, 2, 4, 2, 3, 3, 1, 2, 1, 3, 4, 4, 4, 2, 4, 3, 2, 4, 3, 4, 4, 4, 4, 4]
    pos = 0
    for i, w in enumerate(widths):
        assert fields[i] == int.from_bytes(raw[pos:pos + w], "little")
        pos += w
    assert pos == 77


def test_crc16_matches_independent_implementations(oracle):
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 69, 1000):
        d = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert oracle.orc_crc16(d, n, 0) == crc16(d) == _crc16_fast(d)
    assert oracle.orc_crc16(b"123456789", 9, 0) == 0xD64E       # CRC-16/GENIBUS check value


def test_endpoint_unquant_matches_astc_spec_tables(oracle):
    """ASTC colour unquantisation: pure-bit ranges replicate bits; the trit/quint ranges are monotone
    in the quantisation order and span 0..255 (uastc.rs:585-614)."""
    for r, bits in ((8, 4), (11, 5), (20, 8)):
        for m in range(1 << bits):
            v = oracle.orc_unquant_endpoint(0, m, r)
            rep = m << (8 - bits)
            x = 0
            while rep:
                x |= rep
                rep >>= bits
            assert v == x & 0xFF
    for r, bits, nd in ((7, 2, 3), (12, 3, 5), (13, 4, 3), (18, 5, 5), (19, 6, 3)):
        vals = sorted(oracle.orc_unquant_endpoint(d, m, r) for d in range(nd) for m in range(1 << bits))
        assert vals[0] == 0 and vals[-1] == 255 and len(set(vals)) == nd << bits
