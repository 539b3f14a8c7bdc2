"""CPU tests of the ETC1S half of the oracle.  PARITY UNPINNED: the reference holds no ETC1S vector
(SURVEY.md section 8c), so these tests pin internal consistency only: the test encoder and the oracle
decoder (a restatement of src/basis_lz/mod.rs) must round-trip every construct of the grammar."""
import numpy as np
import pytest

from etc1s_common import bind, decode_bc1, etc1s_file, make_case, oracle_bc1, oracle_etc1, oracle_open, oracle_read_to, oracle_rgba, slice_bytes

CASES = [  # nbx, nby, slices, codebook size, history, raw selectors, video
    (16, 16, 2, 64, 64, False, False), (33, 17, 3, 300, 64, True, False), (1, 1, 1, 4, 0, False, False),
    (64, 64, 2, 4096, 64, False, False), (7, 5, 2, 50, 16, False, True), (128, 96, 1, 1000, 200, False, False),
]


@pytest.mark.parametrize("nbx,nby,ns,ncb,hist,raw,video", CASES)
def test_encoder_oracle_round_trip(oracle, nbx, nby, ns, ncb, hist, raw, video):
    orc = bind(oracle)
    ep_cb, sel_cb, ei, si, enc = make_case(orc, nbx, nby, ns, ncb, hist, raw, video, seed=nbx)
    e, h = oracle_open(orc, enc, ncb, ncb, video)
    assert e == 0
    cb_e = np.zeros((ncb, 4), np.uint8)
    cb_s = np.zeros((ncb, 8), np.uint8)
    orc.orc_etc1s_codebooks(h, cb_e.ctypes.data, cb_s.ctypes.data)
    assert (cb_e == ep_cb).all() and (cb_s[:, :4] == sel_cb).all()
    for k in range(ns):
        a = np.zeros(nbx * nby, np.uint16)
        b = np.zeros(nbx * nby, np.uint16)
        d = slice_bytes(enc, k)
        assert orc.orc_etc1s_decode_indices(h, nbx, nby, d, len(d), a.ctypes.data, b.ctypes.data) == 0
        assert (a == ei[k]).all() and (b == si[k]).all()
    orc.orc_etc1s_close(h)


def test_etc1_block_layout_and_rgba_colours(oracle):
    """mod.rs:163-181 / :122-146 on a hand-checkable 1-block slice."""
    orc = bind(oracle)
    ep_cb = np.array([[3, 10, 20, 30], [5, 1, 2, 3]], dtype=np.uint8)
    sel_cb = np.array([[0b11100100, 0, 0xFF, 0x1B], [0, 0, 0, 0]], dtype=np.uint8)
    ei = np.array([[0]], dtype=np.uint16)
    si = np.array([[0]], dtype=np.uint16)
    from etc1s_synth import encode
    enc = encode(orc, ep_cb, sel_cb, ei, si, 1, 1, 64)
    e, h = oracle_open(orc, enc, 2, 2)
    assert e == 0
    e, etc1 = oracle_etc1(orc, h, 1, 1, slice_bytes(enc, 0))
    assert e == 0
    assert etc1[:4] == bytes([10 << 3, 20 << 3, 30 << 3, (3 << 5) | (3 << 2) | 3])
    e, rgba = oracle_rgba(orc, h, 1, 1, slice_bytes(enc, 0))
    px = np.frombuffer(rgba, dtype=np.uint8).reshape(4, 4, 4)
    base = [(10 << 3) | (10 >> 2), (20 << 3) | (20 >> 2), (30 << 3) | (30 >> 2)]
    mods = [-42, -13, 13, 42]                      # etc.rs:440 intensity table 3
    for x in range(4):                             # row 0 of selector 0 = 0b11100100 -> selectors 0,1,2,3
        want = [min(255, max(0, b + mods[x])) for b in base] + [255]
        assert px[0, x].tolist() == want
    orc.orc_etc1s_close(h)


def test_file_level_etc1s_with_alpha_pairs(oracle):
    orc = bind(oracle)
    nbx, nby, ncb = 9, 6, 128
    _, _, ei, si, enc = make_case(orc, nbx, nby, 4, ncb, seed=5)
    f = etc1s_file(enc, nbx, nby, ncb, alpha_pairs=True, orig=(35, 22))
    e, imgs = oracle_read_to(orc, 0, f)
    assert e == 0 and len(imgs) == 2
    w, h, stride, data = imgs[0]
    assert (w, h, stride) == (35, 22, 16 * 35)            # quirk C-6: stride from orig_width (basis.rs:43-49)
    assert len(data) == nbx * nby * 64
    e, imgs1 = oracle_read_to(orc, 3, f)                    # read_to_etc1 returns every slice, alpha slices included (basis.rs:109-123)
    assert e == 0 and len(imgs1) == 4 and imgs1[0][2] == 8 * nbx
    assert oracle_read_to(orc, 2, f)[0] == 11              # ETC1S -> BC7: reference unimplemented!()


def test_bc1_extension_definition_on_a_hand_checkable_block(oracle):
    """ETC1S -> BC1 is an EXTENSION (absent from the reference); this pins its definition: endpoints = RGB565 of the lowest
    and highest ETC1S colour in use, selectors remapped to the nearest BC1 palette entry, solid blocks use c0 == c1."""
    orc = bind(oracle)
    ep_cb = np.array([[3, 10, 20, 30], [0, 16, 16, 16]], dtype=np.uint8)           # (inten, r5, g5, b5)
    sel_cb = np.array([[0b11100100, 0b11100100, 0b00011011, 0b11111111], [0b01010101] * 4], dtype=np.uint8)
    ei = np.array([[0, 1]], dtype=np.uint16)
    si = np.array([[0, 1]], dtype=np.uint16)
    from etc1s_synth import encode
    enc = encode(orc, ep_cb, sel_cb, ei, si, 2, 1, 64)
    e, h = oracle_open(orc, enc, 2, 2)
    assert e == 0
    e, bc1 = oracle_bc1(orc, h, 2, 1, slice_bytes(enc, 0))
    assert e == 0
    b0, b1 = bc1[:8], bc1[8:]
    base = [(10 << 3) | (10 >> 2), (20 << 3) | (20 >> 2), (30 << 3) | (30 >> 2)]
    lo = [min(255, max(0, v - 42)) for v in base]
    hi = [min(255, max(0, v + 42)) for v in base]
    q = lambda c: ((c[0] * 31 + 127) // 255) << 11 | ((c[1] * 63 + 127) // 255) << 5 | ((c[2] * 31 + 127) // 255)
    c0, c1 = max(q(lo), q(hi)), min(q(lo), q(hi))
    assert b0[0] | b0[1] << 8 == c0 and b0[2] | b0[3] << 8 == c1 and c0 > c1
    # selector 3 (brightest) -> palette 0 (c0 = the brighter endpoint here), selector 0 -> palette 1, 2 -> 2, 1 -> 3
    assert b0[4] == 0b00101101 and b0[5] == 0b00101101 and b0[6] == 0b01111000 and b0[7] == 0
    # block 1 uses one selector only: solid, c0 == c1, all indices 0
    assert b1[0:2] == b1[2:4] and b1[4:] == bytes(4)
    orc.orc_etc1s_close(h)


def test_bc1_extension_stays_close_to_the_etc1s_image(oracle):
    orc = bind(oracle)
    nbx, nby, ncb = 40, 24, 512
    _, _, ei, si, enc = make_case(orc, nbx, nby, 1, ncb, seed=9)
    e, h = oracle_open(orc, enc, ncb, ncb)
    d = slice_bytes(enc, 0)
    e, bc1 = oracle_bc1(orc, h, nbx, nby, d)
    assert e == 0
    _, rgba = oracle_rgba(orc, h, nbx, nby, d)
    ref = np.frombuffer(rgba, dtype=np.uint8).reshape(nby * 4, nbx * 4, 4)[..., :3].astype(np.int32)
    err = np.abs(decode_bc1(bc1, nbx, nby).astype(np.int32) - ref)
    assert err.mean() < 9.0 and err.max() <= 96      # random codebooks: many high-intensity tables whose clamped colours leave the BC1 line
    orc.orc_etc1s_close(h)
