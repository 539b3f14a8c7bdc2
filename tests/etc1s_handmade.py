"""A tiny ETC1S / BasisLZ file body assembled BY HAND, bit by bit, from the published Basis Universal file-format description
(compressed Huffman tables, endpoint / selector codebooks, slice decoding with endpoint prediction and the selector history).
It does not come from oracle/etc1s_encoder.inc or tests/etc1s_synth.py: every field below is written out with its value and
the reason for it, and the expected result (endpoint / selector index per block, texels) is derived in the comments and stated
as literals.  The oracle and the CUDA path must both decode the stream to exactly this.

Bit order: everything is LSB-first; a Huffman code is written most-significant CODE bit first (the decoder's flat table is
indexed by the bit-reversed code), i.e. a canonical code 'ab' appears in the stream as bit a, then bit b.
"""
import numpy as np


class Bits:
    """append-only LSB-first bit string"""

    def __init__(self):
        self.bits = []

    def put(self, value, nbits):                       # plain field: least significant bit first
        for i in range(nbits):
            self.bits.append((value >> i) & 1)
        return self

    def code(self, canonical):                         # Huffman code given as a string of bits, most significant first
        for ch in canonical:
            self.bits.append(int(ch))
        return self

    def bytes(self):
        b = self.bits + [0] * (-len(self.bits) % 8)
        return bytes(sum(b[i + k] << k for k in range(8)) for i in range(0, len(b), 8))


# order in which the code-length code sizes are stored
CL_ORDER = [17, 18, 19, 20, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 16]


def huff_table(w, total_syms, size_of):
    """One compressed Huffman table whose used symbols all have the SAME code size L (1 or 2): size_of maps symbol -> L.
    Header: total_used_syms:14, num_codelength_codes:5, then 3 bits per code-length symbol in CL_ORDER.  The code-length
    alphabet here has exactly two used symbols, 0 ("symbol unused") and L, each with a 1-bit code: canonical assignment gives
    the smaller symbol (0) the code '0' and L the code '1'.  The sizes of symbols 0 .. total_syms-1 follow, one bit each."""
    (L,) = set(size_of.values())
    assert L in (1, 2)
    w.put(total_syms, 14)
    n_cl = CL_ORDER.index(L) + 1                       # entries up to and including the one for code-length symbol L
    w.put(n_cl, 5)
    for i in range(n_cl):
        w.put(1 if CL_ORDER[i] in (0, L) else 0, 3)    # 1-bit codes for code-length symbols 0 and L, everything else unused
    for s in range(total_syms):
        w.code("1" if s in size_of else "0")
    # canonical codes of the table itself: symbols in ascending order take consecutive codes of L bits, starting at 0
    used = sorted(size_of)
    return {s: format(i, "0%db" % L) for i, s in enumerate(used)}


def build():
    """returns dict(endpoints, selectors, tables, slice, nbx, nby, n, expect_ep, expect_sel, ep_cb, sel_cb)"""
    n = 3                                              # endpoints == selectors (the reference passes total_selectors for both)

    # ---------------- endpoint codebook: 3 colour-delta tables (chosen by the previous value's band), intensity-delta table,
    # grayscale flag, then per entry: intensity delta, R, G, B deltas (5-bit colours start at 16, intensity at 0)
    w = Bits()
    col = [huff_table(w, 32, {1: 1, 31: 1}) for _ in range(3)]        # '0' -> +1, '1' -> +31 (= -1 mod 32) in all three bands
    inten = huff_table(w, 8, {0: 1, 3: 1})                            # '0' -> +0, '1' -> +3 (mod 8)
    w.put(0, 1)                                                       # not grayscale
    # every colour stays inside 10..21, so band 1 (col[1]) is the table in use throughout
    w.code(inten[3]).code(col[1][1]).code(col[1][31]).code(col[1][1])     # entry 0: inten 0+3 = 3, rgb (16+1, 16-1, 16+1) = (17, 15, 17)
    w.code(inten[0]).code(col[1][1]).code(col[1][1]).code(col[1][31])     # entry 1: inten 3,       rgb (18, 16, 16)
    w.code(inten[3]).code(col[1][31]).code(col[1][31]).code(col[1][31])   # entry 2: inten 6,       rgb (17, 15, 15)
    endpoints = w.bytes()
    ep_cb = [(3, 17, 15, 17), (3, 18, 16, 16), (6, 17, 15, 15)]       # (intensity table, r5, g5, b5)

    # ---------------- selector codebook: global = 0, hybrid = 0, raw = 1, then 4 row bytes per selector (texel x of a row at bits 2x..2x+1)
    w = Bits()
    w.put(0, 1).put(0, 1).put(1, 1)
    sel_cb = [[0x00, 0x55, 0xAA, 0xFF],                # selector 0: every texel of row y has the value y
              [0xE4, 0xE4, 0xE4, 0xE4],                # selector 1: texel x has the value x (0b11_10_01_00)
              [0x1B, 0xE4, 0x1B, 0xE4]]                # selector 2: rows alternate 3,2,1,0 / 0,1,2,3 (in the codebook, unused by the slice)
    for rows in sel_cb:
        for r in rows:
            w.put(r, 8)
    selectors = w.bytes()

    # ---------------- slice models: endpoint predictors (one symbol per 2x2 group: 2 bits per block, (x,y) (x+1,y) (x,y+1)
    # (x+1,y+1) from the low bits up; 0 = left, 1 = up, 2 = up-left, 3 = delta), endpoint delta, selector, history run length,
    # then history_size:13
    hist = 2
    w = Bits()
    pred = huff_table(w, 148, {79: 1, 147: 1})         # '0' -> 79 = 3|3<<2|0<<4|1<<6, '1' -> 147 = 3|0<<2|1<<4|2<<6
    delta = huff_table(w, 3, {1: 1, 2: 1})             # '0' -> +1, '1' -> +2 (mod 3 endpoints)
    sel = huff_table(w, n + hist, {0: 2, 1: 2, 3: 2, 4: 2})   # '00' selector 0, '01' selector 1, '10' history[0], '11' history[1]
    huff_table(w, 2, {0: 1, 1: 1})                     # run-length model: present, never used (no symbol n + hist = 5 in the slice)
    w.put(hist, 13)
    tables = w.bytes()

    # ---------------- the slice: 4 x 2 blocks, raster order.  Per block: [predictor symbol at even x of even rows], [endpoint delta
    # if its predictor is 3], selector symbol.  prev endpoint = 0 at the start; history = [0, 0], insert position = size/2 = 1
    w = Bits()
    # (0,0): group symbol 147 -> predictors (0,0)=3 (1,0)=0 (0,1)=1 (1,1)=2
    w.code(pred[147]).code(delta[2]).code(sel[1])      # delta: 0 + 2 = 2               | selector 1 -> history [0, 1]
    # (1,0): left = 2
    w.code(sel[4])                                     # history[1] = 1, swap with [0]  -> history [1, 0]           | selector 1
    # (2,0): group symbol 79 -> predictors (2,0)=3 (3,0)=3 (2,1)=0 (3,1)=1
    w.code(pred[79]).code(delta[1]).code(sel[3])       # delta: 2 + 1 = 3 -> wraps to 0 | history[0] = 1 (no swap)  | selector 1
    # (3,0): delta again
    w.code(delta[1]).code(sel[0])                      # 0 + 1 = 1                      | selector 0 -> history [1, 0]
    # (0,1): up = endpoint of (0,0) = 2
    w.code(sel[4])                                     # history[1] = 0, swap           -> history [0, 1]           | selector 0
    # (1,1): up-left = endpoint of (0,0) = 2
    w.code(sel[1])                                     # selector 1 -> history [0, 1]
    # (2,1): left = previous block's endpoint = 2
    w.code(sel[3])                                     # history[0] = 0                                              | selector 0
    # (3,1): up = endpoint of (3,0) = 1
    w.code(sel[1])                                     # selector 1
    expect_ep = np.array([[2, 2, 0, 1], [2, 2, 2, 1]], dtype=np.uint16)
    expect_sel = np.array([[1, 1, 1, 0], [0, 1, 0, 1]], dtype=np.uint16)
    return dict(endpoints=endpoints, selectors=selectors, tables=tables, slice=w.bytes(), nbx=4, nby=2, n=n,
                expect_ep=expect_ep, expect_sel=expect_sel, ep_cb=ep_cb, sel_cb=sel_cb)


def build_runs():
    """A second hand-made body for the paths the first one does not touch: a GRAYSCALE endpoint codebook, a HUFFMAN-coded
    (XOR-DPCM) selector codebook, the predictor repeat symbol (256 + 4-bit-chunk VLC), the selector-history run symbol with a
    plain count and with the 63 escape + 7-bit-chunk VLC, and a history hit that swaps two different entries."""
    n, hist = 4, 4
    # ---------------- endpoint codebook, grayscale: only R is coded, G and B copy it
    w = Bits()
    col = [huff_table(w, 32, {1: 1, 31: 1}) for _ in range(3)]        # '0' -> +1, '1' -> -1
    inten = huff_table(w, 8, {0: 1, 3: 1})                            # '0' -> +0, '1' -> +3
    w.put(1, 1)                                                       # grayscale
    w.code(inten[3]).code(col[1][1])                                  # entry 0: inten 3, grey 17
    w.code(inten[3]).code(col[1][1])                                  # entry 1: inten 6, grey 18
    w.code(inten[0]).code(col[1][31])                                 # entry 2: inten 6, grey 17
    w.code(inten[3]).code(col[1][31])                                 # entry 3: inten (6 + 3) & 7 = 1, grey 16
    endpoints = w.bytes()
    ep_cb = [(3, 17, 17, 17), (6, 18, 18, 18), (6, 17, 17, 17), (1, 16, 16, 16)]

    # ---------------- selector codebook, Huffman coded: global = 0, hybrid = 0, raw = 0, the delta model, selector 0 as four raw
    # bytes, then four symbols per selector, each XOR-ed onto the same row of the previous selector
    w = Bits()
    w.put(0, 1).put(0, 1).put(0, 1)
    dsel = huff_table(w, 256, {0x00: 1, 0xFF: 1})                     # '0' -> row unchanged, '1' -> row inverted
    for r in (0x1B, 0x1B, 0xE4, 0xE4):
        w.put(r, 8)
    w.code(dsel[0x00]).code(dsel[0xFF]).code(dsel[0x00]).code(dsel[0xFF])   # selector 1: 1B E4 E4 1B
    w.code(dsel[0xFF]).code(dsel[0xFF]).code(dsel[0xFF]).code(dsel[0xFF])   # selector 2: E4 1B 1B E4
    w.code(dsel[0x00]).code(dsel[0x00]).code(dsel[0x00]).code(dsel[0x00])   # selector 3: the same rows
    selectors = w.bytes()
    sel_cb = [[0x1B, 0x1B, 0xE4, 0xE4], [0x1B, 0xE4, 0xE4, 0x1B], [0xE4, 0x1B, 0x1B, 0xE4], [0xE4, 0x1B, 0x1B, 0xE4]]

    # ---------------- slice models
    w = Bits()
    pred = huff_table(w, 257, {83: 1, 256: 1})         # '0' -> 83 = 3|0<<2|1<<4|1<<6 (delta, left / up, up), '1' -> 256 = repeat
    delta = huff_table(w, 4, {0: 2, 1: 2, 2: 2, 3: 2})                # '00' +0, '01' +1, '10' +2, '11' +3 (mod 4 endpoints)
    sel = huff_table(w, n + hist + 1, {1: 2, 2: 2, 7: 2, 8: 2})       # '00' selector 1, '01' selector 2, '10' history[3], '11' = n + hist: run
    rle = huff_table(w, 64, {0: 1, 63: 1})                            # '0' -> count 0, '1' -> 63 = escape, a 7-bit-chunk VLC follows
    w.put(hist, 13)
    tables = w.bytes()

    # ---------------- the slice: 8 x 2 blocks.  history = [0,0,0,0], insert position 2 (wraps back to 2 after 3)
    w = Bits()
    # (0,0): predictor symbol 83 for the 2x2 group; delta +1 -> endpoint 1; selector 1 -> history [0,0,1,0]
    w.code(pred[83]).code(delta[1]).code(sel[1])
    # (1,0): left -> 1; selector 2 -> history [0,0,1,2]
    w.code(sel[2])
    # (2,0): predictor symbol 256 = repeat the previous symbol; VLC(4) value 0 (one 5-bit chunk, no continuation) -> this group and
    #        2 more use 83.  delta +3: 1 + 3 = 4 -> wraps to 0; history[3] = 2, swapped with history[1] -> [0,2,1,0]
    w.code(pred[256]).put(0, 5).code(delta[3]).code(sel[7])
    # (3,0): left -> 0; run symbol, count symbol 0 -> this block and 3 + 0 - 1 = 2 more read history[0] = 0
    w.code(sel[8]).code(rle[0])
    # (4,0): (repeat 2 -> 1) delta +2 -> 2; selector from the run
    w.code(delta[2])
    # (5,0): left -> 2; run ends here
    # (6,0): (repeat 1 -> 0) delta +0 -> 2; selector 1 -> history [0,2,1,0] (insert position 2 -> 3)
    w.code(delta[0]).code(sel[1])
    # (7,0): left -> 2; run symbol with the escape: count symbol 63, then VLC(7) value 1 (one 8-bit chunk) -> 3 + 1 - 1 = 3 more blocks
    w.code(sel[8]).code(rle[63]).put(1, 8)
    # row 1: every predictor is "up" (upper nibble of 83), no predictor or delta symbols
    # (0,1) (1,1) (2,1): the run -> history[0] = 0
    # (3,1): selector 2 -> inserted at position 3: history [0,2,1,2], insert position wraps to 2
    w.code(sel[2])
    # (4,1): history[3] = 2, swapped with history[1] = 2
    w.code(sel[7])
    # (5,1): selector 1   (6,1): selector 2   (7,1): selector 1
    w.code(sel[1]).code(sel[2]).code(sel[1])
    expect_ep = np.array([[1, 1, 0, 0, 2, 2, 2, 2], [1, 1, 0, 0, 2, 2, 2, 2]], dtype=np.uint16)
    expect_sel = np.array([[1, 2, 2, 0, 0, 0, 1, 0], [0, 0, 0, 2, 2, 1, 2, 1]], dtype=np.uint16)
    return dict(endpoints=endpoints, selectors=selectors, tables=tables, slice=w.bytes(), nbx=8, nby=2, n=n,
                expect_ep=expect_ep, expect_sel=expect_sel, ep_cb=ep_cb, sel_cb=sel_cb)


# ETC1 intensity modifier table (ETC1 specification), ordered as the ETC1S selector value indexes it: -large, -small, +small, +large
ETC1_MOD = [(8, 2), (17, 5), (29, 9), (42, 13), (60, 18), (80, 24), (106, 33), (183, 47)]


def expected_rgba(case):
    """RGBA image (4*nby, 4*nbx, 4) from the hand-derived indices and the hand-written codebooks: colour = 5-bit base expanded to 8
    bits ((c << 3) | (c >> 2)) plus the modifier of the block's intensity table selected by the texel's 2-bit value, clamped."""
    nbx, nby = case["nbx"], case["nby"]
    img = np.zeros((4 * nby, 4 * nbx, 4), dtype=np.uint8)
    for by in range(nby):
        for bx in range(nbx):
            inten, r5, g5, b5 = case["ep_cb"][case["expect_ep"][by, bx]]
            rows = case["sel_cb"][case["expect_sel"][by, bx]]
            big, small = ETC1_MOD[inten]
            mod = [-big, -small, small, big]
            for y in range(4):
                for x in range(4):
                    m = mod[(rows[y] >> (2 * x)) & 3]
                    px = [min(255, max(0, ((c << 3) | (c >> 2)) + m)) for c in (r5, g5, b5)]
                    img[4 * by + y, 4 * bx + x] = px + [255]
    if case["nbx"] == 4:
        # two texels worked out by hand as literals: block (2,0) has endpoint 0 = intensity 3, (17,15,17) -> (140,123,140) and
        # selector 1 (texel x has value x): x = 0 -> -42 -> (98, 81, 98); x = 3 -> +42 -> (182, 165, 182)
        assert tuple(img[0, 8]) == (98, 81, 98, 255) and tuple(img[0, 11]) == (182, 165, 182, 255)
        # block (3,1): endpoint 1 = intensity 3, (18,16,16) -> (148,132,132), selector 1: x = 1 -> -13 -> (135, 119, 119)
        assert tuple(img[4, 13]) == (135, 119, 119, 255)
    else:
        # build_runs(): block (0,0) has endpoint 1 = intensity 6, grey 18 -> 148, selector 1 row 0 = 0x1B (texel x has value 3 - x):
        # x = 0 -> +106 -> 254; x = 3 -> -106 -> 42
        assert tuple(img[0, 0]) == (254, 254, 254, 255) and tuple(img[0, 3]) == (42, 42, 42, 255)
    return img
