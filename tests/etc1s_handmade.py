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
    # two texels worked out by hand as literals: block (2,0) has endpoint 0 = intensity 3, (17,15,17) -> (140,123,140) and selector 1
    # (texel x has value x): x = 0 -> -42 -> (98, 81, 98); x = 3 -> +42 -> (182, 165, 182)
    assert tuple(img[0, 8]) == (98, 81, 98, 255) and tuple(img[0, 11]) == (182, 165, 182, 255)
    # block (3,1): endpoint 1 = intensity 3, (18,16,16) -> (148,132,132), selector 1: x = 1 -> -13 -> (135, 119, 119)
    assert tuple(img[4, 13]) == (135, 119, 119, 255)
    return img
