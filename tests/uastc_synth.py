"""Valid-random UASTC block synthesiser (test + bench input generator).

Every 128-bit pattern is accepted by the reference's UASTC paths except mode code 69 and
out-of-range partition indices, so a valid block is: 16 random
bytes, then the mode code patched in, then the partition field patched to a legal index.
Covers all 19 modes uniformly, all 60 partitions, all component selectors, void-extent colours
with 0 / 255 channels (BC7 mode-5 path) and every transcoder-hint value.
"""
import numpy as np

# mode id -> (code value, code bits)   (reference src/uastc.rs:560-577 MODE_LUT, Appendix A of SURVEY.md)
MODE_CODES = {0: (0b0001, 4), 1: (0b110101, 6), 2: (0b11101, 5), 3: (0b00011, 5), 4: (0b10011, 5), 5: (0b01011, 5),
              6: (0b11011, 5), 7: (0b00111, 5), 8: (0b10111, 5), 9: (0b01111, 5), 10: (0b010, 3), 11: (0b00, 2),
              12: (0b110, 3), 13: (0b11111, 5), 14: (0b01101, 5), 15: (0b0000101, 7), 16: (0b010101, 6),
              17: (0b100101, 6), 18: (0b1001, 4)}
# mode id -> (bit position of the partition field, bits, number of legal partitions)
PATTERN_FIELD = {2: (20, 5, 30), 3: (20, 4, 11), 4: (20, 5, 30), 7: (20, 5, 19), 9: (28, 5, 30), 16: (29, 5, 30)}


def _set_bits(lo, pos, nbits, val):
    """lo: uint64 array holding the low 64 bits of each block."""
    mask = np.uint64(((1 << nbits) - 1) << pos)
    return (lo & ~mask) | ((val.astype(np.uint64) << np.uint64(pos)) & mask)


def random_blocks(n, seed=1, modes=None, invalid_fraction=0.0):
    """n valid-random UASTC blocks as an (n,16) uint8 array.  `modes` restricts the mode set.
    invalid_fraction > 0 makes that share of blocks invalid (mode code 69 or a bad partition)."""
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 16), dtype=np.uint8)
    lo = raw[:, :8].copy().view(np.uint64).reshape(n)
    mode_list = np.array(sorted(MODE_CODES) if modes is None else list(modes))
    m = mode_list[rng.integers(0, len(mode_list), size=n)]
    for mode, (code, bits) in MODE_CODES.items():
        sel = m == mode
        if sel.any():
            lo[sel] = _set_bits(lo[sel], 0, bits, np.full(sel.sum(), code))
    for mode, (pos, bits, count) in PATTERN_FIELD.items():
        sel = m == mode
        if sel.any():
            lo[sel] = _set_bits(lo[sel], pos, bits, rng.integers(0, count, size=sel.sum()))
    # void-extent: make channels hit 0 and 255 often (BC7 mode 5 needs both in one colour)
    sel = np.where(m == 8)[0]
    if len(sel):
        ch = rng.integers(0, 256, size=(len(sel), 4))
        force = rng.integers(0, 4, size=(len(sel), 4))
        ch = np.where(force == 0, 0, np.where(force == 1, 255, ch))
        rgba = (ch[:, 0] | (ch[:, 1] << 8) | (ch[:, 2] << 16) | (ch[:, 3] << 24)).astype(np.uint64)
        lo[sel] = _set_bits(lo[sel], 5, 32, rgba)
    if invalid_fraction > 0:
        bad = np.where(rng.random(n) < invalid_fraction)[0]
        half = bad[: len(bad) // 2]
        lo[half] = _set_bits(lo[half], 0, 7, np.full(len(half), 69))          # the one invalid mode code
        rest = bad[len(bad) // 2:]
        for i in rest:                                                          # out-of-range partition
            mode = int(rng.choice(list(PATTERN_FIELD)))
            code, bits = MODE_CODES[mode]
            pos, pbits, count = PATTERN_FIELD[mode]
            v = np.array([lo[i]], dtype=np.uint64)
            v = _set_bits(v, 0, bits, np.array([code]))
            hi_vals = list(range(count, 1 << pbits))
            v = _set_bits(v, pos, pbits, np.array([rng.choice(hi_vals)]))
            lo[i] = v[0]
    out = raw.copy()
    out[:, :8] = lo.view(np.uint8).reshape(n, 8)
    return out
