"""Cross-witness checks shared by tests/test_spec_witness.py (oracle outputs, CPU) and tests/test_gpu_witness.py (CUDA outputs).

The witnesses are the decoders of the TARGET formats in tests/spec/spec_decoders.c, written from the public format
specifications.  Each check takes transcoded blocks from whichever implementation is under test and the unpacked RGBA of the
same UASTC blocks, and states what the format relationship demands."""
import numpy as np

from uastc_synth import MODE_CODES, PATTERN_FIELD, random_blocks

# UASTC mode -> BC7 mode (reference src/target_formats/bc7.rs:582-589)
BC7_MODE_OF = {0: 6, 1: 3, 2: 1, 3: 2, 4: 3, 5: 6, 6: 5, 7: 2, 9: 7, 10: 6, 11: 5, 12: 6, 13: 5, 14: 6, 15: 6, 16: 7, 17: 5, 18: 6}


def uastc_modes(blocks):
    """mode id of every block from its prefix code (the synthesiser's own table)"""
    lo = blocks[:, 0].astype(np.int32)
    out = np.full(len(blocks), -1)
    for m, (code, bits) in MODE_CODES.items():
        out[(lo & ((1 << bits) - 1)) == code] = m
    return out


def all_partition_blocks(per_combo=24, seed=77):
    """valid-random blocks that cover EVERY (mode, partition index) combination at least per_combo times, plus every
    unpartitioned mode: 30 + 11 + 30 + 19 + 30 + 30 partitioned combinations (the KATs cover 44 of the 60 partitions)"""
    rng = np.random.default_rng(seed)
    parts = []
    for mode, (pos, bits, count) in PATTERN_FIELD.items():
        blk = random_blocks(per_combo * count, seed=int(rng.integers(1 << 30)), modes=[mode])
        lo = blk[:, :8].copy().view(np.uint64).reshape(-1)
        mask = np.uint64(((1 << bits) - 1) << pos)
        pat = np.repeat(np.arange(count, dtype=np.uint64), per_combo)
        lo = (lo & ~mask) | (pat << np.uint64(pos))
        blk[:, :8] = lo.view(np.uint8).reshape(-1, 8)
        parts.append(blk)
    parts.append(random_blocks(4000, seed=int(rng.integers(1 << 30)), modes=[m for m in MODE_CODES if m not in PATTERN_FIELD]))
    return np.ascontiguousarray(np.concatenate(parts))


def check_astc_lossless(spec, astc_blocks, rgba_blocks):
    """UASTC -> ASTC is lossless (reference src/target_formats/astc.rs:8-181): a spec ASTC decoder must return the texels of
    decode_block_to_rgba (uastc.rs:237-327) exactly -- endpoints, BISE re-encoding, blue-contraction swap, weight order and
    inversion, partition seed and dual-plane selector all have to be right for that."""
    rc, dec = spec.astc(astc_blocks)
    assert (rc == 0).all(), "the spec decoder rejects blocks: codes %s" % np.unique(rc)
    rgba = np.asarray(rgba_blocks, dtype=np.uint8).reshape(-1, 64)
    bad = np.where((dec != rgba).any(axis=1))[0]
    assert len(bad) == 0, "ASTC output of %d blocks does not decode to the RGBA output (first: %d)" % (len(bad), bad[0])


def check_bc7_single_subset(spec, uastc, bc7_blocks, rgba_blocks, max_abs=10, mean_abs=1.0):
    """UASTC modes that become BC7 mode 5 / 6 (one subset): the mode byte must be the mapped one and the spec decode must
    stay within the re-quantisation error of the reference's texels (7-bit + p-bit endpoints, weight LUTs 3/5 -> 4 bits).
    A wrong weight order, channel rotation or endpoint order is an error of tens of levels, not of a few."""
    modes = uastc_modes(uastc)
    bc7 = np.asarray(bc7_blocks, dtype=np.uint8).reshape(-1, 16)
    rgba = np.asarray(rgba_blocks, dtype=np.uint8).reshape(-1, 64).astype(np.int32)
    got_modes = spec.bc7_modes(bc7)
    for m, bm in BC7_MODE_OF.items():
        sel = modes == m
        assert (got_modes[sel] == bm).all(), "UASTC mode %d must become BC7 mode %d" % (m, bm)
    sel = np.isin(modes, [m for m, bm in BC7_MODE_OF.items() if bm in (5, 6)])
    rc, dec = spec.bc7_56(bc7[sel])
    assert (rc == 0).all()
    err = np.abs(dec.astype(np.int32) - rgba[sel])
    assert err.max() <= max_abs, "BC7 mode 5/6 texel error %d" % err.max()
    assert err.mean() <= mean_abs, "BC7 mode 5/6 mean texel error %.3f" % err.mean()
    return int(sel.sum())


def _best_err_tables():
    """smallest |decoded - c| a BC7 mode-6 block with all weights = 5 / a mode-5 block with colour weights = 1 can reach for
    a channel value c (brute force over both 7-bit endpoints; BPTC interpolation, weights 21/64)"""
    l, h = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
    c = np.arange(256)[:, None, None]
    best6 = np.zeros((2, 256), dtype=np.int32)
    for p in range(2):
        v = ((64 - 21) * ((l << 1) | p) + 21 * ((h << 1) | p) + 32) >> 6
        best6[p] = np.abs(v[None] - c).min(axis=(1, 2))
    v5 = ((64 - 21) * ((l << 1) | (l >> 6)) + 21 * ((h << 1) | (h >> 6)) + 32) >> 6
    best5 = np.abs(v5[None] - c).min(axis=(1, 2))
    return best6, best5


def check_bc7_void_extent(spec, uastc, bc7_blocks):
    """Solid-colour blocks (reference src/target_formats/bc7.rs:312-375): mode 6 with every weight = 5, or mode 5 when the
    colour has both a 0 and a 255 channel.  The decoded block must be ONE colour and every channel must sit at the smallest
    error that mode / p-bit can reach at all (the 'optimal endpoint' tables, re-derived here by brute force)."""
    modes = uastc_modes(uastc)
    sel = modes == 8
    blk = uastc[sel]
    lo = blk[:, :8].copy().view(np.uint64).reshape(-1)
    rgba = ((lo >> np.uint64(5)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    col = np.stack([(rgba >> (8 * i)) & 0xFF for i in range(4)], axis=1).astype(np.int32)
    bc7 = np.asarray(bc7_blocks, dtype=np.uint8).reshape(-1, 16)[sel]
    rc, dec = spec.bc7_56(bc7)
    assert (rc == 0).all()
    dec = dec.reshape(-1, 16, 4).astype(np.int32)
    assert (dec == dec[:, :1]).all(), "a void-extent block must decode to a single colour"
    got = dec[:, 0]
    bmode = spec.bc7_modes(bc7)
    want5 = ((col == 0).any(axis=1)) & ((col == 255).any(axis=1))
    assert (bmode[want5] == 5).all() and (bmode[~want5] == 6).all()
    best6, best5 = _best_err_tables()
    err = np.abs(got - col)
    m5 = bmode == 5
    assert (err[m5][:, :3] == best5[col[m5][:, :3]]).all() and (err[m5][:, 3] == 0).all()
    # mode 6: the two p-bits are equal (bc7.rs:362-371); read them from the block: bits 63 and 64
    b6 = bc7[~m5]
    p0 = (b6[:, 7] >> 7) & 1
    p1 = b6[:, 8] & 1
    assert (p0 == p1).all()
    e6 = err[~m5]
    assert (e6 == best6[p0[:, None], col[~m5]]).all(), "mode 6 void-extent colour is not at the optimal-endpoint error"
    # ... and the p-bit is the better one of the two (ties: 0), by total error over the four channels as the reference sums it
    tot = np.stack([best6[p][col[~m5]].astype(np.int64) for p in range(2)], axis=0)
    sq = (tot ** 2).sum(axis=2)
    ab = tot.sum(axis=2)
    better_sq = np.where(sq[1] < sq[0], 1, 0)
    better_ab = np.where(ab[1] < ab[0], 1, 0)
    assert ((p0 == better_sq) | (p0 == better_ab)).all()
    return int(sel.sum()), int(m5.sum())


def check_etc1_matches_rgba(spec, etc1_blocks, rgba_image, nbx, nby, alpha_is_255=True):
    """ETC1S -> ETC1 is lossless too (reference src/basis_lz/mod.rs:153-186 writes base colour, table and selectors verbatim):
    a spec ETC1 decode of the transcoded slice must equal the RGB of decode_to_rgba (mod.rs:97-151)."""
    rc, dec = spec.etc1(np.frombuffer(etc1_blocks, dtype=np.uint8).reshape(-1, 8))
    assert (rc == 0).all()
    img = np.frombuffer(rgba_image, dtype=np.uint8).reshape(nby, 4, nbx, 4, 4).transpose(0, 2, 1, 3, 4).reshape(nbx * nby, 16, 4)
    dec = dec.reshape(-1, 16, 4)
    assert (dec[:, :, :3] == img[:, :, :3]).all(), "ETC1 output does not decode to the RGBA output"
    if alpha_is_255:
        assert (img[:, :, 3] == 255).all()
