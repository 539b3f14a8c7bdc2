"""CPU tests of the KERNEL SOURCE itself: basisu_rs_b200/csrc/uastc_device.cuh compiled for the host
with shimmed intrinsics (tests/emu).  Catches logic errors before any GPU time is spent; the real
parity gate is tests/test_gpu_uastc.py on the B200."""
import numpy as np
import pytest

from conftest import OUT_BYTES, TARGETS, oracle_transcode, rgba_image_to_blocks
from uastc_synth import random_blocks


def emu_transcode(emu, target, blocks, bpr=1):
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    n = blocks.shape[0]
    out = np.zeros(n * OUT_BYTES[target], dtype=np.uint8)
    st = emu.emu_uastc_transcode(target, blocks.ctypes.data, n, bpr, out.ctypes.data)
    return st, out


@pytest.mark.parametrize("name", list(TARGETS))
def test_kernel_source_reproduces_kat_vectors(emu, kat, name):
    t = TARGETS[name]
    st, out = emu_transcode(emu, t, kat.inputs)
    assert st == 0xFFFFFFFFFFFFFFFF
    assert (out.reshape(kat.n, OUT_BYTES[t]) == kat.expected[t]).all()


@pytest.mark.parametrize("name", list(TARGETS))
def test_kernel_source_matches_oracle_on_random_valid_blocks(emu, oracle, name):
    t = TARGETS[name]
    n, bpr = 60000, 100
    blk = random_blocks(n, seed=11)
    st, a = emu_transcode(emu, t, blk, bpr)
    e, _, b = oracle_transcode(oracle, t, blk, bpr)
    assert st == 0xFFFFFFFFFFFFFFFF and e == 0
    assert (a == b).all()


def test_kernel_source_error_codes(emu, oracle):
    blk = random_blocks(5000, seed=5, invalid_fraction=0.01)
    for t in range(5):
        st, _ = emu_transcode(emu, t, blk, 100)
        e, bad, _ = oracle_transcode(oracle, t, blk, 100, threads=1)
        assert e in (2, 3)
        assert (st >> 8, st & 0xFF) == (bad, e)


def test_eac_selector_threshold_rule_is_exhaustively_equal_to_the_reference_search():
    """uastc_device.cuh etc2_alpha_block replaces the reference's 8-candidate search (etc.rs:317-323, first minimum wins)
    by 7 threshold compares in value order.  Checked here for every table row, multiplier, centre and alpha value."""
    mods = np.array([[-3, -6, -9, -15, 2, 5, 8, 14], [-3, -7, -10, -13, 2, 6, 9, 12], [-2, -5, -8, -13, 1, 4, 7, 12], [-2, -4, -6, -13, 1, 3, 5, 12],
                     [-3, -6, -8, -12, 2, 5, 7, 11], [-3, -7, -9, -11, 2, 6, 8, 10], [-4, -7, -8, -11, 3, 6, 7, 10], [-3, -5, -8, -11, 2, 4, 7, 10],
                     [-2, -6, -8, -10, 1, 5, 7, 9], [-2, -5, -8, -10, 1, 4, 7, 9], [-2, -4, -8, -10, 1, 3, 7, 9], [-2, -5, -7, -10, 1, 4, 6, 9],
                     [-3, -4, -7, -10, 2, 3, 6, 9], [-1, -2, -3, -10, 0, 1, 2, 9], [-4, -6, -8, -9, 3, 5, 7, 8], [-3, -5, -7, -9, 2, 4, 6, 8]])   # etc.rs:451-468
    a = np.arange(256)[None, :, None]                                    # alpha
    order = np.array([3, 2, 1, 0, 4, 5, 6, 7])
    for ti in range(16):
        for mult in range(16):
            centre = np.arange(256)[:, None, None]
            vals = np.clip(centre + mods[ti][None, None, :] * mult, 0, 255)                     # (centre, 1, 8)
            want = np.argmin(np.abs(vals - a), axis=2)                                           # first minimum
            sv = vals[:, :, order]
            s, same = sv[:, :, :-1] + sv[:, :, 1:], sv[:, :, :-1] == sv[:, :, 1:]
            thr = np.where(np.arange(7) < 3, np.where(same, 0, (s + 1) >> 1), (s >> 1) + 1)
            for j in (6, 5, 4, 3):                                      # equal neighbours above k = 0 pass together with the next boundary
                thr[:, :, j] = np.where(same[:, :, j], 256 if j == 6 else thr[:, :, min(j + 1, 6)], thr[:, :, j])
            got = order[(a >= thr).sum(axis=2)]
            assert (got == want).all(), (ti, mult)
