"""Synthetic ETC1S payloads (test + bench inputs): seeded codebooks and spatially coherent index maps,
encoded with the test-only encoder in oracle/etc1s_encoder.inc (no ETC1S data exists in the reference)."""
import ctypes

import numpy as np


def make_codebooks(n_endpoints, n_selectors, seed=0):
    rng = np.random.default_rng(seed)
    ep = np.zeros((n_endpoints, 4), dtype=np.uint8)
    ep[:, 0] = rng.integers(0, 8, n_endpoints)
    base = rng.integers(0, 32, size=(n_endpoints, 3))
    # neighbouring codebook entries are similar (DPCM friendly), like a real sorted codebook
    walk = np.cumsum(rng.integers(-2, 3, size=(n_endpoints, 3)), axis=0)
    ep[:, 1:] = ((base // 4) + walk) % 32
    sel = rng.integers(0, 256, size=(n_selectors, 4), dtype=np.uint8)
    sel[::3] = (sel[::3] & 0xF0) | (sel[::3] >> 4)          # some structure for the XOR-DPCM coder
    return ep, sel


def make_indices(nbx, nby, num_slices, n_endpoints, n_selectors, seed=0, flat=0.35):
    """value-noise style maps with flat regions so that predictor reuse, RLE and history hits all occur."""
    rng = np.random.default_rng(seed)
    n = nbx * nby
    ep = np.empty((num_slices, nby, nbx), dtype=np.uint16)
    se = np.empty((num_slices, nby, nbx), dtype=np.uint16)
    for k in range(num_slices):
        gy, gx = (nby + 7) // 8 + 1, (nbx + 7) // 8 + 1
        coarse_e = rng.integers(0, n_endpoints, size=(gy, gx))
        coarse_s = rng.integers(0, n_selectors, size=(gy, gx))
        e = np.kron(coarse_e, np.ones((8, 8), dtype=np.int64))[:nby, :nbx]
        s = np.kron(coarse_s, np.ones((8, 8), dtype=np.int64))[:nby, :nbx]
        noise = rng.random((nby, nbx)) > flat
        e = np.where(noise, (e + rng.integers(-3, 4, size=(nby, nbx))) % n_endpoints, e)
        noise2 = rng.random((nby, nbx)) > (flat + 0.25)
        s = np.where(noise2, rng.integers(0, n_selectors, size=(nby, nbx)), s)
        # long flat runs (exercise the VLC-coded run lengths)
        if nby > 4:
            e[nby // 2] = e[nby // 2, 0]
            s[nby // 2] = s[nby // 2, 0]
        ep[k], se[k] = e, s
    return ep.reshape(num_slices, n), se.reshape(num_slices, n)


def encode(orc, ep_cb, sel_cb, ep_idx, sel_idx, nbx, nby, hist_size=64, raw_selectors=False, is_video=False):
    """Returns dict(endpoints, selectors, tables, slice_data, slice_ofs, slice_len)."""
    c = ctypes
    num_slices = ep_idx.shape[0]
    ep_idx = np.ascontiguousarray(ep_idx, dtype=np.uint16)
    sel_idx = np.ascontiguousarray(sel_idx, dtype=np.uint16)
    ptrs = [c.POINTER(c.c_uint8)() for _ in range(4)]
    lens = [c.c_size_t(0) for _ in range(4)]
    ofs = (c.c_uint64 * num_slices)()
    sl = (c.c_uint64 * num_slices)()
    orc.orc_etc1s_encode.argtypes = [c.c_void_p, c.c_uint, c.c_void_p, c.c_uint, c.c_int, c.c_uint, c.c_int, c.c_uint, c.c_uint, c.c_uint,
                                     c.c_void_p, c.c_void_p] + [c.c_void_p] * 8 + [c.c_void_p, c.c_void_p]
    e = orc.orc_etc1s_encode(ep_cb.ctypes.data, ep_cb.shape[0], sel_cb.ctypes.data, sel_cb.shape[0], int(raw_selectors), hist_size,
                             int(is_video), nbx, nby, num_slices, ep_idx.ctypes.data, sel_idx.ctypes.data,
                             c.byref(ptrs[0]), c.byref(lens[0]), c.byref(ptrs[1]), c.byref(lens[1]), c.byref(ptrs[2]), c.byref(lens[2]),
                             c.byref(ptrs[3]), c.byref(lens[3]), ofs, sl)
    assert e == 0, e
    out = {}
    for name, p, ln in zip(("endpoints", "selectors", "tables", "slice_data"), ptrs, lens):
        out[name] = c.string_at(p, ln.value)
        orc.orc_free.argtypes = [c.c_void_p]
        orc.orc_free(p)
    out["slice_ofs"] = list(ofs)
    out["slice_len"] = list(sl)
    return out
