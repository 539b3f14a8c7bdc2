"""Shared helpers of the ETC1S tests: oracle bindings + synthetic file construction."""
import ctypes

import numpy as np

from basis_writer import build_basis
from etc1s_synth import encode, make_codebooks, make_indices

c = ctypes


def bind(orc):
    orc.orc_etc1s_open.argtypes = [c.c_uint, c.c_uint, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_int,
                                   c.POINTER(c.c_void_p)]
    orc.orc_etc1s_close.argtypes = [c.c_void_p]
    orc.orc_etc1s_decode_indices.argtypes = [c.c_void_p, c.c_uint, c.c_uint, c.c_void_p, c.c_size_t, c.c_void_p, c.c_void_p]
    orc.orc_etc1s_codebooks.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
    orc.orc_etc1s_transcode_to_etc1.argtypes = [c.c_void_p, c.c_uint, c.c_uint, c.c_void_p, c.c_size_t, c.c_void_p]
    orc.orc_etc1s_transcode_to_bc1.argtypes = [c.c_void_p, c.c_uint, c.c_uint, c.c_void_p, c.c_size_t, c.c_void_p]
    orc.orc_etc1s_decode_to_rgba.argtypes = [c.c_void_p, c.c_uint, c.c_uint, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_void_p]
    orc.orc_read_to.argtypes = [c.c_int, c.c_void_p, c.c_size_t, c.c_void_p, c.c_uint32, c.POINTER(c.c_uint32), c.c_void_p, c.c_uint64,
                                c.POINTER(c.c_uint64), c.c_int]
    return orc


class OrcImage(c.Structure):
    _fields_ = [("w", c.c_uint32), ("h", c.c_uint32), ("stride", c.c_uint32), ("nbytes", c.c_uint64)]


def oracle_open(orc, enc, n_ep, n_sel, is_video=False):
    h = c.c_void_p()
    e = orc.orc_etc1s_open(n_ep, n_sel, enc["endpoints"], len(enc["endpoints"]), enc["selectors"], len(enc["selectors"]),
                           enc["tables"], len(enc["tables"]), int(is_video), c.byref(h))
    return e, h


def slice_bytes(enc, k):
    return enc["slice_data"][enc["slice_ofs"][k]:enc["slice_ofs"][k] + enc["slice_len"][k]]


def oracle_etc1(orc, h, nbx, nby, data):
    out = np.zeros(nbx * nby * 8, dtype=np.uint8)
    e = orc.orc_etc1s_transcode_to_etc1(h, nbx, nby, data, len(data), out.ctypes.data)
    return e, out.tobytes()


def oracle_bc1(orc, h, nbx, nby, data):
    out = np.zeros(nbx * nby * 8, dtype=np.uint8)
    e = orc.orc_etc1s_transcode_to_bc1(h, nbx, nby, data, len(data), out.ctypes.data)
    return e, out.tobytes()


def decode_bc1(blocks: bytes, nbx, nby):
    """Plain BC1 decoder (four-colour and three-colour modes) -> (4 nby, 4 nbx, 3) uint8 image; test helper."""
    b = np.frombuffer(blocks, dtype=np.uint8).reshape(nby, nbx, 8).astype(np.int32)
    c = [b[..., 0] | (b[..., 1] << 8), b[..., 2] | (b[..., 3] << 8)]
    pal = np.zeros((nby, nbx, 4, 3), dtype=np.int32)
    for e in range(2):
        r5, g6, b5 = c[e] >> 11, (c[e] >> 5) & 63, c[e] & 31
        pal[..., e, 0] = (r5 << 3) | (r5 >> 2); pal[..., e, 1] = (g6 << 2) | (g6 >> 4); pal[..., e, 2] = (b5 << 3) | (b5 >> 2)
    four = (c[0] > c[1])[..., None]
    pal[..., 2, :] = np.where(four, (2 * pal[..., 0, :] + pal[..., 1, :]) // 3, (pal[..., 0, :] + pal[..., 1, :]) // 2)
    pal[..., 3, :] = np.where(four, (pal[..., 0, :] + 2 * pal[..., 1, :]) // 3, 0)
    img = np.zeros((nby * 4, nbx * 4, 3), dtype=np.uint8)
    for y in range(4):
        for x in range(4):
            sel = (b[..., 4 + y] >> (2 * x)) & 3
            img[y::4, x::4] = np.take_along_axis(pal, sel[..., None, None].repeat(3, -1), axis=2)[..., 0, :]
    return img


def oracle_rgba(orc, h, nbx, nby, rgb, alpha=None):
    out = np.zeros(nbx * nby * 64, dtype=np.uint8)
    e = orc.orc_etc1s_decode_to_rgba(h, nbx, nby, rgb, len(rgb), alpha, len(alpha) if alpha else 0, out.ctypes.data)
    return e, out.tobytes()


def oracle_read_to(orc, fmt, f):
    cnt = c.c_uint32(0)
    need = c.c_uint64(0)
    imgs = (OrcImage * 64)()
    e = orc.orc_read_to(fmt, f, len(f), imgs, 64, c.byref(cnt), None, 0, c.byref(need), 4)
    if e:
        return e, []
    out = np.zeros(max(need.value, 1), dtype=np.uint8)
    e = orc.orc_read_to(fmt, f, len(f), imgs, 64, c.byref(cnt), out.ctypes.data, need.value, c.byref(need), 4)
    res, pos = [], 0
    for i in range(cnt.value):
        res.append((imgs[i].w, imgs[i].h, imgs[i].stride, out[pos:pos + imgs[i].nbytes].tobytes()))
        pos += imgs[i].nbytes
    return e, res


def make_case(orc, nbx, nby, num_slices, n_cb, hist=64, raw=False, video=False, seed=0):
    """codebooks with n_cb endpoints == n_cb selectors (quirk C-1 safe), indices and the encoded payload."""
    ep_cb, sel_cb = make_codebooks(n_cb, n_cb, seed=seed)
    ei, si = make_indices(nbx, nby, num_slices, n_cb, n_cb, seed=seed + 1)
    if video:
        ei[:, ::5] = 0
        si[:, ::5] = 0
    enc = encode(orc, ep_cb, sel_cb, ei, si, nbx, nby, hist, raw, video)
    return ep_cb, sel_cb, ei, si, enc


def etc1s_file(enc, nbx, nby, n_cb, alpha_pairs=False, video=False, orig=None):
    n = len(enc["slice_ofs"])
    slices = []
    for k in range(n):
        ow, oh = orig if orig else (4 * nbx, 4 * nby)
        slices.append(dict(data=slice_bytes(enc, k), orig_width=ow, orig_height=oh, num_blocks_x=nbx, num_blocks_y=nby,
                           flags=(1 if (alpha_pairs and k % 2 == 1) else 0), image_index=k // 2 if alpha_pairs else k))
    return build_basis(slices, tex_format=0, flags=1 | (4 if alpha_pairs else 0), tex_type=3 if video else 0,
                       etc1s=dict(endpoints=enc["endpoints"], selectors=enc["selectors"], tables=enc["tables"],
                                  total_endpoints=n_cb, total_selectors=n_cb))
