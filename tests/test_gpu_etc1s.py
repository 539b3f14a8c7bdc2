"""GPU parity tests of the ETC1S / BasisLZ path (K2 entropy decode + K3 gather) against the CPU oracle
on generated bitstreams.  PARITY UNPINNED against the reference itself: it ships no ETC1S vector."""
import ctypes

import numpy as np
import pytest

from etc1s_common import bind, decode_bc1, etc1s_file, make_case, oracle_bc1, oracle_etc1, oracle_open, oracle_read_to, oracle_rgba, slice_bytes

pytestmark = pytest.mark.gpu

CASES = [  # nbx, nby, slices, codebook size, history, raw selectors, video
    (16, 16, 2, 64, 64, False, False), (33, 17, 3, 300, 64, True, False), (1, 1, 1, 4, 0, False, False), (2, 1, 1, 4, 64, False, False),
    (64, 64, 2, 4096, 64, False, False), (7, 5, 2, 50, 16, False, True), (128, 96, 2, 1000, 200, False, False), (31, 1, 1, 20, 64, False, False),
    (1, 40, 1, 20, 64, False, False), (257, 130, 1, 8000, 64, False, False),
    (60001, 3, 1, 300, 64, False, False),          # wider than the shared-memory row state: previous-row state in the global scratch area
]


@pytest.mark.parametrize("nbx,nby,ns,ncb,hist,raw,video", CASES)
def test_slice_api_matches_oracle(gpu_lib, oracle, nbx, nby, ns, ncb, hist, raw, video):
    orc = bind(oracle)
    _, _, ei, si, enc = make_case(orc, nbx, nby, ns, ncb, hist, raw, video, seed=nbx + nby)
    e, h = oracle_open(orc, enc, ncb, ncb, video)
    assert e == 0
    dec = gpu_lib.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"], b"", video)
    for k in range(ns):
        d = slice_bytes(enc, k)
        e, want = oracle_etc1(orc, h, nbx, nby, d)
        assert e == 0
        assert dec.transcode_to_etc1(nbx, nby, d) == want
        e, want = oracle_rgba(orc, h, nbx, nby, d)
        assert dec.decode_to_rgba(nbx, nby, d) == want
    if ns >= 2:
        e, want = oracle_rgba(orc, h, nbx, nby, slice_bytes(enc, 0), slice_bytes(enc, 1))
        assert e == 0
        assert dec.decode_to_rgba(nbx, nby, slice_bytes(enc, 0), slice_bytes(enc, 1)) == want
    dec.close()
    orc.orc_etc1s_close(h)


def test_batched_slices_entry_point(gpu_lib, oracle):
    orc = bind(oracle)
    nbx, nby, ns, ncb = 64, 48, 12, 2048
    _, _, ei, si, enc = make_case(orc, nbx, nby, ns, ncb, seed=3)
    e, h = oracle_open(orc, enc, ncb, ncb)
    dec = gpu_lib.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"])
    L = gpu_lib.lib()
    ofs = (ctypes.c_uint64 * ns)(*enc["slice_ofs"])
    ln = (ctypes.c_uint64 * ns)(*enc["slice_len"])
    for target, per in ((gpu_lib.ETC1, 8), (gpu_lib.RGBA, 64)):
        out = np.zeros(ns * nbx * nby * per, dtype=np.uint8)
        st = L.b2bu_etc1s_transcode_slices(dec._h, target, nbx, nby, enc["slice_data"], len(enc["slice_data"]), ofs, ln, ns, out.ctypes.data, out.size)
        assert st == 0
        for k in range(ns):
            fn = oracle_etc1 if target == gpu_lib.ETC1 else oracle_rgba
            e, want = fn(orc, h, nbx, nby, slice_bytes(enc, k))
            assert out[k * nbx * nby * per:(k + 1) * nbx * nby * per].tobytes() == want, (target, k)
    dec.close()
    orc.orc_etc1s_close(h)


def test_file_level_read_to_rgba_and_etc1(gpu_lib, oracle):
    orc = bind(oracle)
    for alpha in (False, True):
        nbx, nby, ncb = 21, 10, 256
        _, _, ei, si, enc = make_case(orc, nbx, nby, 4, ncb, seed=17)
        f = etc1s_file(enc, nbx, nby, ncb, alpha_pairs=alpha, orig=(83, 39))
        e, want = oracle_read_to(orc, 0, f)
        assert e == 0
        header, images = gpu_lib.read_to_rgba(f)
        assert header.texture_format() == "ETC1S" and header.has_alpha() == alpha
        assert [(im.w, im.h, im.stride, im.data) for im in images] == want
        e, want = oracle_read_to(orc, 3, f)
        images = gpu_lib.read_to_etc1(f)
        assert [(im.w, im.h, im.stride, im.data) for im in images] == want
        with pytest.raises(gpu_lib.BasisuError) as ei_:
            gpu_lib.read_to_bc7(f)                       # reference: unimplemented!() for ETC1S files (basis.rs:258)
        assert ei_.value.status == 14


def test_file_level_large_file_device_crc_path(gpu_lib, oracle):
    """An ETC1S file above the 256 KiB threshold: one upload, CRC-16 on the GPU, slices gathered device-to-device."""
    orc = bind(oracle)
    nbx, nby, ncb = 200, 160, 4000
    _, _, ei, si, enc = make_case(orc, nbx, nby, 6, ncb, seed=29)
    f = etc1s_file(enc, nbx, nby, ncb, alpha_pairs=True)
    assert len(f) > 256 * 1024
    e, want = oracle_read_to(orc, 0, f)
    assert e == 0
    header, images = gpu_lib.read_to_rgba(f)
    assert [(im.w, im.h, im.stride, im.data) for im in images] == want
    e, want = oracle_read_to(orc, 3, f)
    assert [(im.w, im.h, im.stride, im.data) for im in gpu_lib.read_to_etc1(f)] == want
    g = bytearray(f); g[len(g) - 1000] ^= 0x11
    with pytest.raises(gpu_lib.BasisuError) as ei_:
        gpu_lib.read_to_etc1(bytes(g))
    assert str(ei_.value) == "Data CRC16 failed"
    g = bytearray(f); g[100] ^= 0x11                      # inside the slice descriptors / codebooks: the CRC error still wins
    with pytest.raises(gpu_lib.BasisuError) as ei_:
        gpu_lib.read_to_rgba(bytes(g))
    assert str(ei_.value) == "Data CRC16 failed"


@pytest.mark.parametrize("nbx,nby,ncb", [(64, 64, 4096), (33, 17, 300), (1, 1, 4)])
def test_bc1_extension_matches_its_definition(gpu_lib, oracle, nbx, nby, ncb):
    """ETC1S -> BC1 does not exist in the reference: the definition is the oracle's etc1s_emit_bc1; the kernel must
    reproduce it bit for bit (the sanity of the definition itself is a CPU test, tests/test_etc1s_oracle.py)."""
    orc = bind(oracle)
    _, _, ei, si, enc = make_case(orc, nbx, nby, 2, ncb, seed=nbx * 7 + nby)
    e, h = oracle_open(orc, enc, ncb, ncb)
    dec = gpu_lib.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"])
    for k in range(2):
        d = slice_bytes(enc, k)
        e, want = oracle_bc1(orc, h, nbx, nby, d)
        assert e == 0
        got = dec.transcode_to_bc1(nbx, nby, d)
        assert got == want
    dec.close()
    orc.orc_etc1s_close(h)


def test_corrupt_streams_report_like_the_oracle(gpu_lib, oracle):
    orc = bind(oracle)
    nbx, nby, ncb = 40, 30, 500
    _, _, ei, si, enc = make_case(orc, nbx, nby, 1, ncb, seed=23)
    e, h = oracle_open(orc, enc, ncb, ncb)
    dec = gpu_lib.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"])
    good = slice_bytes(enc, 0)
    rng = np.random.default_rng(1)
    for trial in range(12):
        bad = bytearray(good)
        if trial % 3 == 0:
            bad = bad[: len(bad) // 2]                    # truncated: the reader sees zeros past the end
        else:
            for _ in range(3):
                bad[int(rng.integers(0, len(bad)))] ^= int(rng.integers(1, 256))
        e, want = oracle_etc1(orc, h, nbx, nby, bytes(bad))
        if e == 0:
            assert dec.transcode_to_etc1(nbx, nby, bytes(bad)) == want
        else:
            with pytest.raises(gpu_lib.BasisuError) as ei_:
                dec.transcode_to_etc1(nbx, nby, bytes(bad))
            assert ei_.value.status == e
    dec.close()
    orc.orc_etc1s_close(h)


@pytest.mark.parametrize("nbx,nby,ncb,hist,video", [(64, 20, 300, 64, False), (37, 9, 40, 16, False), (32, 6, 64, 64, True), (95, 4, 2000, 0, False)])
def test_corrupt_stream_fuzz_keeps_the_reference_error_order(gpu_lib, oracle, nbx, nby, ncb, hist, video):
    """The two-warp pipeline finds errors in two places (bit-level in the tokenizer, index-level in the resolver) and must
    still report the one the reference's serial loop meets first; and streams that decode despite the damage must decode
    to the same indices.  Bit flips, byte splices and truncations over several shapes (odd widths, video, no history)."""
    orc = bind(oracle)
    _, _, ei, si, enc = make_case(orc, nbx, nby, 1, ncb, hist, False, video, seed=nbx + 3 * nby)
    e, h = oracle_open(orc, enc, ncb, ncb, video)
    assert e == 0
    dec = gpu_lib.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"], b"", video)
    good = slice_bytes(enc, 0)
    rng = np.random.default_rng(nbx * 1000 + nby)
    seen = set()
    for trial in range(40):
        bad = bytearray(good)
        kind = trial % 4
        if kind == 0:
            bad = bad[: int(rng.integers(0, len(bad)))]
        elif kind == 1:
            bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 2:
            at = int(rng.integers(0, len(bad)))
            bad[at:at + 4] = bytes(rng.integers(0, 256, size=4, dtype=np.uint8))
        else:
            at = int(rng.integers(0, max(1, len(bad) - 8)))
            bad[at:at + 8] = b"\xff" * 8
        e, want = oracle_etc1(orc, h, nbx, nby, bytes(bad))
        seen.add(e)
        if e == 0:
            assert dec.transcode_to_etc1(nbx, nby, bytes(bad)) == want, trial
        else:
            with pytest.raises(gpu_lib.BasisuError) as ei_:
                dec.transcode_to_etc1(nbx, nby, bytes(bad))
            assert ei_.value.status == e, (trial, kind)
    assert len(seen) >= 2
    dec.close()
    orc.orc_etc1s_close(h)


def test_config4_shape_slices_property(gpu_lib, oracle):
    """BASELINE config 4 shape (1024x1024 blocks per slice) on a few slices: decode -> re-encode round trip
    is not available, so check against the oracle on one slice and self-consistency (ETC1 vs RGBA colours)."""
    orc = bind(oracle)
    nbx = nby = 1024
    ncb = 4096
    _, _, ei, si, enc = make_case(orc, nbx, nby, 2, ncb, seed=41)
    e, h = oracle_open(orc, enc, ncb, ncb)
    dec = gpu_lib.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"])
    d = slice_bytes(enc, 0)
    e, want = oracle_etc1(orc, h, nbx, nby, d)
    got = dec.transcode_to_etc1(nbx, nby, d)
    assert got == want
    e, want = oracle_rgba(orc, h, nbx, nby, slice_bytes(enc, 1))
    assert dec.decode_to_rgba(nbx, nby, slice_bytes(enc, 1)) == want
    dec.close()
    orc.orc_etc1s_close(h)


@pytest.mark.parametrize("nbx,nby,ns,ncb", [(64, 33, 300, 700), (40, 40, 1300, 256)])
def test_many_slices_pack_several_pipelines_per_cta(gpu_lib, oracle, nbx, nby, ns, ncb):
    """More slices than SMs: the launch packs up to 8 two-warp pipelines into a CTA (shared tables, the narrow table set) and
    runs several waves.  Every slice must still decode exactly, run after run."""
    orc = bind(oracle)
    _, _, ei, si, enc = make_case(orc, nbx, nby, 10, ncb, seed=ns)
    uniq = len(enc["slice_ofs"])
    e, h = oracle_open(orc, enc, ncb, ncb)
    want = [oracle_etc1(orc, h, nbx, nby, slice_bytes(enc, k))[1] for k in range(uniq)]
    dec = gpu_lib.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"])
    L = gpu_lib.lib()
    ofs = (ctypes.c_uint64 * ns)(*[enc["slice_ofs"][k % uniq] for k in range(ns)])
    ln = (ctypes.c_uint64 * ns)(*[enc["slice_len"][k % uniq] for k in range(ns)])
    per = nbx * nby * 8
    out = np.zeros(ns * per, dtype=np.uint8)
    for rep in range(3):
        out[:] = 0
        assert L.b2bu_etc1s_transcode_slices(dec._h, gpu_lib.ETC1, nbx, nby, enc["slice_data"], len(enc["slice_data"]), ofs, ln, ns,
                                             out.ctypes.data, out.size) == 0
        for k in range(ns):
            assert out[k * per:(k + 1) * per].tobytes() == want[k % uniq], (rep, k)
    dec.close()
    orc.orc_etc1s_close(h)
