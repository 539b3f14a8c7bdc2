"""GPU cross-witness tests (pytest -m gpu): the CUDA outputs, obtained through the C ABI, examined by the spec-derived
target-format decoders of tests/spec (written from the public specifications, validated on the reference's own vectors in
tests/test_spec_witness.py) and by the hand-assembled ETC1S stream -- i.e. by something that is neither the oracle nor the
kernels.  Plus the exhaustive sweep that licenses the table / integer form of BC7's shared p-bit search."""
import numpy as np
import pytest

import etc1s_common as ec
import witness_checks as wc
from conftest import OUT_BYTES, TARGETS, oracle_transcode
from uastc_synth import random_blocks

pytestmark = pytest.mark.gpu


def gpu_transcode(b, target, blocks, bpr=1):
    raw = np.ascontiguousarray(blocks, dtype=np.uint8).tobytes()
    if target == 0:
        return np.frombuffer(b.uastc_decode_rgba(raw, bpr), dtype=np.uint8)
    return np.frombuffer(b.uastc_transcode(target, raw), dtype=np.uint8)


def test_astc_output_decodes_to_the_rgba_output_for_all_60_partitions(gpu_lib, spec):
    """UASTC -> ASTC is lossless: spec ASTC decode of the CUDA ASTC blocks == CUDA RGBA texels, on blocks covering every
    (mode, partition) combination (the golden vectors cover 44 of the 60 partitions)."""
    blk = wc.all_partition_blocks()
    n = len(blk)
    pad = (-n) % 64                                                      # a whole number of block rows for the RGBA image
    blk = np.concatenate([blk, blk[:pad]])
    astc = gpu_transcode(gpu_lib, TARGETS["astc"], blk)
    img = gpu_transcode(gpu_lib, TARGETS["rgba"], blk, bpr=64)
    rgba = img.reshape(len(blk) // 64, 4, 64, 16).transpose(0, 2, 1, 3).reshape(len(blk), 64)
    wc.check_astc_lossless(spec, astc.reshape(-1, 16), rgba)


def test_bc7_output_against_the_bptc_witness(gpu_lib, spec):
    blk = wc.all_partition_blocks(per_combo=4)
    blk = np.concatenate([blk, blk[:(-len(blk)) % 64]])
    bc7 = gpu_transcode(gpu_lib, TARGETS["bc7"], blk)
    img = gpu_transcode(gpu_lib, TARGETS["rgba"], blk, bpr=64)
    rgba = img.reshape(len(blk) // 64, 4, 64, 16).transpose(0, 2, 1, 3).reshape(len(blk), 64)
    assert wc.check_bc7_single_subset(spec, blk, bc7.reshape(-1, 16), rgba) > 1000
    ve = random_blocks(20000, seed=5, modes=[8])
    n, n5 = wc.check_bc7_void_extent(spec, ve, gpu_transcode(gpu_lib, TARGETS["bc7"], ve).reshape(-1, 16))
    assert n == 20000 and n5 > 1000


def test_etc2_colour_half_is_the_etc1_block_and_alpha_tracks_rgba(gpu_lib, spec, kat):
    blk = random_blocks(20000, seed=13)
    etc1 = gpu_transcode(gpu_lib, TARGETS["etc1"], blk).reshape(-1, 8)
    etc2 = gpu_transcode(gpu_lib, TARGETS["etc2"], blk).reshape(-1, 16)
    assert (etc2[:, 8:] == etc1).all()                                   # etc.rs:22-28
    a = spec.eac_alpha(np.ascontiguousarray(etc2[:, :8]))
    img = gpu_transcode(gpu_lib, TARGETS["rgba"], blk, bpr=100)
    alpha = img.reshape(200, 4, 100, 16).transpose(0, 2, 1, 3).reshape(20000, 16, 4)[:, :, 3]
    opaque = (alpha == 255).all(axis=1)
    assert opaque.sum() > 5000 and (a[opaque] == 255).all()             # RGB modes: solid 255 alpha block


@pytest.mark.parametrize("shape", [(48, 40, 2, 512, 64, False), (33, 7, 3, 300, 64, True), (64, 64, 1, 4096, 0, False)])
def test_etc1s_etc1_output_decodes_to_the_rgba_output(gpu_lib, spec, oracle, shape):
    """the encoder that makes the stream is ours (test infrastructure); the CHECK is not: ETC1 specification decode of the CUDA
    ETC1 output == CUDA RGBA output"""
    nbx, nby, ns, ncb, hist, raw = shape
    eo = ec.bind(oracle)
    _, _, _, _, enc = ec.make_case(eo, nbx, nby, ns, ncb, hist=hist, raw=raw, seed=9)
    dec = gpu_lib.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"])
    for k in range(ns):
        d = ec.slice_bytes(enc, k)
        wc.check_etc1_matches_rgba(spec, dec.transcode_to_etc1(nbx, nby, d), dec.decode_to_rgba(nbx, nby, d), nbx, nby)
    dec.close()


@pytest.mark.parametrize("which", ["build", "build_runs"])
def test_hand_assembled_etc1s_stream_on_the_gpu(gpu_lib, spec, which):
    """the bit-by-bit hand-made streams of tests/etc1s_handmade.py through b2bu_etc1s_open / K2 / K3: hand-derived texels"""
    import etc1s_handmade as hm
    case = getattr(hm, which)()
    want = hm.expected_rgba(case)
    dec = gpu_lib.Etc1sDecoder(case["n"], case["n"], case["endpoints"], case["selectors"], case["tables"])
    nbx, nby = case["nbx"], case["nby"]
    assert dec.decode_to_rgba(nbx, nby, case["slice"]) == want.tobytes()
    wc.check_etc1_matches_rgba(spec, dec.transcode_to_etc1(nbx, nby, case["slice"]), want.tobytes(), nbx, nby)
    dec.close()


def test_bc7_shared_pbits_exhaustive_over_all_16_pow_6_endpoint_sets(gpu_lib, oracle):
    """UASTC mode 2 -> BC7 mode 1 decides one shared p-bit per subset from the subset's six 4-bit endpoint values with an f32
    error search (reference src/target_formats/bc7.rs:408-475).  The kernel uses a table of the reference's f32 terms; SURVEY.md
    section 7 asks for the whole domain: all 16^6 = 16,777,216 (r0,r1,g0,g1,b0,b1) sets, two per block, GPU == oracle."""
    n = 16 ** 6 // 2
    i = np.arange(n, dtype=np.uint64)
    blk = np.zeros((n, 2), dtype=np.uint64)
    # mode 2: code 0b11101 (5 bits), 15 flag bits, 5-bit partition at bit 20, twelve 4-bit endpoints from bit 25
    # (subset 0: r0 r1 g0 g1 b0 b1, then subset 1), 46 weight bits from bit 73
    pat = i % np.uint64(30)
    ep = (i * np.uint64(2)) | (((i * np.uint64(2)) + np.uint64(1)) << np.uint64(24))      # 48 bits: combination 2i, then 2i+1
    w = (i * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(18)                              # 46 pseudo-random weight bits (anchor MSBs vary)
    blk[:, 0] = np.uint64(0b11101) | (pat << np.uint64(20)) | (ep << np.uint64(25))
    blk[:, 1] = (ep >> np.uint64(39)) | (w << np.uint64(9))
    raw = blk.view(np.uint8).reshape(n, 16)
    got = gpu_transcode(gpu_lib, TARGETS["bc7"], raw)
    e, _, want = oracle_transcode(oracle, TARGETS["bc7"], raw, threads=16)
    assert e == 0
    bad = np.where((got.reshape(n, 16) != want.reshape(n, 16)).any(axis=1))[0]
    assert len(bad) == 0, "%d of %d blocks differ, first %d" % (len(bad), n, bad[0])
    # the sweep really reached both p-bit values for both subsets: BC7 mode 1 stores them after 6 + 2*2*3*6 = 78 endpoint bits + 8 header bits
    pb = (got.reshape(n, 16)[:, 10] >> 0) & 3                                             # bits 80, 81
    assert set(np.unique(pb)) == {0, 1, 2, 3}
