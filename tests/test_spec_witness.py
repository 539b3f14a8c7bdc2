"""Independent witnesses on the CPU (pytest -m "not gpu"): the spec-derived target-format decoders of tests/spec are first
checked against the reference's OWN vectors (tests/golden/uastc_kat.bin comes from the upstream transcoder), then used to
cross-examine the oracle on inputs the golden vectors do not cover (16 of the 60 partitions, BC7 mode-5 void extent, every ETC1S
path).  tests/test_gpu_witness.py repeats the examination on the CUDA outputs."""
import numpy as np
import pytest

import etc1s_common as ec
import witness_checks as wc
from conftest import TARGETS, oracle_transcode, rgba_image_to_blocks


def test_spec_astc_decoder_reproduces_the_reference_vectors(spec, kat):
    """608 reference ASTC blocks decode to the 608 reference RGBA blocks: pins the witness itself"""
    wc.check_astc_lossless(spec, kat.expected[TARGETS["astc"]], kat.expected[TARGETS["rgba"]])


def test_spec_decoders_agree_with_the_reference_etc_and_bc7_vectors(spec, kat):
    rgba = kat.expected[TARGETS["rgba"]].reshape(-1, 16, 4).astype(np.int32)
    # ETC1 / ETC2 are lossy re-encodings: the reference's own blocks must decode NEAR the reference's texels
    rc, e1 = spec.etc1(kat.expected[TARGETS["etc1"]])
    assert (rc == 0).all()
    err = np.abs(e1.reshape(-1, 16, 4)[:, :, :3].astype(np.int32) - rgba[:, :, :3])
    assert err.mean() < 5.0 and np.median(err) <= 3
    etc2 = kat.expected[TARGETS["etc2"]]
    rc, e2 = spec.etc1(np.ascontiguousarray(etc2[:, 8:]))
    assert (rc == 0).all() and (e2 == e1).all()                       # colour half of ETC2 == the ETC1 block
    a = spec.eac_alpha(np.ascontiguousarray(etc2[:, :8])).astype(np.int32)
    aerr = np.abs(a - rgba[:, :, 3])
    assert aerr.mean() < 1.0
    opaque = (rgba[:, :, 3] == 255).all(axis=1)
    assert (a[opaque] == 255).all()
    # BC7: mode mapping and modes 5 / 6 near the reference texels
    n = wc.check_bc7_single_subset(spec, kat.inputs, kat.expected[TARGETS["bc7"]], kat.expected[TARGETS["rgba"]])
    assert n >= 300


def test_astc_hash_partitions_equal_the_uastc_pattern_tables(spec, oracle):
    """The ASTC partition of a transcoded block is whatever the spec's hash makes of the 10-bit seed; UASTC stores a table index.
    For all 60 partitions the oracle's RGBA (pattern tables, uastc.rs:742-811) and the witness' decode of the oracle's ASTC
    (seed tables astc.rs:183-297 + hash) agree texel for texel -- checked through blocks whose subsets have different colours."""
    blk = wc.all_partition_blocks()
    e, _, astc = oracle_transcode(oracle, TARGETS["astc"], blk)
    e2, _, rgba = oracle_transcode(oracle, TARGETS["rgba"], blk, blocks_per_row=1)
    assert e == 0 and e2 == 0
    wc.check_astc_lossless(spec, astc.reshape(-1, 16), rgba.reshape(-1, 64))


def test_oracle_bc7_against_the_witness(spec, oracle):
    blk = wc.all_partition_blocks(per_combo=4)
    e, _, bc7 = oracle_transcode(oracle, TARGETS["bc7"], blk)
    e2, _, rgba = oracle_transcode(oracle, TARGETS["rgba"], blk, blocks_per_row=1)
    assert e == 0 and e2 == 0
    wc.check_bc7_single_subset(spec, blk, bc7.reshape(-1, 16), rgba.reshape(-1, 64))
    from uastc_synth import random_blocks
    ve = random_blocks(20000, seed=5, modes=[8])
    e, _, bc7 = oracle_transcode(oracle, TARGETS["bc7"], ve)
    assert e == 0
    n, n5 = wc.check_bc7_void_extent(spec, ve, bc7.reshape(-1, 16))
    assert n == 20000 and n5 > 1000                                     # the mode-5 path (no golden vector has it) is exercised


@pytest.mark.parametrize("shape", [(48, 40, 2, 512, 64, False), (33, 7, 3, 300, 64, True), (64, 64, 1, 4096, 0, False)])
def test_oracle_etc1s_etc1_decodes_to_its_rgba(spec, oracle, shape):
    """ETC1S: the ETC1 transcode, decoded by the spec ETC1 decoder, is the RGBA decode -- two outputs of one slice decode
    that only meet in the codebook entries, so endpoint packing (5-bit colour, intensity table) and the selector bit planes
    (etc.rs:363-393) are checked against the ETC1 specification rather than against ourselves."""
    nbx, nby, ns, ncb, hist, raw = shape
    eo = ec.bind(oracle)
    _, _, _, _, enc = ec.make_case(eo, nbx, nby, ns, ncb, hist=hist, raw=raw, seed=9)
    e, h = ec.oracle_open(eo, enc, ncb, ncb)
    assert e == 0
    for k in range(ns):
        d = ec.slice_bytes(enc, k)
        e1, etc1 = ec.oracle_etc1(eo, h, nbx, nby, d)
        e2, rgba = ec.oracle_rgba(eo, h, nbx, nby, d)
        assert e1 == 0 and e2 == 0
        wc.check_etc1_matches_rgba(spec, etc1, rgba, nbx, nby)
    eo.orc_etc1s_close(h)


@pytest.mark.parametrize("which", ["build", "build_runs"])
def test_hand_assembled_etc1s_stream_decodes_to_the_hand_derived_result(spec, oracle, which):
    """tests/etc1s_handmade.py builds ETC1S bodies bit by bit from the published format description (no encoder of ours involved)
    with the expected indices and texels worked out in its comments: pins the ETC1S oracle on something we did not generate.
    build(): colour codebook, raw selectors, predictors, delta wrap, history insert / hit.  build_runs(): grayscale codebook,
    Huffman-coded selectors, predictor repeat + VLC, selector runs (plain count and escape + VLC), a swapping history hit."""
    import etc1s_handmade as hm
    case = getattr(hm, which)()
    eo = ec.bind(oracle)
    enc = dict(endpoints=case["endpoints"], selectors=case["selectors"], tables=case["tables"])
    e, h = ec.oracle_open(eo, enc, case["n"], case["n"])
    assert e == 0
    # codebooks as decoded by the oracle == the hand-written values
    cb_e = np.zeros((case["n"], 4), dtype=np.uint8)
    cb_s = np.zeros((case["n"], 8), dtype=np.uint8)
    eo.orc_etc1s_codebooks(h, cb_e.ctypes.data, cb_s.ctypes.data)
    assert [tuple(int(v) for v in r) for r in cb_e] == case["ep_cb"]
    assert [list(int(v) for v in r[:4]) for r in cb_s] == case["sel_cb"]
    nbx, nby = case["nbx"], case["nby"]
    ep = np.zeros(nbx * nby, dtype=np.uint16)
    sel = np.zeros(nbx * nby, dtype=np.uint16)
    assert eo.orc_etc1s_decode_indices(h, nbx, nby, case["slice"], len(case["slice"]), ep.ctypes.data, sel.ctypes.data) == 0
    assert (ep.reshape(nby, nbx) == case["expect_ep"]).all() and (sel.reshape(nby, nbx) == case["expect_sel"]).all()
    want = hm.expected_rgba(case)
    e, rgba = ec.oracle_rgba(eo, h, nbx, nby, case["slice"])
    assert e == 0 and rgba == want.tobytes()
    e, etc1 = ec.oracle_etc1(eo, h, nbx, nby, case["slice"])
    assert e == 0
    wc.check_etc1_matches_rgba(spec, etc1, want.tobytes(), nbx, nby)
    eo.orc_etc1s_close(h)
