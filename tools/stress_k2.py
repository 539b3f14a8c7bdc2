#!/usr/bin/env python3
"""Stress / determinism check of the ETC1S pipeline: decodes a batch of different slices repeatedly (one slice per SM and
packed several per CTA) and compares every run with the oracle's indices-derived output of the first.  usage: stress_k2.py [reps]"""
import ctypes, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import basisu_rs_b200 as b
import etc1s_common as ec
from bench import load_oracle
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L = b.lib(); assert L.b2bu_init(0) == 0
orc = ec.bind(load_oracle())
bad = 0
for nbx, nby, ns, ncb in ((96, 64, 40, 1500), (257, 33, 300, 700), (64, 64, 1300, 256)):
    _, _, _, _, enc = ec.make_case(orc, nbx, nby, min(ns, 12), ncb, seed=nbx + ns)
    uniq = len(enc["slice_ofs"])
    ofs = (ctypes.c_uint64 * ns)(*[enc["slice_ofs"][k % uniq] for k in range(ns)])
    ln = (ctypes.c_uint64 * ns)(*[enc["slice_len"][k % uniq] for k in range(ns)])
    e, h = ec.oracle_open(orc, enc, ncb, ncb)
    want = [ec.oracle_etc1(orc, h, nbx, nby, ec.slice_bytes(enc, k))[1] for k in range(uniq)]
    dec = b.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"])
    out = np.zeros(ns * nbx * nby * 8, dtype=np.uint8)
    per = nbx * nby * 8
    for r in range(reps):
        out[:] = 0
        st = L.b2bu_etc1s_transcode_slices(dec._h, 3, nbx, nby, enc["slice_data"], len(enc["slice_data"]), ofs, ln, ns, out.ctypes.data, out.size)
        assert st == 0, st
        for k in range(ns):
            if out[k * per:(k + 1) * per].tobytes() != want[k % uniq]:
                bad += 1
    dec.close(); orc.orc_etc1s_close(h)
    print("%dx%d blocks x %d slices x %d runs: mismatches so far %d" % (nbx, nby, ns, reps, bad))
sys.exit(1 if bad else 0)
