#!/usr/bin/env python3
"""Tuning aid: runs the ETC1S entropy kernel (K2) from a -DB2BU_K2_TRACE build on a config-4 shaped slice and prints
the per-slice stage counters (cycles per block in each warp, wait shares, slow-path rate).
usage: trace_k2.py [lib] [blocks_x] [blocks_y] [slices] [flat share | -1 = the benchmark's config-4 set]"""
import ctypes, os, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["B2BU_LIBRARY"] = str(ROOT / "basisu_rs_b200" / (sys.argv[1] if len(sys.argv) > 1 else "libb2bu_k2trace.so"))
import numpy as np
import basisu_rs_b200 as b
import etc1s_common as ec
from etc1s_synth import encode, make_codebooks, make_indices
from bench import load_oracle
nbx = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nby = int(sys.argv[3]) if len(sys.argv) > 3 else 256
slices = int(sys.argv[4]) if len(sys.argv) > 4 else 8
n_cb = 4096
L = b.lib(); assert L.b2bu_init(0) == 0
orc = ec.bind(load_oracle())
ep_cb, sel_cb = make_codebooks(n_cb, n_cb, seed=3)
flat = float(sys.argv[5]) if len(sys.argv) > 5 else None      # none: the synthesiser's default; < 0: the benchmark's config 4
if flat is not None and flat < 0:          # `slices` different images encoded together, as bench.py does
    eis, sis = [], []
    for k in range(slices):
        e1, s1 = make_indices(nbx, nby, 1, n_cb, n_cb, seed=4 + k, flat=0.15 + 0.45 * ((k * 7) % slices) / max(1, slices - 1))
        eis.append(e1[0]); sis.append(s1[0])
    enc = encode(orc, ep_cb, sel_cb, np.stack(eis), np.stack(sis), nbx, nby, 64, False, False)
    parts, ofs_l, lens_l, pos = [], [], [], 0
    for k in range(slices):
        one = ec.slice_bytes(enc, k); pad = (-len(one)) % 16
        parts.append(one + b"\0" * pad); ofs_l.append(pos); lens_l.append(len(one)); pos += len(one) + pad
    data = b"".join(parts)
    ofs = (ctypes.c_uint64 * slices)(*ofs_l); lens = (ctypes.c_uint64 * slices)(*lens_l)
else:
    ei, si = make_indices(nbx, nby, 1, n_cb, n_cb, seed=4) if flat is None else make_indices(nbx, nby, 1, n_cb, n_cb, seed=4, flat=flat)
    enc = encode(orc, ep_cb, sel_cb, ei, si, nbx, nby, 64, False, False)
    one = ec.slice_bytes(enc, 0); pad = (-len(one)) % 16
    data = (one + b"\0" * pad) * slices
    ofs = (ctypes.c_uint64 * slices)(*[i * (len(one) + pad) for i in range(slices)])
    lens = (ctypes.c_uint64 * slices)(*[len(one)] * slices)
dec = b.Etc1sDecoder(n_cb, n_cb, enc["endpoints"], enc["selectors"], enc["tables"])
bits = (ctypes.c_uint32 * 4)(); mx = (ctypes.c_uint32 * 4)()
L.b2bu_etc1s_table_info.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
L.b2bu_etc1s_table_info(dec._h, bits, mx)
print("first-level bits", list(bits), "max code length", list(mx))
out = np.zeros(nbx * nby * slices * 8, dtype=np.uint8)
L.b2bu_debug_k2_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
for rep in range(2):
    L.b2bu_debug_k2_trace(None, 1)
    st = L.b2bu_etc1s_transcode_slices(dec._h, 3, nbx, nby, data, len(data), ofs, lens, slices, out.ctypes.data, out.size)
    assert st == 0, st
k2 = ctypes.c_float(); L.b2bu_etc1s_last_timing(dec._h, ctypes.byref(k2), None, None, None)
tr = np.zeros((64, 16), dtype=np.uint64); L.b2bu_debug_k2_trace(tr.ctypes.data, 0)
tr2 = np.zeros((64, 2), dtype=np.uint64); L.b2bu_debug_k2_trace(tr2.ctypes.data, 2)
nblk = nbx * nby
print("K2 %.2f ms for %d slices of %d blocks (%.1f bits/block)" % (k2.value, slices, nblk, 8.0 * sum(lens) / slices / nblk))
print("slice: bits/block -> tokenizer cyc/block (lean share, slow steps): " + "  ".join("%d: %.1f -> %.0f (%.3f, %d)" % (k, 8.0 * lens[k] / nblk, tr[k, 0] / nblk, tr[k, 15] * 8 / nblk, tr[k, 3]) for k in range(0, slices, max(1, slices // 16))))
print("slice: SM -> cyc/block: " + "  ".join("%d: %d -> %.0f" % (k, int(tr2[k, 0]), tr[k, 0] / nblk) for k in range(slices)))
order = np.argsort(tr[:slices, 0]); print("tokenizer cycles/block over the slices: min %.0f median %.0f max %.0f (slice %d)" % (tr[order[0], 0] / nblk, tr[order[slices // 2], 0] / nblk, tr[order[-1], 0] / nblk, order[-1]))
for s in (int(order[0]), int(order[-1])):
    r = tr[s].astype(np.float64)
    print("slice %d: tokenizer %.0f cyc/block (waiting %.0f%%), %.2f symbols/block, slow path %.2f%% of symbols; "
          "resolver %.0f cyc/block (waiting %.0f%%), history hits %.2f/block, serial endpoint rounds %d"
          % (s, r[0] / nblk, 100 * r[1] / max(r[0], 1), r[2] / nblk, 100 * r[3] / max(r[2], 1), r[4] / nblk, 100 * r[5] / max(r[4], 1), r[6] / nblk, int(r[7])))
    print("   lean steps: %d of %d" % (int(r[15]), nblk // 8))
    print("   resolver cycles per round: tokens/checks %.0f, endpoint scan %.0f, selectors %.0f, checks/stores %.0f" % tuple(r[8 + k] * 32 / nblk for k in range(4)))
