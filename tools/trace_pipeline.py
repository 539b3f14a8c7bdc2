#!/usr/bin/env python3
"""Tuning aid: runs one launch of the pipelined kernel from a -DB2BU_TRACE build and prints the per-CTA
timeline of the tile hand-offs (cycles relative to kernel start).  usage: trace_pipeline.py <target> [lib]"""
import ctypes, sys, os, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["B2BU_LIBRARY"] = str(ROOT / "basisu_rs_b200" / (sys.argv[2] if len(sys.argv) > 2 else "libb2bu_trace.so"))
import numpy as np, torch
import basisu_rs_b200 as b
from bench import make_payload, TARGET_NAMES, OUT_BYTES
L = b.lib(); assert L.b2bu_init(0) == 0
t = TARGET_NAMES[sys.argv[1]]; n = 2048 * 2048
blk = make_payload("kat-shuffled", n)
if len(sys.argv) > 3:      # one UASTC mode only: the 32 golden blocks of that mode, tiled and shuffled (instruction-footprint experiment)
    from conftest import Kat
    k = Kat(); sel = k.inputs[k.modes == int(sys.argv[3])]
    blk = np.ascontiguousarray(sel[np.random.default_rng(0).integers(0, len(sel), n)])
# a ring of buffer sets larger than the L2, like bench.py: the traced launch reads its input from HBM
ring = [(torch.from_numpy(np.roll(blk, 7 * i, axis=0).reshape(-1).copy()).cuda(), torch.empty(n * OUT_BYTES[t], dtype=torch.uint8, device="cuda")) for i in range(4)]
st = torch.zeros(1, dtype=torch.int64, device="cuda"); L.b2bu_status_reset_dev(st.data_ptr(), None)
for i in range(7):
    di, do = ring[i % 4]
    L.b2bu_uastc_transcode_dev(t, di.data_ptr(), n * 16, 2048, do.data_ptr(), n * OUT_BYTES[t], st.data_ptr(), None)
torch.cuda.synchronize()
L.b2bu_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.b2bu_debug_trace(None, 1)
di, do = ring[3]
L.b2bu_uastc_transcode_dev(t, di.data_ptr(), n * 16, 2048, do.data_ptr(), n * OUT_BYTES[t], st.data_ptr(), None)
torch.cuda.synchronize()
tr = np.zeros((160, 64), dtype=np.uint64); L.b2bu_debug_trace(tr.ctypes.data, 0)
ends = (tr[:148, 59].astype(np.int64) - tr[:148, 60].astype(np.int64))
print("DMA end per CTA (cycles after start): min %d median %d max %d; start skew %d" % (ends.min(), np.median(ends), ends.max(), int(tr[:148, 60].max() - tr[:148, 60].min())))
g0, g1 = tr[:148, 57].astype(np.int64), tr[:148, 58].astype(np.int64)
print("globaltimer (ns): CTA start spread %d, end spread %d, first start -> last end %d; per-CTA duration min %d median %d max %d" % (
    g0.max() - g0.min(), g1.max() - g1.min(), g1.max() - g0.min(), (g1 - g0).min(), np.median(g1 - g0), (g1 - g0).max()))
order = np.argsort(ends)
print("fastest CTAs:", [(int(c), int(ends[c])) for c in order[:6]])
print("slowest CTAs:", [(int(c), int(ends[c])) for c in order[-12:]])
print("waits of worker 0 (sorter, load) of the slowest:", [(int(tr[c, 61]), int(tr[c, 62])) for c in order[-6:]], "of the fastest:", [(int(tr[c, 61]), int(tr[c, 62])) for c in order[:6]])
for cta in (0, 73, int(order[-1])):
    r = tr[cta].astype(np.int64); t0 = r[60]
    print(f"CTA {cta}: worker0 end {r[63]-t0}, dma end {r[59]-t0}; worker 0 waited {r[61]} cycles for the sorter, {r[62]} for the load")
    for k in range(6):
        v = r[k*6:k*6+5]
        if v[0] == 0: break
        print(f"  tile {k}: load issued@{v[0]-t0} classified@{v[1]-t0} sorted@{v[2]-t0} w0start@{v[3]-t0} w0done@{v[4]-t0}")
