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
di = torch.from_numpy(blk.reshape(-1)).cuda(); do = torch.empty(n * OUT_BYTES[t], dtype=torch.uint8, device="cuda")
st = torch.zeros(1, dtype=torch.int64, device="cuda"); L.b2bu_status_reset_dev(st.data_ptr(), None)
for i in range(3):
    L.b2bu_uastc_transcode_dev(t, di.data_ptr(), n * 16, 2048, do.data_ptr(), n * OUT_BYTES[t], st.data_ptr(), None)
torch.cuda.synchronize()
L.b2bu_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.b2bu_debug_trace(None, 1)
L.b2bu_uastc_transcode_dev(t, di.data_ptr(), n * 16, 2048, do.data_ptr(), n * OUT_BYTES[t], st.data_ptr(), None)
torch.cuda.synchronize()
tr = np.zeros((160, 64), dtype=np.uint64); L.b2bu_debug_trace(tr.ctypes.data, 0)
for cta in (0, 73, 147):
    r = tr[cta].astype(np.int64); t0 = r[60]
    print(f"CTA {cta}: worker0 end {r[63]-t0}, dma end {r[59]-t0}, wait(sorted) first worker {r[61]}, last worker {r[62]}")
    print("  sorter phases tile 4 (warp0 / last warp), rel. to full:", [int(x - r[4*6+1]) for x in r[40:48]], [int(x - r[4*6+1]) for x in r[50:58]])
    for k in range(6):
        v = r[k*6:k*6+5]
        if v[0] == 0: break
        print(f"  tile {k}: load@{v[0]-t0} full@{v[1]-t0} sorted@{v[2]-t0} w0start@{v[3]-t0} w0done@{v[4]-t0}")
