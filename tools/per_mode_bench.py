#!/usr/bin/env python3
"""Tuning aid: device time of the tile kernel per UASTC mode (payload = the KAT blocks of one mode, tiled), so that the
instruction work can be attributed.  usage: per_mode_bench.py [target]"""
import ctypes, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import basisu_rs_b200 as b
from conftest import Kat, OUT_BYTES, TARGETS
t = TARGETS[sys.argv[1] if len(sys.argv) > 1 else "astc"]
L = b.lib(); assert L.b2bu_init(0) == 0
kat = Kat(); n = 2048 * 2048
status = torch.zeros(1, dtype=torch.int64, device="cuda"); L.b2bu_status_reset_dev(status.data_ptr(), None)
rng = np.random.default_rng(0)
res = {}
for mode in list(range(19)) + [-1]:
    src = kat.inputs if mode < 0 else kat.inputs[kat.modes == mode]
    blk = src[rng.integers(0, len(src), size=n)]
    bufs = [(torch.from_numpy(blk.reshape(-1).copy()).cuda(), torch.empty(n * OUT_BYTES[t], dtype=torch.uint8, device="cuda")) for _ in range(3)]
    def run(i):
        di, do = bufs[i % 3]
        assert L.b2bu_uastc_transcode_dev(t, di.data_ptr(), n * 16, 2048, do.data_ptr(), n * OUT_BYTES[t], status.data_ptr(), None) == 0
    for i in range(3): run(i)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(30): run(i)
    e.record(); torch.cuda.synchronize()
    res[mode] = a.elapsed_time(e) / 30 * 1e3
print(" ".join("m%d=%.0f" % (m, us) if m >= 0 else "all=%.0f" % us for m, us in res.items()))
print("mean of single-mode runs %.1f us" % (sum(v for m, v in res.items() if m >= 0) / 19))
