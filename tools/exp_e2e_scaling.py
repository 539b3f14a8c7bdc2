#!/usr/bin/env python3
"""e2e time of b2bu_uastc_transcode (pinned host buffers) against the input size: T(n) = a + b n separates the fixed cost
of a call (ramp, last kernel + copy, synchronisation) from the per-byte rate.  usage: tools/exp_e2e_scaling.py [lib.so]"""
import ctypes, os, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
if len(sys.argv) > 1: os.environ["B2BU_LIBRARY"] = str(ROOT / "basisu_rs_b200" / sys.argv[1])
import numpy as np, torch
import basisu_rs_b200 as b
from bench import make_payload
L = b.lib(); assert L.b2bu_init(0) == 0
res = []
for mib in (8, 16, 32, 64, 128, 256):
    n = mib * 65536
    blk = make_payload("kat-shuffled", n)
    h_in = torch.from_numpy(blk.reshape(-1)).pin_memory(); h_out = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
    fb = ctypes.c_uint64(0)
    for _ in range(3): assert L.b2bu_uastc_transcode(1, h_in.data_ptr(), n * 16, h_out.data_ptr(), n * 16, ctypes.byref(fb)) == 0
    reps = max(5, 2048 // mib // 4)
    t0 = time.perf_counter()
    for _ in range(reps): L.b2bu_uastc_transcode(1, h_in.data_ptr(), n * 16, h_out.data_ptr(), n * 16, ctypes.byref(fb))
    dt = (time.perf_counter() - t0) / reps
    res.append((mib, dt))
    print("%4d MiB: %8.1f us  %6.2f GB/s per direction" % (mib, dt * 1e6, n * 16 / dt / 1e9))
x = np.array([m * 1048576.0 for m, _ in res]); y = np.array([t for _, t in res])
bb, aa = np.polyfit(x, y, 1)
print("fit: fixed %.1f us per call, %.2f GB/s per direction asymptotically" % (aa * 1e6, 1 / bb / 1e9))
