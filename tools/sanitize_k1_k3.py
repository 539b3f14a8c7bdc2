#!/usr/bin/env python3
"""Driver for compute-sanitizer runs over K1 (the tile pipeline, all five targets) and K3 (the ETC1S gathers):
one modest launch of each through the C ABI, checked against the oracle; K2 (ETC1S entropy decode: one slice per call = the
wide pipelines with their helper warps) runs on the way to K3.  usage (under gpurun):
  compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_k1_k3.py [k2]      (k2: skip K1)"""
import ctypes, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import basisu_rs_b200 as b
import conftest, etc1s_common as ec
from uastc_synth import random_blocks
L = b.lib(); assert L.b2bu_init(0) == 0
orc = ctypes.CDLL(str(conftest.build_oracle()))
orc.orc_uastc_transcode_slice.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
n, bpr = 148 * 1100 + 37 * 128, 128                       # every SM gets a short first tile and a ragged second one
n -= n % bpr
blk = random_blocks(n, seed=8)
d_in = torch.from_numpy(blk.reshape(-1)).cuda()
status = torch.zeros(1, dtype=torch.int64, device="cuda")
sh = torch.cuda.current_stream().cuda_stream
L.b2bu_status_reset_dev(status.data_ptr(), sh)
for t in (range(5) if "k2" not in sys.argv[1:] else ()):
    ob = conftest.OUT_BYTES[t]
    d_out = torch.zeros(n * ob, dtype=torch.uint8, device="cuda")
    assert L.b2bu_uastc_transcode_dev(t, d_in.data_ptr(), n * 16, bpr, d_out.data_ptr(), n * ob, status.data_ptr(), sh) == 0
    torch.cuda.synchronize()
    want = np.zeros(n * ob, dtype=np.uint8)
    orc.orc_uastc_transcode_slice(t, blk.ctypes.data, n * 16, bpr, want.ctypes.data, 8, None)
    assert (d_out.cpu().numpy() == want).all(), t
    print("K1 target", t, "ok,", n, "blocks")
eo = ec.bind(orc)
nbx, nby, ncb = 96, 64, 1024
_, _, _, _, enc = ec.make_case(eo, nbx, nby, 3, ncb, seed=2)
e, h = ec.oracle_open(eo, enc, ncb, ncb)
dec = b.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"])
for k in range(3):
    d = ec.slice_bytes(enc, k)
    assert dec.transcode_to_etc1(nbx, nby, d) == ec.oracle_etc1(eo, h, nbx, nby, d)[1]
    assert dec.transcode_to_bc1(nbx, nby, d) == ec.oracle_bc1(eo, h, nbx, nby, d)[1]
    assert dec.decode_to_rgba(nbx, nby, d) == ec.oracle_rgba(eo, h, nbx, nby, d)[1]
assert dec.decode_to_rgba(nbx, nby, ec.slice_bytes(enc, 0), ec.slice_bytes(enc, 1)) == ec.oracle_rgba(eo, h, nbx, nby, ec.slice_bytes(enc, 0), ec.slice_bytes(enc, 1))[1]
dec.close()
print("K2 + K3 (etc1, bc1, rgba, rgba + alpha) ok")
# a ragged width (every row ends in a partial round: the slow path), a small history buffer, and a truncated stream
nbx, nby, ncb = 75, 21, 300
_, _, _, _, enc = ec.make_case(eo, nbx, nby, 2, ncb, seed=5, hist=16)
e, h = ec.oracle_open(eo, enc, ncb, ncb)
dec = b.Etc1sDecoder(ncb, ncb, enc["endpoints"], enc["selectors"], enc["tables"])
d = ec.slice_bytes(enc, 0)
assert dec.transcode_to_etc1(nbx, nby, d) == ec.oracle_etc1(eo, h, nbx, nby, d)[1]
try:
    dec.transcode_to_etc1(nbx, nby, d[: len(d) // 3])
    cut = "decoded"
except Exception as ex:
    cut = "refused: %s" % ex
dec.close()
print("K2 ragged width ok; truncated stream", cut)
