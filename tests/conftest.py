import ctypes
import pathlib
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

TARGETS = {"rgba": 0, "astc": 1, "bc7": 2, "etc1": 3, "etc2": 4}
OUT_BYTES = {0: 64, 1: 16, 2: 16, 3: 8, 4: 16}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _newer(target: pathlib.Path, deps):
    return (not target.exists()) or any(d.stat().st_mtime > target.stat().st_mtime for d in deps if d.exists())


def build_oracle() -> pathlib.Path:
    so = ROOT / "oracle" / "libbasisu_oracle.so"
    deps = [p for p in (ROOT / "oracle").iterdir() if p.suffix in (".c", ".inc", ".h")]
    if _newer(so, deps):
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, stdout=subprocess.DEVNULL)
    return so


def build_emu() -> pathlib.Path:
    so = ROOT / "tests" / "emu" / "libemu_uastc.so"
    csrc = ROOT / "basisu_rs_b200" / "csrc"
    deps = [ROOT / "tests" / "emu" / "emu_uastc.cpp", csrc / "uastc_device.cuh", csrc / "device_tables.h", csrc / "device_tables_gen.inc"]
    if _newer(so, deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fvisibility=hidden",
                        "-Wno-attributes", "-o", str(so), str(deps[0])], check=True)
    return so


def build_spec() -> pathlib.Path:
    """tests/spec/spec_decoders.c: decoders of the TARGET formats written from the public specifications (independent witnesses)"""
    src = ROOT / "tests" / "spec" / "spec_decoders.c"
    so = ROOT / "tests" / "spec" / "libspec_decoders.so"
    if _newer(so, [src]):
        subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-o", str(so), str(src)], check=True)
    return so


class Spec:
    """numpy front-end of the spec decoders: arrays of blocks in, arrays of texels out"""

    def __init__(self):
        c = ctypes
        self.L = L = c.CDLL(str(build_spec()))
        for name in ("spec_astc_decode_4x4", "spec_etc1_decode", "spec_bc7_decode_mode56"):
            getattr(L, name).argtypes = [c.c_void_p, c.c_void_p]
            getattr(L, name).restype = c.c_int
        L.spec_eac_alpha_decode.argtypes = [c.c_void_p, c.c_void_p]
        L.spec_eac_alpha_decode.restype = None
        L.spec_bc7_mode.argtypes = [c.c_void_p]
        L.spec_astc_partition_map.argtypes = [c.c_int, c.c_int, c.c_void_p]
        L.spec_astc_partition_map.restype = None

    def _run(self, fn, blocks, in_bytes, out_bytes):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, in_bytes)
        out = np.zeros((blocks.shape[0], out_bytes), dtype=np.uint8)
        rc = np.zeros(blocks.shape[0], dtype=np.int32)
        pi, po = blocks.ctypes.data, out.ctypes.data
        for i in range(blocks.shape[0]):
            r = fn(pi + i * in_bytes, po + i * out_bytes)
            rc[i] = 0 if r is None else r
        return rc, out

    def astc(self, blocks):
        return self._run(self.L.spec_astc_decode_4x4, blocks, 16, 64)

    def etc1(self, blocks):
        return self._run(self.L.spec_etc1_decode, blocks, 8, 64)

    def bc7_56(self, blocks):
        return self._run(self.L.spec_bc7_decode_mode56, blocks, 16, 64)

    def eac_alpha(self, blocks):
        return self._run(self.L.spec_eac_alpha_decode, blocks, 8, 16)[1]

    def bc7_modes(self, blocks):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 16)
        low = blocks[:, 0].astype(np.int32)
        return np.array([(int(v) & -int(v)).bit_length() - 1 if v else 8 for v in low])


@pytest.fixture(scope="session")
def spec():
    return Spec()


@pytest.fixture(scope="session")
def oracle():
    L = ctypes.CDLL(str(build_oracle()))
    c = ctypes
    L.orc_uastc_transcode_slice.argtypes = [c.c_int, c.c_void_p, c.c_size_t, c.c_size_t, c.c_void_p, c.c_int, c.c_void_p]
    L.orc_crc16.restype = c.c_uint16
    L.orc_crc16.argtypes = [c.c_void_p, c.c_size_t, c.c_uint16]
    L.orc_bitreader_read_at.restype = c.c_uint32
    L.orc_bitreader_read_at.argtypes = [c.c_void_p, c.c_size_t, c.c_size_t, c.c_uint]
    L.orc_bitwriter_lsb.argtypes = [c.c_void_p, c.c_size_t, c.c_size_t, c.c_uint, c.c_uint32]
    L.orc_bitwriter_msb.argtypes = [c.c_void_p, c.c_size_t, c.c_size_t, c.c_uint, c.c_uint32, c.c_int]
    L.orc_unquant_endpoint.restype = c.c_uint8
    L.orc_unquant_endpoint.argtypes = [c.c_uint, c.c_uint, c.c_uint]
    return L


@pytest.fixture(scope="session")
def emu():
    L = ctypes.CDLL(str(build_emu()))
    L.emu_uastc_transcode.restype = ctypes.c_uint64
    L.emu_uastc_transcode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
    return L


class Kat:
    """tests/golden/uastc_kat.bin: the reference's 608 x 5 known-answer vectors."""

    def __init__(self):
        blob = (ROOT / "tests" / "golden" / "uastc_kat.bin").read_bytes()
        assert blob[:4] == b"UKAT"
        self.n = struct.unpack_from("<I", blob, 4)[0]
        recs = np.frombuffer(blob, dtype=np.uint8, offset=8).reshape(self.n, 137)
        self.modes = recs[:, 0].copy()
        self.inputs = np.ascontiguousarray(recs[:, 1:17])
        self.expected = {0: np.ascontiguousarray(recs[:, 17:81]), 1: np.ascontiguousarray(recs[:, 81:97]),
                         2: np.ascontiguousarray(recs[:, 97:113]), 3: np.ascontiguousarray(recs[:, 113:121]),
                         4: np.ascontiguousarray(recs[:, 121:137])}


@pytest.fixture(scope="session")
def kat():
    return Kat()


def oracle_transcode(L, target, blocks, blocks_per_row=1, threads=8):
    """Runs the oracle on an (n,16) uint8 array.  Returns (err, first_bad, out bytes as 1-D array)."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    n = blocks.shape[0]
    out = np.zeros(n * OUT_BYTES[target], dtype=np.uint8)
    bad = ctypes.c_size_t(0)
    e = L.orc_uastc_transcode_slice(target, blocks.ctypes.data, n * 16, blocks_per_row, out.ctypes.data, threads, ctypes.byref(bad))
    return e, bad.value, out


def rgba_image_to_blocks(img, n, bpr):
    """row-major RGBA image bytes -> (n,64) per-block bytes (raster order inside the block)."""
    return img.reshape(n // bpr, 4, bpr, 16).transpose(0, 2, 1, 3).reshape(n, 64)


def cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_lib():
    if not cuda_available():
        pytest.skip("no CUDA device")
    import basisu_rs_b200 as b
    L = b.lib()           # raises loudly if the extension is missing
    st = L.b2bu_init(0)
    assert st == 0, L.b2bu_last_cuda_error()
    return b
