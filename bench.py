#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native Basis Universal transcoder.

Contract (one JSON line on stdout from rank 0):
  metric   Gtexels/s of the UASTC -> ASTC 4x4 transcode (BASELINE.json configs[1]: synthetic
           8192x8192 texture = 4,194,304 blocks per step per GPU)
  value    device-resident throughput (inputs already in HBM, CUDA events on the launching stream,
           max over ranks), whole job over all N GPUs (weak scaling: every GPU transcodes its own texture)
  e2e      the same metric through the reference-facing C-ABI call b2bu_uastc_transcode() with
           pinned HOST buffers: H2D + kernels + D2H inside the timed region
  roofline achieved algorithmic GB/s (32 B per block) of the K1<ASTC> kernel vs
           the measured HBM copy peak
  cpu_baseline  the CPU oracle port (C restatement of the reference's scalar path) on the host cores
--impl reference times that CPU port on all host threads instead (the reference is Rust and cannot
be built in this image).
"""
import argparse
import ctypes
import json
import os
import pathlib
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np

TARGET_NAMES = {"rgba": 0, "astc": 1, "bc7": 2, "etc1": 3, "etc2": 4}
OUT_BYTES = {0: 64, 1: 16, 2: 16, 3: 8, 4: 16}
ALGO_BYTES = {0: 80, 1: 32, 2: 32, 3: 24, 4: 32}          # input + output per block
INT_OPS = {0: 400, 1: 300, 2: 550, 3: 800, 4: 1300}        # algorithmic integer ops per block



# The contract is ONE JSON line on stdout.  Libraries under us write there too (NCCL prints its version banner to
# stdout), so everything else is routed to stderr: fd 1 is pointed at fd 2 for the whole run and the line goes to the
# saved descriptor.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def bind_to_gpu_numa_node(device_index):
    """One process per GPU: run this rank (and first-touch its pinned host buffers) on the CPU socket the GPU hangs off, so
    that its H2D / D2H copies do not cross the socket interconnect.  Returns (node, why): node is None when no binding was
    made, and `why` says what was found (a container that exposes a single NUMA node reports -1 for every device)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                       # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        path = "/sys/bus/pci/devices/%s/numa_node" % bus
        if not os.path.exists(path):
            return None, "no %s (PCI device not visible in this container)" % path
        node = int(open(path).read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] if os.path.isdir("/sys/devices/system/node") else []
        if node < 0:
            return None, "sysfs reports numa_node = -1 for the GPU (%d NUMA node(s) visible): nothing to bind to" % len(nodes)
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node, "bound to the %d allowed CPUs of node %d (%d node(s) visible)" % (len(cpus), node, len(nodes))
        return None, "node %d has no CPU this process may run on" % node
    except Exception as e:                                    # noqa: BLE001 -- reported in the JSON line
        return None, "lookup failed: %r" % (e,)


def pcie_ceiling(torch, dist, world, seconds=0.4, nbytes=64 << 20):
    """What the host side can feed: every rank copies pinned host -> device and device -> pinned host CONCURRENTLY (two streams,
    64 MiB pieces) for `seconds`, all ranks at the same time; returns this rank's GB/s per direction and the sum over ranks.
    The end-to-end number cannot exceed min(h2d, d2h) / 16 B per block (x 16 texels)."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def burst(reps):
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    burst(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(2)]
    reps = 0
    e[0][0].record(s1)
    e[1][0].record(s2)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        burst(4)
        reps += 4
        s1.synchronize()
    e[0][1].record(s1)
    e[1][1].record(s2)
    torch.cuda.synchronize()
    h2d = nbytes * reps / (e[0][0].elapsed_time(e[0][1]) * 1e-3) / 1e9
    d2h = nbytes * reps / (e[1][0].elapsed_time(e[1][1]) * 1e-3) / 1e9
    t = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    agg = [float(x) for x in t.tolist()]
    return {"h2d_gbs_this_rank": h2d, "d2h_gbs_this_rank": d2h, "h2d_gbs_all_ranks": agg[0], "d2h_gbs_all_ranks": agg[1],
            "how": "pinned 64 MiB pieces, H2D and D2H concurrently on two streams, all %d rank(s) at once, %.1f s" % (world, seconds)}


def emit_line(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())

def make_payload(kind: str, nblocks: int, seed: int = 0) -> np.ndarray:
    """Synthetic UASTC payloads of SURVEY.md section 8d, built from the reference's 608 golden input
    blocks (tests/golden/uastc_kat.bin): kat-coherent = tiled in file order (runs of 32 same-mode
    blocks), kat-shuffled = same multiset in a seeded permutation, random = valid-random blocks."""
    if kind == "random":
        from uastc_synth import random_blocks
        return random_blocks(nblocks, seed=seed + 1)
    blob = (ROOT / "tests" / "golden" / "uastc_kat.bin").read_bytes()
    recs = np.frombuffer(blob, dtype=np.uint8, offset=8).reshape(-1, 137)
    inputs = np.ascontiguousarray(recs[:, 1:17])
    reps = (nblocks + len(inputs) - 1) // len(inputs)
    tiled = np.tile(inputs, (reps, 1))[:nblocks]
    if kind == "kat-coherent":
        return np.ascontiguousarray(tiled)
    if kind == "kat-shuffled":
        rng = np.random.default_rng(seed)
        return np.ascontiguousarray(tiled[rng.permutation(nblocks)])
    raise ValueError(kind)


def load_oracle():
    import conftest
    L = ctypes.CDLL(str(conftest.build_oracle()))
    L.orc_uastc_transcode_slice.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_void_p]
    return L


def time_oracle(L, target, blocks, bpr, threads, min_seconds=2.0, max_reps=50):
    """Gtexel/s of the CPU port on `blocks` with `threads` threads."""
    n = blocks.shape[0]
    out = np.zeros(n * OUT_BYTES[target], dtype=np.uint8)
    L.orc_uastc_transcode_slice(target, blocks.ctypes.data, n * 16, bpr, out.ctypes.data, threads, None)   # warm
    reps, t0 = 0, time.perf_counter()
    while True:
        e = L.orc_uastc_transcode_slice(target, blocks.ctypes.data, n * 16, bpr, out.ctypes.data, threads, None)
        assert e == 0
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or reps >= max_reps:
            break
    return reps * n * 16 / dt / 1e9, dt, reps, out


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.02)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(target_name):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(target_name)
        except Exception:
            return None
    return None


def workload_config(args, n, ob):
    """the `config` object of the JSON line: ONE definition for both arms (the driver compares them key by key)"""
    return {"workload": "UASTC->%s 4x4 transcode, synthetic 8192x8192 texture (%d blocks) per GPU per step" % (args.target.upper(), n),
            "payload": args.payload + " (reference KAT blocks tiled/permuted, seed = rank)",
            "l2_policy": "GPU arm: ring of %d distinct in/out buffer sets (%d MiB) cycled, larger than the 126 MB L2"
                         % (args.ring, args.ring * n * (16 + ob) >> 20),
            "sharding": "one texture per GPU, no collective"}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Rust and this
    image has no rustc/cargo, so the arm runs the C oracle port (oracle/basisu_oracle.c) with every
    host thread.  A step is the WHOLE workload of the GPU arm (all args.blocks blocks, same payload): about 70 ms on 16 cores."""
    if rank != 0:
        return
    target = TARGET_NAMES[args.target]
    cores = os.cpu_count() or 1
    n = args.blocks
    blocks = make_payload(args.payload, n)
    L = load_oracle()
    out = np.zeros(n * OUT_BYTES[target], dtype=np.uint8)
    for _ in range(args.warmup):
        L.orc_uastc_transcode_slice(target, blocks.ctypes.data, n * 16, args.bpr, out.ctypes.data, cores, None)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        assert L.orc_uastc_transcode_slice(target, blocks.ctypes.data, n * 16, args.bpr, out.ctypes.data, cores, None) == 0
    dt = time.perf_counter() - t0
    value = args.steps * n * 16 / dt / 1e9
    sample = f"the full workload: {n} blocks per step ({args.payload}), {cores} threads, static block partition"
    line = {
        "impl": "reference", "metric": "Gtexels/s UASTC->%s" % args.target.upper(), "value": value, "unit": "Gtexel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, n, OUT_BYTES[target]),
        "cpu_baseline": {"value": value, "unit": "Gtexel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gtexel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference (Rust) cannot be built here: no rustc/cargo; this is the C restatement pinned on the reference's 3,040 KATs",
    }
    emit_line(line)



def mip_chain_blocks(size=8192):
    """BASELINE configs[2]: 8192^2 with the full mip chain, level k has ceil(max(1, size >> k) / 4)^2 blocks."""
    out, k = [], 0
    while True:
        d = max(1, size >> k)
        nb = (d + 3) // 4
        out.append(nb)
        if d == 1:
            break
        k += 1
    return out


def bench_c3_bc7_mips(L, b, torch, payload, steps, status, sh):
    """configs[2]: UASTC -> BC7 over an 8192^2 texture with its full mip chain (14 slices).  Device-resident: one call of
    b2bu_uastc_transcode_slices_dev (the levels are contiguous, so the chain is ONE launch; the per-level launch loop is
    timed beside it).  File level: the same chain as a .basis file through b2bu_read_to (one upload, CRC-16 on the GPU, one
    launch, one copy back), with the CRC kernel and the host CRC loop it replaces timed on their own."""
    import ctypes as c
    dims = mip_chain_blocks()
    total = sum(d * d for d in dims)
    blocks = make_payload(payload, total, seed=7)
    # three copies of the 89 MB chain on either side (537 MB cycled: nothing a launch reads or writes is still in the 126 MB L2)
    R = 3
    d_ins = [torch.from_numpy(blocks.reshape(-1)).cuda() for _ in range(R)]
    d_out = [torch.empty(total * 16, dtype=torch.uint8, device="cuda") for _ in range(R)]
    offs = np.cumsum([0] + [d * d for d in dims])
    sl = (b.SliceDev * len(dims))()
    for lv, d in enumerate(dims):
        sl[lv] = b.SliceDev(int(offs[lv]) * 16, int(offs[lv]) * 16, d * d, d, 0)

    def chain_levels(i):
        for lv, d in enumerate(dims):
            n = d * d
            st = L.b2bu_uastc_transcode_dev(2, d_ins[i % R].data_ptr() + int(offs[lv]) * 16, n * 16, d, d_out[i % R].data_ptr() + int(offs[lv]) * 16, n * 16,
                                            status.data_ptr(), sh)
            assert st == 0

    def chain_one(i):
        assert L.b2bu_uastc_transcode_slices_dev(2, d_ins[i % R].data_ptr(), d_out[i % R].data_ptr(), sl, len(dims), status.data_ptr(), sh) == 0

    def timed(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        e.record()
        torch.cuda.synchronize()
        return a.elapsed_time(e) / steps * 1e3
    us_levels = timed(chain_levels)
    before = L.b2bu_launch_count()
    us = timed(chain_one)
    launches = (L.b2bu_launch_count() - before) // (steps + 3)
    orc = load_oracle()
    ns = min(total, 1 << 18)
    want = np.zeros(ns * 16, dtype=np.uint8)
    orc.orc_uastc_transcode_slice(2, blocks.ctypes.data, ns * 16, 1, want.ctypes.data, os.cpu_count() or 1, None)
    # level 0 holds the first blocks of the payload; the tail levels are checked whole
    ok = bool((d_out[(steps - 1) % R][: ns * 16].cpu().numpy() == want).all())
    tail0 = int(offs[3])
    want_t = np.zeros((total - tail0) * 16, dtype=np.uint8)
    orc.orc_uastc_transcode_slice(2, blocks[tail0:].ctypes.data, (total - tail0) * 16, 1, want_t.ctypes.data, os.cpu_count() or 1, None)
    ok = ok and bool((d_out[(steps - 1) % R][tail0 * 16: total * 16].cpu().numpy() == want_t).all())
    res = {"workload": "UASTC->BC7, 8192x8192 + full mip chain (%d levels, %d blocks), contiguous levels merged into one launch" % (len(dims), total),
           "us_per_chain": us, "gtexel_s": total * 16 / us / 1e3, "algo_gb_s": total * 32 / us / 1e3, "launches_per_chain": int(launches),
           "us_per_chain_one_launch_per_level": us_levels, "parity_vs_oracle": ok,
           "l2_policy": "ring of %d chains in and out (%d MB cycled), larger than L2" % (R, 2 * R * total * 16 // 1000000)}
    del d_ins
    # ---- file level: the chain as a .basis file ----
    from basis_writer import build_basis
    slices = [dict(data=blocks[int(offs[lv]):int(offs[lv + 1])].tobytes(), orig_width=max(1, 8192 >> lv), orig_height=max(1, 8192 >> lv),
                   num_blocks_x=d, num_blocks_y=d, level_index=lv, image_index=0) for lv, d in enumerate(dims)]
    f = build_basis(slices, tex_format=1, total_images=1)
    n = len(f)
    hbuf = L.b2bu_host_alloc(n)
    hout = L.b2bu_host_alloc(total * 16)
    c.memmove(hbuf, f, n)
    imgs = (b._CImage * len(dims))()
    cnt, need = c.c_uint32(0), c.c_uint64(0)
    src = c.cast(hbuf, c.POINTER(c.c_uint8))
    dst = c.cast(hout, c.POINTER(c.c_uint8))
    best = None
    for r in range(4):
        t0 = time.perf_counter()
        st = L.b2bu_read_to(2, src, n, None, imgs, len(dims), c.byref(cnt), dst, total * 16, c.byref(need))
        dt = time.perf_counter() - t0
        assert st == 0, st
        best = dt if best is None or (r and dt < best) else best
    got = np.ctypeslib.as_array(dst, shape=(total * 16,))
    ok_file = bool((got[: ns * 16] == want).all() and (got[tail0 * 16:] == want_t).all())
    # CRC-16 of the payload: GPU kernel (device-resident file) vs the reference's single-core loop (host)
    d_file = torch.from_numpy(np.frombuffer(f, dtype=np.uint8).copy()).cuda()
    crc = c.c_uint16(0)
    for r in range(3):
        t0 = time.perf_counter()
        assert L.b2bu_crc16_dev(d_file.data_ptr() + 77, n - 77, 0, c.byref(crc), sh) == 0
        dt_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    host_crc = L.b2bu_crc16(src, n, 0) if False else L.b2bu_crc16(c.cast(hbuf + 77, c.POINTER(c.c_uint8)), n - 77, 0)
    dt_host = time.perf_counter() - t0
    res["file_level"] = {"api": "b2bu_read_to(BC7) on the chain as a .basis file (%d bytes, pinned buffers)" % n, "ms": best * 1e3,
                         "gtexel_s": total * 16 / best / 1e9, "parity_vs_oracle": ok_file,
                         "crc16_gpu_ms_incl_sync": dt_dev * 1e3, "crc16_gpu_gb_s": (n - 77) / dt_dev / 1e9,
                         "crc16_host_single_core_ms": dt_host * 1e3, "crc_match": bool(crc.value == host_crc)}
    L.b2bu_host_free(hbuf)
    L.b2bu_host_free(hout)
    return res


def bench_c4_etc1s(L, b, nb=1024, slices=64, n_cb=4096, reps=2):
    """configs[3]: ETC1S -> ETC1 / RGBA, warp-per-slice entropy decode (K2) + codebook gather (K3), nb x nb blocks x `slices` slices.
    One slice is encoded by the test encoder on a procedural index map and replicated (the decoder keeps no state across slices)."""
    import etc1s_common as ec
    from etc1s_synth import encode, make_codebooks, make_indices
    orc = ec.bind(load_oracle())
    ep_cb, sel_cb = make_codebooks(n_cb, n_cb, seed=3)
    # every slice is a different procedural image: own seed, and a share of flat regions that varies from 15 % to 60 % so that the
    # slices differ in compressed length (K2 decodes one slice per warp pair: unequal slices are its real load-balance problem)
    eis, sis = [], []
    for k in range(slices):
        e1, s1 = make_indices(nb, nb, 1, n_cb, n_cb, seed=4 + k, flat=0.15 + 0.45 * ((k * 7) % slices) / max(1, slices - 1))
        eis.append(e1[0])
        sis.append(s1[0])
    enc = encode(orc, ep_cb, sel_cb, np.stack(eis), np.stack(sis), nb, nb, 64, False, False)
    del eis, sis
    parts, ofs_l, lens_l, pos = [], [], [], 0
    for k in range(slices):
        one = ec.slice_bytes(enc, k)
        pad = (-len(one)) % 16
        parts.append(one + b"\0" * pad)
        ofs_l.append(pos)
        lens_l.append(len(one))
        pos += len(one) + pad
    data = b"".join(parts)
    del parts
    ofs = (ctypes.c_uint64 * slices)(*ofs_l)
    lens = (ctypes.c_uint64 * slices)(*lens_l)
    dec = b.Etc1sDecoder(n_cb, n_cb, enc["endpoints"], enc["selectors"], enc["tables"])
    nblk = nb * nb * slices
    res = {"workload": "ETC1S %dx%d blocks x %d DIFFERENT slices (procedural images, seeds 4..%d, 15-60 %% flat regions), %d-entry codebooks"
                       % (nb, nb, slices, 3 + slices, n_cb),
           "compressed_bytes_per_slice_min_max": [min(lens_l), max(lens_l)], "bits_per_block": 8.0 * sum(lens_l) / nblk}
    buf = ctypes.create_string_buffer(data, len(data))
    import torch
    for tname, t, ob in (("etc1", 3, 8), ("bc1", 6, 8), ("rgba", 0, 64)):
        if t == 0 and nblk * 64 > (8 << 30):
            continue
        out = torch.empty(nblk * ob, dtype=torch.uint8).pin_memory()
        best = None
        for r in range(reps + 1):
            t0 = time.perf_counter()
            st = L.b2bu_etc1s_transcode_slices(dec._h, t, nb, nb, buf, len(data), ofs, lens, slices, out.data_ptr(), nblk * ob)
            wall = time.perf_counter() - t0
            assert st == 0, st
            k2, k3, d2h, nbk = ctypes.c_float(), ctypes.c_float(), ctypes.c_float(), ctypes.c_uint64()
            L.b2bu_etc1s_last_timing(dec._h, ctypes.byref(k2), ctypes.byref(k3), ctypes.byref(d2h), ctypes.byref(nbk))
            if r and (best is None or k2.value + k3.value < best[0] + best[1]):
                best = (k2.value, k3.value, d2h.value, wall)
        k2, k3, d2h, wall = best
        # parity on slice 0 and the last slice against the oracle's serial decode
        h = ctypes.c_void_p()
        assert orc.orc_etc1s_open(n_cb, n_cb, enc["endpoints"], len(enc["endpoints"]), enc["selectors"], len(enc["selectors"]), enc["tables"],
                                  len(enc["tables"]), 0, ctypes.byref(h)) == 0
        per = nb * nb * ob
        got = out.numpy()
        ok = True
        for k in (0, slices - 1):                       # first and last slice (different streams) against the oracle's serial decode
            one = ec.slice_bytes(enc, k)
            if t == 3:
                e, want = ec.oracle_etc1(orc, h, nb, nb, one)
            elif t == 6:
                e, want = ec.oracle_bc1(orc, h, nb, nb, one)   # EXTENSION: no BC1 in the reference, the oracle function is the definition
            else:
                e, want = ec.oracle_rgba(orc, h, nb, nb, one)
            ok = ok and e == 0 and got[k * per:(k + 1) * per].tobytes() == want
        orc.orc_etc1s_close(h)
        res[tname] = {"entropy_ms": k2, "gather_ms": k3, "d2h_ms": d2h, "wall_ms": wall * 1e3,
                      "device_gtexel_s": nblk * 16 / ((k2 + k3) * 1e-3) / 1e9, "e2e_gtexel_s": nblk * 16 / wall / 1e9,
                      "gather_algo_gb_s": nblk * (4 + ob) / (k3 * 1e-3) / 1e9, "entropy_mblocks_s_per_slice": nb * nb / (k2 * 1e-3) / 1e6,
                      "parity_vs_oracle": bool(ok)}
        del out
    # CPU port on the same slice, one thread per slice
    t0 = time.perf_counter()
    h = ctypes.c_void_p()
    orc.orc_etc1s_open(n_cb, n_cb, enc["endpoints"], len(enc["endpoints"]), enc["selectors"], len(enc["selectors"]), enc["tables"], len(enc["tables"]), 0,
                       ctypes.byref(h))
    ec.oracle_etc1(orc, h, nb, nb, ec.slice_bytes(enc, 0))
    orc.orc_etc1s_close(h)
    dt = time.perf_counter() - t0
    res["cpu_port_one_thread_gtexel_s"] = nb * nb * 16 / dt / 1e9
    dec.close()
    return res


def bench_c5_mixed_batch(L, b, torch, dist, rank, world, payload, total_images, status, sh):
    """configs[4]: a batch of mixed UASTC / ETC1S 2048x2048 textures sharded by image over the ranks (image i -> rank i mod world,
    basisu_rs_b200.shard.plan_shards), no collective on the data path.  BASELINE names 4096 images on 8 GPUs = 512 per GPU; the
    default is 512 per GPU at every rank count (weak scaling).  UASTC images (even i, payload seeded by i) go to RGBA and to BC7
    through the device-resident slice-table entry point (the rank's images in ONE launch per target); ETC1S images (odd i, every
    one a different procedural slice) go to RGBA through b2bu_etc1s_transcode_slices (the reference has no ETC1S -> BC7), whose
    device phases (K2 + K3) are timed by the library.  Every rank checks its first and last image of either kind against the
    oracle.  End to end: the same images through the host-pointer calls with pinned buffers."""
    import etc1s_common as ec
    from etc1s_synth import encode, make_codebooks, make_indices
    from basisu_rs_b200.shard import plan_shards
    nb = 512                                              # 2048 texels = 512 blocks per edge
    nblk = nb * nb
    tex = nblk * 16
    mine = plan_shards([1.0] * total_images, world)[rank]
    # kinds alternate along every rank's own list (image i = r + G m is UASTC when m + r is even): with "i even" and an even
    # number of ranks, all UASTC images would land on the even ranks and all ETC1S images on the odd ones
    is_uastc = lambda i: ((i // world) + (i % world)) % 2 == 0
    ua = [i for i in mine if is_uastc(i)]
    es = [i for i in mine if not is_uastc(i)]
    res = {"workload": "%d textures of 2048x2048, half UASTC half ETC1S (alternating on every rank), image i on rank i mod %d: %d per GPU" % (total_images, world, len(mine)),
           "images_total": total_images, "images_this_rank": len(mine)}
    orc_u = load_oracle()
    cores = max(1, (os.cpu_count() or 1) // world)
    ok_all = True
    t_rgba = t_bc7 = 0.0
    e2e_u = {0: (0.0, 0), 2: (0.0, 0)}
    if ua:
        nimg = len(ua)
        allb = np.empty((nimg * nblk, 16), dtype=np.uint8)
        for k, i in enumerate(ua):
            allb[k * nblk:(k + 1) * nblk] = make_payload(payload, nblk, seed=1000 + i)
        d_all = torch.from_numpy(allb.reshape(-1)).cuda()
        for tgt, ob in ((0, 64), (2, 16)):
            d_out = torch.empty(nimg * nblk * ob, dtype=torch.uint8, device="cuda")
            sl = (b.SliceDev * nimg)()
            for k in range(nimg):
                sl[k] = b.SliceDev(k * nblk * 16, k * nblk * ob, nblk, nb, 0)
            best = None
            for rep in range(3):
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                assert L.b2bu_uastc_transcode_slices_dev(tgt, d_all.data_ptr(), d_out.data_ptr(), sl, nimg, status.data_ptr(), sh) == 0
                c.record()
                torch.cuda.synchronize()
                t = a.elapsed_time(c) * 1e-3
                best = t if best is None or (rep and t < best) else best
            if tgt == 0:
                t_rgba = best
            else:
                t_bc7 = best
            for k in (0, nimg - 1):                      # parity: first and last image of this rank
                want = np.zeros(nblk * ob, dtype=np.uint8)
                orc_u.orc_uastc_transcode_slice(tgt, allb[k * nblk:].ctypes.data, nblk * 16, nb, want.ctypes.data, cores, None)
                ok_all = ok_all and bool((d_out[k * nblk * ob:(k + 1) * nblk * ob].cpu().numpy() == want).all())
            del d_out
            # end to end: image by image through the host-pointer call (pinned buffers), a bounded number of images
            m = min(nimg, 64 if tgt == 0 else 128)
            h_in = torch.from_numpy(allb[: m * nblk].reshape(-1)).pin_memory()
            h_out = torch.empty(m * nblk * ob, dtype=torch.uint8).pin_memory()
            fb = ctypes.c_uint64(0)

            def host_images():
                for k in range(m):
                    if tgt == 0:
                        st = L.b2bu_uastc_decode_rgba(h_in.data_ptr() + k * nblk * 16, nblk * 16, nb, h_out.data_ptr() + k * nblk * ob, nblk * 16, ctypes.byref(fb))
                    else:
                        st = L.b2bu_uastc_transcode(tgt, h_in.data_ptr() + k * nblk * 16, nblk * 16, h_out.data_ptr() + k * nblk * ob, nblk * ob, ctypes.byref(fb))
                    assert st == 0, st
            host_images()
            t0 = time.perf_counter()
            host_images()
            e2e_u[tgt] = (time.perf_counter() - t0, m)
            del h_in, h_out
        del d_all, allb
    # ---- ETC1S share: every image its own procedural 512x512-block slice (same codebooks), the rank's slices in calls of <= 128 ----
    t_etc = 0.0
    e2e_etc = 0.0
    if es:
        orc = ec.bind(orc_u)
        n_cb = 2048
        ep_cb, sel_cb = make_codebooks(n_cb, n_cb, seed=9)
        eis, sis = [], []
        for i in es:
            e1, s1 = make_indices(nb, nb, 1, n_cb, n_cb, seed=10 + i, flat=0.15 + 0.45 * ((i * 7) % 64) / 63.0)
            eis.append(e1[0])
            sis.append(s1[0])
        enc = encode(orc, ep_cb, sel_cb, np.stack(eis), np.stack(sis), nb, nb, 64, False, False)
        del eis, sis
        dec = b.Etc1sDecoder(n_cb, n_cb, enc["endpoints"], enc["selectors"], enc["tables"])
        CH = max(1, len(es))                          # all of the rank's slices in one call: K2 packs them two per SM
        out = torch.empty(nblk * 64 * min(CH, len(es)), dtype=torch.uint8).pin_memory()
        h = ctypes.c_void_p()
        assert orc.orc_etc1s_open(n_cb, n_cb, enc["endpoints"], len(enc["endpoints"]), enc["selectors"], len(enc["selectors"]), enc["tables"],
                                  len(enc["tables"]), 0, ctypes.byref(h)) == 0
        k2s = k3s = 0.0
        for c0 in range(0, len(es), CH):
            ks = list(range(c0, min(c0 + CH, len(es))))
            parts, ofs_l, lens_l, pos = [], [], [], 0
            for k in ks:
                one = ec.slice_bytes(enc, k)
                pad = (-len(one)) % 16
                parts.append(one + b"\0" * pad)
                ofs_l.append(pos)
                lens_l.append(len(one))
                pos += len(one) + pad
            data = b"".join(parts)
            buf = ctypes.create_string_buffer(data, len(data))
            ofs = (ctypes.c_uint64 * len(ks))(*ofs_l)
            lens = (ctypes.c_uint64 * len(ks))(*lens_l)
            best = None
            for rep in range(2):
                t0 = time.perf_counter()
                assert L.b2bu_etc1s_transcode_slices(dec._h, 0, nb, nb, buf, len(data), ofs, lens, len(ks), out.data_ptr(), nblk * 64 * len(ks)) == 0
                wall = time.perf_counter() - t0
                k2, k3 = ctypes.c_float(), ctypes.c_float()
                L.b2bu_etc1s_last_timing(dec._h, ctypes.byref(k2), ctypes.byref(k3), None, None)
                if best is None or k2.value + k3.value < best[0] + best[1]:
                    best = (k2.value, k3.value, wall)
            k2s += best[0]
            k3s += best[1]
            e2e_etc += best[2]
            got = out.numpy()
            for k in ks:
                if k in (0, len(es) - 1):                # parity: first and last ETC1S image of this rank
                    e, want = ec.oracle_rgba(orc, h, nb, nb, ec.slice_bytes(enc, k))
                    j = k - c0
                    ok_all = ok_all and e == 0 and got[j * nblk * 64:(j + 1) * nblk * 64].tobytes() == want
        orc.orc_etc1s_close(h)
        t_etc = (k2s + k3s) * 1e-3
        res["etc1s_entropy_ms_this_rank"] = k2s
        res["etc1s_gather_ms_this_rank"] = k3s
        dec.close()
        del out
    tu_r, mu_r = e2e_u[0]
    tu_b, mu_b = e2e_u[2]
    times = torch.tensor([t_rgba, t_bc7, t_etc, t_rgba + t_etc, tu_r, tu_b, e2e_etc], dtype=torch.float64, device="cuda")
    counts = torch.tensor([len(ua), len(es), mu_r, mu_b], dtype=torch.float64, device="cuda")
    okt = torch.tensor([1.0 if ok_all else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)      # timing only: max over ranks
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    tr, tb, te, tm, er, eb, ee = [float(x) for x in times.tolist()]
    nu, ne, mr, mb = [float(x) for x in counts.tolist()]
    res.update({"uastc_to_rgba_gtexel_s": nu * tex / tr / 1e9 if tr else None, "uastc_to_bc7_gtexel_s": nu * tex / tb / 1e9 if tb else None,
                "etc1s_to_rgba_device_gtexel_s": ne * tex / te / 1e9 if te else None,
                "mixed_to_rgba_device_gtexel_s": (nu + ne) * tex / tm / 1e9 if tm else None,
                "e2e": {"uastc_to_rgba_gtexel_s": mr * tex / er / 1e9 if er else None, "uastc_to_rgba_images_timed": int(mr),
                        "uastc_to_bc7_gtexel_s": mb * tex / eb / 1e9 if eb else None, "uastc_to_bc7_images_timed": int(mb),
                        "etc1s_to_rgba_gtexel_s": ne * tex / ee / 1e9 if ee else None, "etc1s_images_timed": int(ne),
                        "api": "b2bu_uastc_decode_rgba / b2bu_uastc_transcode per image, b2bu_etc1s_transcode_slices over all slices of the rank; pinned host buffers, wall clock, max over ranks"},
                "parity_first_and_last_image_of_every_rank": bool(okt.item() > 0.5), "n_gpus": world})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--target", default="astc", choices=list(TARGET_NAMES))
    ap.add_argument("--payload", default="kat-shuffled", choices=["kat-shuffled", "kat-coherent", "random"])
    ap.add_argument("--blocks", type=int, default=2048 * 2048)        # 8192 x 8192 texels
    ap.add_argument("--bpr", type=int, default=2048)
    ap.add_argument("--ring", type=int, default=4, help="distinct device buffer sets cycled through (ring * 128 MiB > L2)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--prewarm-s", type=float, default=0.25, help="seconds of untimed launches before the counted warm-up")
    ap.add_argument("--cpu-sample-blocks", type=int, default=1 << 20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--all-targets", action="store_true", help="also report the other targets / payloads in 'extra'")
    ap.add_argument("--c5-images", type=int, default=0, help="images in the mixed batch of configs[4]; 0 = 512 per GPU (BASELINE: 4096 on 8 GPUs)")
    ap.add_argument("--configs", default="c3,c4,c5", help="extra BASELINE configs measured on rank 0 and reported under 'configs' (c3 = BC7 mip chain, "
                    "c4 = ETC1S slices, c5 = mixed batch sharded by image over the ranks); 'none' to skip")
    ap.add_argument("--c4-blocks", type=int, default=1024, help="ETC1S slice edge in blocks (1024 = 4096x4096 texels)")
    ap.add_argument("--c4-slices", type=int, default=64)
    args = ap.parse_args()
    capture_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import basisu_rs_b200 as b

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: basisu_rs_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa, numa_why = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a mismatched collective must fail in minutes, not after NCCL's default 10-minute watchdog on every rank
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=420))
    L = b.lib()
    assert L.b2bu_init(local_rank) == 0, L.b2bu_last_cuda_error().decode()

    target = TARGET_NAMES[args.target]
    n = args.blocks
    ob = OUT_BYTES[target]
    blocks = make_payload(args.payload, n, seed=rank)          # each GPU owns its own texture (sharded by image)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    # ---- device-resident run: ring of distinct input/output buffers, total footprint > L2 ----
    d_in = [torch.from_numpy(blocks.reshape(-1)).cuda() for _ in range(args.ring)]
    d_out = [torch.empty(n * ob, dtype=torch.uint8, device="cuda") for _ in range(args.ring)]
    status = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert L.b2bu_status_reset_dev(status.data_ptr(), sh) == 0

    def step(i):
        k = i % args.ring
        st = L.b2bu_uastc_transcode_dev(target, d_in[k].data_ptr(), n * 16, args.bpr, d_out[k].data_ptr(), n * ob, status.data_ptr(), sh)
        assert st == 0, L.b2bu_last_cuda_error().decode()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # pre-warm by TIME before the counted warm-up: clocks, caches and the instruction cache settle over ~0.2 s of launches (with a
    # handful of 55 us launches only, one hiccup on one rank moved the max-over-ranks by a quarter)
    t_pre = time.perf_counter()
    prewarm = 0
    while time.perf_counter() - t_pre < args.prewarm_s:
        for i in range(64):
            step(prewarm + i)
        prewarm += 64
        torch.cuda.synchronize()
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.b2bu_launch_count()
    # the contract's number: ONE event pair around exactly `steps` steps, nothing else on the stream in between (an event record
    # between two launches would also keep the next launch's prologue from overlapping the previous launch's tail, which the
    # kernels allow through programmatic dependent launch)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step(i)
    ev1.record(stream)
    barrier()
    launches = L.b2bu_launch_count() - launches0
    # a second pass of the same steps with an event after every launch: the distribution of single steps (median, min, max)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    marks[0].record(stream)
    for i in range(args.steps):
        step(i)
        marks[i + 1].record(stream)
    barrier()
    if args.steps * 1e-4 < 0.3:                    # a short timed region: keep sampling clocks under the same load for a moment
        t_s = time.perf_counter()
        while time.perf_counter() - t_s < 0.3:
            for i in range(64):
                step(i)
            torch.cuda.synchronize()
    clocks = sampler.finish()
    ms = ev0.elapsed_time(ev1)
    per_step_ms = np.array([marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)])
    bad = ctypes.c_uint64(0)
    assert L.b2bu_status_read_dev(status.data_ptr(), sh, ctypes.byref(bad)) == 0, "payload contained invalid blocks"
    t = torch.tensor([ms, float(np.median(per_step_ms)) * args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t[0].item())
    ms_per_step = ms_max / args.steps
    ms_per_step_median = float(t[1].item()) / args.steps
    value = world * n * 16 / (ms_per_step * 1e-3) / 1e9

    # (the end-to-end pass runs BEFORE the CPU baseline: sixteen busy host threads for tens of seconds leave a container
    # with a CPU quota throttled, and the call under test is ~80 driver calls per step from one host thread -- measured on
    # one box: 26.5 Gtexel/s right behind the CPU phase against 38.6 without it)
    # ---- what the host side can feed at this rank count (every rank at once) ----
    ceiling = pcie_ceiling(torch, dist, world)

    # ---- end to end: host pinned buffers through the reference-facing C-ABI call ----
    h_in = torch.from_numpy(blocks.reshape(-1)).pin_memory()
    h_out = torch.empty(n * ob, dtype=torch.uint8).pin_memory()
    fb = ctypes.c_uint64(0)

    def e2e_step():
        if target == 0:
            st = L.b2bu_uastc_decode_rgba(h_in.data_ptr(), n * 16, args.bpr, h_out.data_ptr(), n * 16, ctypes.byref(fb))
        else:
            st = L.b2bu_uastc_transcode(target, h_in.data_ptr(), n * 16, h_out.data_ptr(), n * ob, ctypes.byref(fb))
        assert st == 0, (st, L.b2bu_last_cuda_error().decode())

    for _ in range(3):
        e2e_step()
    barrier()
    l0 = L.b2bu_launch_count()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_launches = L.b2bu_launch_count() - l0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * 16 * args.e2e_steps / float(te.item()) / 1e9

    # ---- parity on the exact benchmark buffers (rank 0, bounded oracle sample + golden tiling) ----
    parity = None
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        orc = load_oracle()
        cores = os.cpu_count() or 1
        ns = min(n, args.cpu_sample_blocks)
        ns -= ns % args.bpr if target == 0 and ns >= args.bpr else 0
        gt, dt, reps, want = time_oracle(orc, target, blocks[:ns], args.bpr, cores)
        got = d_out[(args.steps - 1) % args.ring][: ns * ob].cpu().numpy()
        parity = bool((got == want).all()) if target != 0 else None
        g1, dt1, reps1, _ = time_oracle(orc, target, blocks[: max(ns // 8, 1)], args.bpr, 1, min_seconds=1.0)
        cpu = {"value": gt, "unit": "Gtexel/s", "cores": cores, "kind": "port",
               "sample": f"{ns} of {n} blocks x {reps} reps in {dt:.1f} s, {cores} threads (static block partition); 1 thread: {g1:.4f} Gtexel/s",
               "single_thread_value": g1}

    e2e_parity = None
    if rank == 0 and parity is not None:
        e2e_parity = bool((h_out[: want.size].numpy() == want).all())

    extra = {}
    if args.all_targets and rank == 0:
        for tn, tt in TARGET_NAMES.items():
            for pk in ("kat-shuffled", "kat-coherent", "random"):
                blk = make_payload(pk, n)
                di = torch.from_numpy(blk.reshape(-1)).cuda()
                do = [torch.empty(n * OUT_BYTES[tt], dtype=torch.uint8, device="cuda") for _ in range(2)]
                for i in range(3):
                    L.b2bu_uastc_transcode_dev(tt, di.data_ptr(), n * 16, args.bpr, do[i % 2].data_ptr(), n * OUT_BYTES[tt], status.data_ptr(), sh)
                torch.cuda.synchronize()
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for i in range(20):
                    L.b2bu_uastc_transcode_dev(tt, di.data_ptr(), n * 16, args.bpr, do[i % 2].data_ptr(), n * OUT_BYTES[tt], status.data_ptr(), sh)
                c.record(stream)
                torch.cuda.synchronize()
                us = a.elapsed_time(c) / 20 * 1e3
                extra[f"{tn}/{pk}"] = {"us_per_launch": us, "gtexel_s": n * 16 / us / 1e3, "algo_gb_s": n * ALGO_BYTES[tt] / us / 1e3}
                del di, do

    cfgs = {}
    if rank == 0 and args.configs:
        want_cfg = set(args.configs.split(","))
        if "c3" in want_cfg:
            cfgs["c3_bc7_mip_chain"] = bench_c3_bc7_mips(L, b, torch, args.payload, 50, status, sh)
        if "c4" in want_cfg:
            cfgs["c4_etc1s"] = bench_c4_etc1s(L, b, args.c4_blocks, args.c4_slices)
    if args.configs and "c5" in args.configs.split(","):
        c5 = bench_c5_mixed_batch(L, b, torch, dist, rank, world, args.payload, args.c5_images or 512 * world, status, sh)     # every rank takes part
        if rank == 0:
            cfgs["c5_mixed_batch"] = c5
    if world > 1:
        dist.barrier()

    int_bound = None
    if rank == 0:
        # secondary bound of north_star's roofline definition: integer ops / INT throughput, with the op count per block
        # fixed by SURVEY.md section 8d and the INT throughput measured on this device by the library's probe (inline-PTX
        # instruction streams whose SASS holds exactly the counted instructions).  The SM issues 128 thread-instructions per
        # clock, 64 on the alu pipe (LOP3 / SHF / PRMT / SEL / ISETP) and 64 on the fma pipe (IMAD); the bound uses the
        # balanced two-pipe peak -- the stricter reading: a kernel that is all bit logic can only use half of it.
        alu, mix = ctypes.c_double(), ctypes.c_double()
        if L.b2bu_probe_int_peak(ctypes.byref(alu), ctypes.byref(mix)) == 0 and mix.value > 0:
            ops = INT_OPS[target]
            t_int = n * ops / (mix.value * 1e12)
            t_alu = n * ops / (alu.value * 1e12)
            int_bound = {"alu_pipe_tops": alu.value, "alu_fma_mix_tops": mix.value, "contract_ops_per_block": ops,
                         "bound_us_at_mix_peak": t_int * 1e6, "frac_of_int_bound": t_int / (ms_per_step * 1e-3),
                         "bound_us_at_alu_pipe_peak": t_alu * 1e6,
                         "note": "SURVEY 8d op counts are estimates of a minimal table-driven formulation; frac = bound / measured time"}
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        achieved = n * ALGO_BYTES[target] / (ms_per_step * 1e-3) / 1e9
        t_hbm_us = n * ALGO_BYTES[target] / (peak * 1e9) * 1e6
        # north_star: the roofline of a path is the SLOWER of bytes / HBM bandwidth and integer ops / INT throughput
        t_roof_us = max(t_hbm_us, int_bound["bound_us_at_mix_peak"]) if int_bound else t_hbm_us
        per_path = {"hbm_bound_us": t_hbm_us, "int_bound_us": int_bound["bound_us_at_mix_peak"] if int_bound else None,
                    "roofline_us": t_roof_us, "measured_us": ms_per_step * 1e3, "frac_of_per_path_roofline": t_roof_us / (ms_per_step * 1e3),
                    "binding": "int" if int_bound and int_bound["bound_us_at_mix_peak"] > t_hbm_us else "hbm"}
        e2e_bytes_gbs = e2e_value * (16 + ob) / 16.0                           # GB/s over PCIe, both directions together
        e2e_ceiling = max(1e-9, min(ceiling["h2d_gbs_all_ranks"], ceiling["d2h_gbs_all_ranks"] * 16.0 / ob))
        line = {
            "metric": "Gtexels/s UASTC->%s" % args.target.upper(),
            "value": value, "unit": "Gtexel/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "ms_per_step_median": ms_per_step_median,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, n, ob),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.target), "peak_source": peak_src,
                         "algorithmic_bytes_per_block": ALGO_BYTES[target], "kernel": "uastc_sorted_kernel<%s>" % args.target.upper(),
                         "int_bound_us": per_path["int_bound_us"], "hbm_bound_us": t_hbm_us,
                         "frac_of_per_path_roofline": per_path["frac_of_per_path_roofline"], "int_bound": int_bound},
            "int_bound": int_bound,
            "per_path_roofline": per_path,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Gtexel/s", "h2d_bytes_per_step": n * 16, "d2h_bytes_per_step": n * ob,
                    "steps": args.e2e_steps, "launches": int(e2e_launches), "api": "b2bu_uastc_transcode (pinned host buffers)",
                    "pcie_ceiling": ceiling,
                    "pcie_ceiling_gbs": min(ceiling["h2d_gbs_all_ranks"], ceiling["d2h_gbs_all_ranks"]),
                    # 16 B up and `ob` B down per block of 16 texels: the copies alone allow this many Gtexel/s
                    "gtexel_s_at_pcie_ceiling": e2e_ceiling, "frac_of_pcie_ceiling": e2e_value / e2e_ceiling,
                    "pcie_gbs_both_directions": e2e_bytes_gbs,
                    "host_numa_node_rank0": numa, "host_numa_binding": numa_why},
            "gpu_launches": int(launches),
            "timing": {"prewarm_launches": int(prewarm), "prewarm_s": args.prewarm_s, "per_step_us_min": float(per_step_ms.min() * 1e3),
                       "per_step_us_median": float(np.median(per_step_ms) * 1e3), "per_step_us_max": float(per_step_ms.max() * 1e3),
                       "note": "per_step_*: rank 0, a second pass of the same steps with an event after every launch; ms_per_step is the "
                               "one event pair around all steps of the first pass, max over ranks",
                       "overlap": "consecutive launches of a stream overlap by the kernel's prologue (programmatic dependent launch: "
                                  "griddepcontrol.wait in front of the first global access keeps stream order), so ms_per_step of "
                                  "back-to-back steps is below the duration of a launch measured on its own (per_step_*, ncu)"},
            "clocks": clocks,
            "parity": {"device_vs_oracle_sample": parity, "e2e_vs_oracle_sample": e2e_parity},
        }
        if extra:
            line["extra"] = extra
        if cfgs:
            line["configs"] = cfgs
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
