"""Builds libb2bu.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

Usage: python -m basisu_rs_b200.build [--force]
The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
import os
import pathlib
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libb2bu.so"
SOURCES = ["uastc_kernels.cu", "etc1s_kernels.cu", "etc1s_host.cu", "basis_file.cu", "capi.cu", "probe.cu", "crc_kernels.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=hidden", "--use_fast_math=false",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc"))
    deps.append(HERE.parent / "include" / "b2bu.h")
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants (e.g. defines=["-DB2BU_TILE=2048"], out="libb2bu_t2048.so")."""
    lib = HERE / out if out else LIB
    if not force and not out and not needs_build():
        return LIB
    objs = []
    procs = []
    builddir = HERE / "build" / (out or "default")
    builddir.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    for src in SOURCES:
        if not (CSRC / src).exists():
            continue
        obj = builddir / (src + ".o")
        cmd = [nvcc] + flags + list(defines) + ["-Xptxas", "-v"] * int(verbose) + ["-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    link = [nvcc, "-shared", "-o", str(lib)] + objs + ["-Xcompiler", "-fPIC", "-lcudart_static", "-lpthread", "-ldl", "-lrt",
                                                       "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return lib


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
