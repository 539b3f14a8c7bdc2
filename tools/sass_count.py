#!/usr/bin/env python3
"""Tuning aid: per UASTC mode, the SASS instruction count of the specialised transcode code by pipe
(alu = LOP3/SHF/IADD3/PRMT/SEL/ISETP/LEA..., fma = IMAD/FMUL.., lsu = LDS/STS/LDG.., xu = POPC/BREV/FLO/I2F..).
usage: tools/sass_count.py astc|bc7|rgba"""
import re, subprocess, sys, pathlib, collections
ROOT = pathlib.Path(__file__).resolve().parent.parent
tgt = sys.argv[1]
T = {"rgba": 0, "astc": 1, "bc7": 2, "etc1": 3, "etc2": 4}[tgt]
modes = [0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18]
src = ['#include "%s/basisu_rs_b200/csrc/uastc_device.cuh"\nusing namespace b2bu;\n__device__ DevTables g_t;\n' % ROOT]
for m in modes:
    if tgt in ("astc", "bc7"):
        body = "o.v = %s_block<%d>(b, T, pat, cs);" % (tgt, m)
        store = "out[i] = o.v;"
    else:
        body = "Canon c; canon_front<" + str(m) + ">(b, T, pat, cs, c); out[i] = make_uint4(c.lo_rb[0] ^ c.lo_rb[1] ^ c.lo_rb[2], c.hi_rb[0]^c.hi_rb[1]^c.hi_rb[2]^c.lo_ga[0]^c.lo_ga[1]^c.lo_ga[2]^c.hi_ga[0]^c.hi_ga[1]^c.hi_ga[2], c.w0.x^c.w0.y^c.w0.z^c.w0.w^c.pw, c.w1.x^c.w1.y^c.w1.z^c.w1.w^c.mrb^c.mga);"
        store = ""
    src.append("extern \"C\" __global__ void k_m%d(const uint4* in, uint4* out) { __shared__ DevTables T; if (threadIdx.x == 0) T = g_t; __syncthreads(); int i = threadIdx.x; uint4 b = in[i]; BlockOut o; uint32_t pat, cs; if (!header_ok<%d>(b, pat, cs)) return; %s %s }\n" % (m, m, body, store))
if tgt == "rgba":
    for name, targs in (("interp_single", "false,false"), ("interp_multi", "true,false"), ("interp_dual", "false,true")):
        src.append("extern \"C\" __global__ void k_%s(const Canon* in, uint4* out) { Canon c = in[threadIdx.x]; StridedRowSink s{out + threadIdx.x, 1024}; interp_rows<%s>(c, s); }\n" % (name, targs))
p = pathlib.Path("/tmp/sc/k.cu"); p.write_text("".join(src))
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-cubin", "-o", "/tmp/sc/k.cubin", str(p)], check=True)
sass = subprocess.run(["cuobjdump", "-sass", "/tmp/sc/k.cubin"], stdout=subprocess.PIPE, text=True).stdout
cls = lambda op: ("fma" if op.startswith(("IMAD", "FMUL", "FADD", "FFMA", "HFMA")) else "lsu" if op.startswith(("LDS", "STS", "LDG", "STG", "LDC", "ATOM", "LD", "ST")) else
                  "xu" if op.startswith(("POPC", "BREV", "FLO", "I2F", "F2I", "MUFU", "FRND")) else "ctl" if op.startswith(("BRA", "EXIT", "BAR", "BSSY", "BSYNC", "NOP", "S2R", "CS2R", "WARPSYNC", "S2UR", "RET", "CALL")) else
                  "uni" if op.startswith(("U", "R2UR")) else "alu")
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur: counts[cur][cls(m.group(1))] += 1; counts[cur]["total"] += 1
for k, c in sorted(counts.items(), key=lambda kv: (len(kv[0]), kv[0])):
    print("%-18s total %4d  alu %4d  fma %3d  lsu %3d  xu %2d  uni %2d ctl %2d" % (k, c["total"], c["alu"], c["fma"], c["lsu"], c["xu"], c["uni"], c["ctl"]))
print("mean total %.0f  mean alu %.0f" % (sum(c["total"] for c in counts.values()) / len(counts), sum(c["alu"] for c in counts.values()) / len(counts)))
