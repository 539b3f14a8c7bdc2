#!/usr/bin/env python3
"""Emit basisu_rs_b200/csrc/device_tables_gen.inc: the initialiser of b2bu::DevTables.

Inputs: tools/format_tables.json (format constants) + derivations done here:
  * base-3 / base-5 digit tables (uastc.rs:629-685 semantics: digit_k = (g / b^k) % b)
  * ASTC endpoint unquantisation for the five trit/quint ranges UASTC uses (uastc.rs:585-614)
  * BC7 solid-colour endpoint LUTs by brute force (bc7.rs:1158-1250)
  * p-bit search terms evaluated in IEEE f32 with numpy (bc7.rs:408-553), one op at a time
  * EAC alpha lerp fractions in f32 (etc.rs:301-307)
"""
import json, pathlib
import numpy as np
HERE = pathlib.Path(__file__).resolve().parent
T = json.loads((HERE / "format_tables.json").read_text())
f32 = np.float32

def chunks(v, n): return [v[i:i + n] for i in range(0, len(v), n)]

def digits(g, base, count, bits):
    x = 0
    for k in range(count):
        x |= (g % base) << (bits * k)
        g //= base
    return x

BISE = {7: (2, 1, 0, "b000b0bb0", 93), 12: (3, 0, 1, "cb0000cbc", 26), 13: (4, 1, 0, "dcb000dcb", 22),
        18: (5, 0, 1, "edcb0000e", 6), 19: (6, 1, 0, "fedcb000f", 5)}

def unquant(range_index, d, m):
    bits, trits, quints, deq_b, deq_c = BISE[range_index]
    a = 511 if (m & 1) else 0
    b = 0
    for j in range(9):
        b <<= 1
        ch = deq_b[j]
        if ch != "0":
            b |= (m >> (ord(ch) - ord("a"))) & 1
    val = (d * deq_c + b) & 0xFFFF
    val ^= a
    return ((a & 0x80) | (val >> 2)) & 0xFF

def unq_table(r):
    bits, trits, quints, _, _ = BISE[r]
    nd = 3 if trits else 5
    return [unquant(r, d, m) for d in range(nd) for m in range(1 << bits)]

def pack_pat(rows): return [sum((v & 3) << (2 * i) for i, v in enumerate(r)) for r in rows]

def derive_m6():
    w = 21; lo = []; hi = []
    for c in range(256):
        best = (0, 0); be = 1 << 30
        for l in range(128):
            for h in range(l, 128):
                k = ((l << 1) * (64 - w) + (h << 1) * w + 32) >> 6
                e = (k - c) ** 2
                if e < be: be = e; best = (l, h)
        lo.append(best[0]); hi.append(best[1])
    return [0] + lo, [0] + hi          # entry 0 is the p=1 alias of c=0 (bc7.rs:1126-1131)

def derive_m5():
    w = 21; lo = []; hi = []
    for c in range(256):
        best = (0, 0); be = 1 << 30
        for l in range(128):
            for h in range(l, 128):
                k = (((l << 1) | (l >> 6)) * (64 - w) + ((h << 1) | (h >> 6)) * w + 32) >> 6
                e = (k - c) ** 2
                if e < be: be = e; best = (l, h)
        lo.append(best[0]); hi.append(best[1])
    return lo, hi

def quantise(v, p, total_bits):
    """bc7.rs:509-514: (((x*scalep - p)/2 + .5) as i32 * 2 + p).clamp(p, iscalep-1+p) in f32."""
    iscalep = (1 << total_bits) - 1
    x = f32(v) / f32(255.0)
    t = x * f32(iscalep)
    t = t - f32(p)
    t = t / f32(2.0)
    t = t + f32(0.5)
    q = int(t) * 2 + p
    return min(max(q, p), iscalep - 1 + p)

def scaled(q, total_bits):
    s = (q << (8 - total_bits)) & 0xFF
    return s | (s >> (total_bits & 7))

def carr(name, vals, per=24):
    rows = chunks([str(v) for v in vals], per)
    return "  /* %s */ {%s},\n" % (name, ",\n    ".join(",".join(r) for r in rows))

def main():
    out = []
    out.append(carr("mode_lut", T["MODE_LUT"]))
    out.append(carr("trit_dec", [digits(g, 3, 5, 2) for g in range(256)]))
    out.append(carr("quint_dec", [digits(g, 5, 3, 3) for g in range(128)]))
    for r in (7, 12, 13, 18, 19):
        out.append(carr("unq%d" % r, unq_table(r)))
    p2 = chunks(T["PATTERNS_2"], 16); p3 = chunks(T["PATTERNS_3"], 16); p23 = chunks(T["PATTERNS_2_3"], 16)
    out.append(carr("pat2", pack_pat(p2) + [0, 0]))
    out.append(carr("pat3", pack_pat(p3) + [0]))
    out.append(carr("pat23", pack_pat(p23) + [0]))
    out.append(carr("pat2w3", ["0x%Xull" % sum(7 << (3 * i) for i, v in enumerate(r) if v == 1) for r in p2] + ["0", "0"]))
    def anc(rows, pats):
        res = []
        for a, pat in zip(rows, pats):
            assert 0 in a and len(set(a)) == len(a)
            for s, t in enumerate(a):
                assert pat[t] == s, "anchor must belong to its subset"
            nz = sorted(x for x in a if x != 0)
            res.append(nz[0] | ((nz[1] if len(nz) > 1 else 0) << 4))
        return res
    out.append(carr("anc2", anc(chunks(T["PATTERNS_2_ANCHORS"], 2), p2) + [0, 0]))
    out.append(carr("anc3", anc(chunks(T["PATTERNS_3_ANCHORS"], 3), p3) + [0]))
    out.append(carr("anc23", anc(chunks(T["PATTERNS_2_3_ANCHORS"], 2), p23) + [0]))
    out.append(carr("trit_enc", T["ASTC_TRIT_ENCODE_LUT"] + [0]))
    out.append(carr("quint_enc", T["ASTC_QUINT_ENCODE_LUT"] + [0, 0, 0]))
    out.append(carr("seed2", T["PATTERNS_2_ASTC_INDEX_10"] + [0, 0]))
    out.append(carr("seed3", T["PATTERNS_3_ASTC_INDEX_10"] + [0]))
    out.append(carr("seed23", T["PATTERNS_2_3_ASTC_INDEX_10"] + [0]))
    a2 = chunks(T["PATTERNS_2_BC7_ANCHORS"], 2); a3 = chunks(T["PATTERNS_3_BC7_ANCHORS"], 3)
    perm3 = chunks(T["PATTERNS_3_BC7_TO_ASTC_PERMUTATIONS"], 3)
    perm23 = chunks(T["PATTERNS_2_3_BC7_TO_ASTC_PERMUTATIONS"], 3)
    def packp(index, perm, anchors, nsub):
        assert anchors[0] == 0
        v = index | sum(p << (8 + 2 * i) for i, p in enumerate(perm))
        v |= anchors[1] << 16
        if nsub == 3: v |= anchors[2] << 20
        return v | (nsub << 24)
    bc7p2 = [packp(i, [1, 0] if inv else [0, 1], a2[i], 2) for i, inv in chunks(T["PATTERNS_2_BC7_INDEX_INV"], 2)]
    bc7p3 = [packp(i, perm3[p], a3[i], 3) for i, p in chunks(T["PATTERNS_3_BC7_INDEX_PERM"], 2)]
    bc7p23 = [packp(i, perm23[p], a3[i], 3) for i, p in chunks(T["PATTERNS_2_3_BC7_INDEX_PERM"], 2)]
    out.append(carr("bc7p2", bc7p2 + [0, 0])); out.append(carr("bc7p3", bc7p3 + [0])); out.append(carr("bc7p23", bc7p23 + [0]))
    out.append(carr("bc7pat2", pack_pat(chunks(T["PATTERNS_2_BC7"], 16)) + [0, 0]))
    out.append(carr("bc7pat3", pack_pat(chunks(T["PATTERNS_3_BC7"], 16)) + [0]))
    out.append(carr("bc7pat23", pack_pat(chunks(T["PATTERNS_2_3_BC7"], 16)) + [0]))
    m5lo, m5hi = derive_m5(); m6lo, m6hi = derive_m6()
    assert [x for pr in zip(m5lo, m5hi) for x in pr] == T["BC7_MODE_5_OPTIMAL_ENDPOINTS"]
    assert [x for pr in zip(m6lo, m6hi) for x in pr] == T["BC7_MODE_6_OPTIMAL_ENDPOINTS"]
    out.append(carr("m5lo", m5lo)); out.append(carr("m5hi", m5hi))
    out.append(carr("m6lo", m6lo + [0] * 3)); out.append(carr("m6hi", m6hi + [0] * 3))
    pq6 = [[quantise(v, p, 6) for v in range(256)] for p in range(2)]
    pe6 = [[(scaled(pq6[p][v], 6) - v) ** 2 for v in range(256)] for p in range(2)]
    out.append("  /* pq6 */ {{%s},\n   {%s}},\n" % (",".join(map(str, pq6[0])), ",".join(map(str, pq6[1]))))
    out.append("  /* pe6 */ {{%s},\n   {%s}},\n" % (",".join(map(str, pe6[0])), ",".join(map(str, pe6[1]))))
    sq7 = [[quantise(17 * k, p, 7) for k in range(16)] for p in range(2)]
    se7 = [[0] * 16 for _ in range(2)]
    for p in range(2):
        for k in range(16):
            x = f32(17 * k) / f32(255.0)
            dlt = f32(scaled(sq7[p][k], 7)) / f32(255.0) - x
            se7[p][k] = int(np.array(f32(dlt) * f32(dlt), dtype=np.float32).view(np.uint32))
    out.append("  /* sq7 */ {{%s},{%s}},\n" % (",".join(map(str, sq7[0])), ",".join(map(str, sq7[1]))))
    out.append("  /* se7_bits */ {{%s},{%s}},\n" % (",".join("0x%08Xu" % v for v in se7[0]), ",".join("0x%08Xu" % v for v in se7[1])))
    out.append(carr("w5to4", [0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 6, 7, 8, 9, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15]))
    out.append(carr("pe6p", ["0x%08Xu" % (pe6[0][v] | (pe6[1][v] << 16)) for v in range(256)]))
    out.append(carr("q5", [(v * 31 + 127) // 255 for v in range(256)]))      # bc7.rs:264-271 without p-bits, 5 bits
    out.append(carr("q7", [(v * 127 + 127) // 255 for v in range(256)]))     # 7 bits
    etc1_mod = [[-8, -2, 2, 8], [-17, -5, 5, 17], [-29, -9, 9, 29], [-42, -13, 13, 42], [-60, -18, 18, 60],
                [-80, -24, 24, 80], [-106, -33, 33, 106], [-183, -47, 47, 183]]
    eac = [[-3, -6, -9, -15, 2, 5, 8, 14], [-3, -7, -10, -13, 2, 6, 9, 12], [-2, -5, -8, -13, 1, 4, 7, 12],
           [-2, -4, -6, -13, 1, 3, 5, 12], [-3, -6, -8, -12, 2, 5, 7, 11], [-3, -7, -9, -11, 2, 6, 8, 10],
           [-4, -7, -8, -11, 3, 6, 7, 10], [-3, -5, -8, -11, 2, 4, 7, 10], [-2, -6, -8, -10, 1, 5, 7, 9],
           [-2, -5, -8, -10, 1, 4, 7, 9], [-2, -4, -8, -10, 1, 3, 7, 9], [-2, -5, -7, -10, 1, 4, 6, 9],
           [-3, -4, -7, -10, 2, 3, 6, 9], [-1, -2, -3, -10, 0, 1, 2, 9], [-4, -6, -8, -9, 3, 5, 7, 8],
           [-3, -5, -7, -9, 2, 4, 6, 8]]
    out.append("  /* etc1_mod */ {%s},\n" % ",".join("{%s}" % ",".join(map(str, r)) for r in etc1_mod))
    out.append("  /* eac_mod */ {%s},\n" % ",".join("{%s}" % ",".join(map(str, r)) for r in eac))
    amt = []; one_m = []
    for r in eac:
        a = -f32(r[3]) / f32(r[7] - r[3])
        amt.append(int(np.array(a, dtype=np.float32).view(np.uint32)))
        one_m.append(int(np.array(f32(1.0) - a, dtype=np.float32).view(np.uint32)))
    out.append(carr("eac_amt_bits", ["0x%08Xu" % v for v in amt]))
    out.append(carr("eac_1m_amt_bits", ["0x%08Xu" % v for v in one_m]))
    text = "// GENERATED by tools/gen_device_tables.py -- do not edit.\n{\n" + "".join(out) + "}\n"
    (HERE.parent / "basisu_rs_b200/csrc/device_tables_gen.inc").write_text(text)
    print("ok", len(text))

if __name__ == "__main__":
    main()
