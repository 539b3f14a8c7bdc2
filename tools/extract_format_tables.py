#!/usr/bin/env python3
"""Dev-time helper: read the numeric FORMAT CONSTANTS (UASTC/ASTC/BC7 partition tables,
BISE trit/quint encodings, block-mode words) out of the reference crate and emit them as
neutral JSON (tools/format_tables.json).  The JSON is committed; nothing at build, test or
run time reads /root/reference.  Only numbers are taken -- no code.

Source locations (relative to /root/reference):
  src/uastc.rs:560-577   MODE_LUT            src/uastc.rs:748-811  partition tables + anchors
  src/target_formats/astc.rs:183-193  ASTC partition seeds
  src/target_formats/astc.rs:208,247  BISE quint / trit encode LUTs
  src/target_formats/astc.rs:333-354  UASTC->ASTC block-mode words
  src/target_formats/bc7.rs:582-722   BC7 mode map, partition index/perm/anchor tables
  src/target_formats/bc7.rs:734-1124  BC7 mode-5 / mode-6 solid-colour endpoint LUTs
     (these two are ALSO re-derived by brute force in tools/gen_tables.py and must match)
"""
import json, re, sys, pathlib

REF = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")

def body(text, name):
    m = re.search(r"(?:static|const)\s+" + name + r"\s*:[^=]*=\s*\[(.*?)\];", text, re.S)
    if not m:
        raise SystemExit("table %s not found" % name)
    b = re.sub(r"//[^\n]*", "", m.group(1))
    return b

def ints(text, name):
    b = body(text, name)
    b = b.replace("true", "1").replace("false", "0")
    b = re.sub(r"\b(lo|hi)\s*:", "", b)
    b = b.replace("OptimalEndpoint", "")
    return [int(x, 0) for x in re.findall(r"0x[0-9A-Fa-f]+|\d+", b)]

def main():
    u = (REF / "src/uastc.rs").read_text()
    a = (REF / "src/target_formats/astc.rs").read_text()
    b = (REF / "src/target_formats/bc7.rs").read_text()
    out = {}
    out["MODE_LUT"] = ints(u, "MODE_LUT")
    for n in ["PATTERNS_2", "PATTERNS_3", "PATTERNS_2_3", "PATTERNS_2_ANCHORS",
              "PATTERNS_3_ANCHORS", "PATTERNS_2_3_ANCHORS"]:
        out[n] = ints(u, n)
    for n in ["PATTERNS_2_ASTC_INDEX_10", "PATTERNS_3_ASTC_INDEX_10", "PATTERNS_2_3_ASTC_INDEX_10",
              "ASTC_QUINT_ENCODE_LUT", "ASTC_TRIT_ENCODE_LUT", "UASTC_TO_ASTC_BLOCK_MODE_13"]:
        out[n] = ints(a, n)
    for n in ["UASTC_TO_BC7_MODES", "PATTERNS_2_BC7_INDEX_INV", "PATTERNS_3_BC7_INDEX_PERM",
              "PATTERNS_3_BC7_TO_ASTC_PERMUTATIONS", "PATTERNS_2_3_BC7_INDEX_PERM",
              "PATTERNS_2_3_BC7_TO_ASTC_PERMUTATIONS", "PATTERNS_2_BC7", "PATTERNS_3_BC7",
              "PATTERNS_2_3_BC7", "PATTERNS_2_BC7_ANCHORS", "PATTERNS_3_BC7_ANCHORS",
              "BC7_MODE_5_OPTIMAL_ENDPOINTS", "BC7_MODE_6_OPTIMAL_ENDPOINTS"]:
        out[n] = ints(b, n)
    expect = {"MODE_LUT": 128, "PATTERNS_2": 480, "PATTERNS_3": 176, "PATTERNS_2_3": 304,
              "PATTERNS_2_ANCHORS": 60, "PATTERNS_3_ANCHORS": 33, "PATTERNS_2_3_ANCHORS": 38,
              "PATTERNS_2_ASTC_INDEX_10": 30, "PATTERNS_3_ASTC_INDEX_10": 11,
              "PATTERNS_2_3_ASTC_INDEX_10": 19, "ASTC_QUINT_ENCODE_LUT": 125,
              "ASTC_TRIT_ENCODE_LUT": 243, "UASTC_TO_ASTC_BLOCK_MODE_13": 20,
              "UASTC_TO_BC7_MODES": 20, "PATTERNS_2_BC7_INDEX_INV": 60,
              "PATTERNS_3_BC7_INDEX_PERM": 22, "PATTERNS_3_BC7_TO_ASTC_PERMUTATIONS": 18,
              "PATTERNS_2_3_BC7_INDEX_PERM": 38, "PATTERNS_2_3_BC7_TO_ASTC_PERMUTATIONS": 18,
              "PATTERNS_2_BC7": 480, "PATTERNS_3_BC7": 176, "PATTERNS_2_3_BC7": 304,
              "PATTERNS_2_BC7_ANCHORS": 128, "PATTERNS_3_BC7_ANCHORS": 192,
              "BC7_MODE_5_OPTIMAL_ENDPOINTS": 512, "BC7_MODE_6_OPTIMAL_ENDPOINTS": 514}
    for k, n in expect.items():
        assert len(out[k]) == n, (k, len(out[k]), n)
    dst = pathlib.Path(__file__).with_name("format_tables.json")
    dst.write_text(json.dumps(out, separators=(",", ":")) + "\n")
    print("wrote", dst, {k: len(v) for k, v in out.items()})

if __name__ == "__main__":
    main()
