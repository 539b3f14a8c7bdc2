#!/usr/bin/env python3
"""Dev-time generator for tests/golden/uastc_kat.bin.

Converts the reference's known-answer vectors (tests/block_test_cases/uastc_{rgba,astc,bc7,
etc1,etc2}.rs, driven by tests/transcode_uastc_block.rs:35-78; 19 modes x 32 blocks x 5
targets) into one neutral binary fixture.  The vectors are data (outputs of the upstream
Basis Universal transcoder captured by the reference's tests/test_block_export.rs); licence
MIT OR Apache-2.0 (reference Cargo.toml:8).  /root/reference is needed only to re-run this
script; tests read the committed .bin.

Layout (little endian):  magic "UKAT" | u32 count | count records of
  u8 mode | u8[16] uastc | u32[16] rgba (0xAABBGGRR) | u8[16] astc | u8[16] bc7 | u8[8] etc1 | u8[16] etc2
"""
import re, struct, sys, pathlib

REF = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
D = REF / "tests/block_test_cases"

def parse(name):
    text = (D / name).read_text()
    text = text[text.index("= ["):]
    groups = re.split(r"&\[\s*//\s*\d+", text)[1:]
    assert len(groups) == 19, len(groups)
    out = []
    for mode, g in enumerate(groups):
        for m in re.finditer(r"\(\[([^\]]*)\],\s*\[([^\]]*)\]\)", g):
            a = [int(x, 0) for x in m.group(1).replace(" ", "").split(",") if x]
            b = [int(x, 0) for x in m.group(2).replace(" ", "").split(",") if x]
            assert len(a) == 16
            out.append((mode, bytes(a), b))
    assert len(out) == 608, len(out)
    return out

def main():
    rgba = parse("uastc_rgba.rs"); astc = parse("uastc_astc.rs"); bc7 = parse("uastc_bc7.rs")
    etc1 = parse("uastc_etc1.rs"); etc2 = parse("uastc_etc2.rs")
    blob = bytearray(b"UKAT" + struct.pack("<I", 608))
    for i in range(608):
        mode, inp, px = rgba[i]
        for other in (astc, bc7, etc1, etc2):
            assert other[i][0] == mode and other[i][1] == inp
        assert len(px) == 16 and len(astc[i][2]) == 16 and len(bc7[i][2]) == 16
        assert len(etc1[i][2]) == 8 and len(etc2[i][2]) == 16
        blob += bytes([mode]) + inp + struct.pack("<16I", *px)
        blob += bytes(astc[i][2]) + bytes(bc7[i][2]) + bytes(etc1[i][2]) + bytes(etc2[i][2])
    dst = pathlib.Path(__file__).resolve().parent.parent / "tests/golden/uastc_kat.bin"
    dst.write_bytes(blob)
    print("wrote", dst, len(blob), "bytes")

if __name__ == "__main__":
    main()
