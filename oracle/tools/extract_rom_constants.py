#!/usr/bin/env python3
"""Dev-time tool (runs only in the build container, never on the GPU box).

Decodes the base-2^58 limb arrays of the reference ROM
(/root/reference/incubator-milagro-crypto-rust/src/roms/rom_bls381_64.rs:28-223 and
 .../src/bls381/iso_constants_x64.rs:5-188) into plain integers and writes
oracle/rom_constants.py.  Only *numbers* (curve parameters published in RFC 9380 / the
BLS12-381 spec) are extracted; no reference code is copied.
"""
import re, sys, pathlib

A = pathlib.Path("/root/reference/incubator-milagro-crypto-rust/src")
BASEBITS = 58


def parse_consts(path):
    txt = path.read_text()
    out = {}
    # 1-D arrays:  pub const NAME: [Chunk; NLEN] = [ ... ];
    for m in re.finditer(r"pub const (\w+): \[Chunk; NLEN\] = \[(.*?)\];", txt, re.S):
        limbs = [int(v, 16) for v in re.findall(r"0x[0-9A-Fa-f]+", m.group(2))]
        out[m.group(1)] = sum(l << (BASEBITS * i) for i, l in enumerate(limbs))
    # 2-D arrays:  pub const NAME: [[Chunk; NLEN]; K] = [ [..], [..] ];
    for m in re.finditer(r"pub const (\w+): \[\[Chunk; NLEN\]; (\d+)\] = \[(.*?)\n\];", txt, re.S):
        rows = re.findall(r"\[([^\[\]]*?)\]", m.group(3), re.S)
        vals = []
        for row in rows:
            limbs = [int(v, 16) for v in re.findall(r"0x[0-9A-Fa-f]+", row)]
            vals.append(sum(l << (BASEBITS * i) for i, l in enumerate(limbs)))
        if len(vals) == int(m.group(2)):   # (unused BN-curve tables have a different shape; skip)
            out[m.group(1)] = vals
    return out


rom = parse_consts(A / "roms/rom_bls381_64.rs")
iso = parse_consts(A / "bls381/iso_constants_x64.rs")
names = ["MODULUS", "FRA", "FRB", "CURVE_ORDER", "CURVE_GX", "CURVE_GY", "CURVE_BNX", "CURVE_CRU",
         "CURVE_PXA", "CURVE_PXB", "CURVE_PYA", "CURVE_PYB",
         "SSWU_A2_A", "SSWU_A2_B", "SSWU_B2_A", "SSWU_B2_B", "SSWU_Z2_A", "SSWU_Z2_B"]
lines = ['"""Numbers decoded from the reference ROM by oracle/tools/extract_rom_constants.py (do not edit).',
         'Sources: A/roms/rom_bls381_64.rs:28-223, A/bls381/iso_constants_x64.rs:5-188 (A = amcl src dir)."""', ""]
for n in names:
    lines.append(f"{n} = 0x{rom[n]:x}")
for n in ["ISO3_XNUM", "ISO3_XDEN", "ISO3_YNUM", "ISO3_YDEN"]:
    v = iso[n]
    pairs = [(v[2 * i], v[2 * i + 1]) for i in range(len(v) // 2)]
    lines.append(f"{n} = [" + ", ".join(f"(0x{a:x}, 0x{b:x})" for a, b in pairs) + "]")
pathlib.Path(__file__).resolve().parents[1].joinpath("rom_constants.py").write_text("\n".join(lines) + "\n")
print("ok", {k: (hex(v) if isinstance(v, int) else len(v)) for k, v in rom.items() if k in names})
