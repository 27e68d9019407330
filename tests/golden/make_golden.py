#!/usr/bin/env python3
"""Dev-time generator for the committed golden fixtures (runs in the build container only).

Reads the reference's OWN test vectors / known answers and rewrites them as small JSON fixtures:
  * A/test_utils/hash_to_curve_vectors/BLS12381G2_XMDSHA-256_SSWU_RO_.json  (consumed by the
    reference test A/bls381/core.rs:858-937)                       -> h2c_g2_ro.json
  * the fixed compressed G1/G2 points of M/src/amcl_utils.rs:83-144 (= A/bls381/core.rs:1188-1224),
    the README secret key M/src/signature.rs:105-108, the sk=1 / sk=r-1 pair of
    M/src/aggregates.rs:395-403 and the encoding edge cases of M/src/keys.rs:250-350
                                                                     -> known_points.json
  * oracle-derived cross-check values (NOT reference-pinned; flagged "derived")
                                                                     -> derived.json
A = /root/reference/incubator-milagro-crypto-rust/src, M = /root/reference/src.
"""
import json, pathlib, re, sys

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
A = pathlib.Path("/root/reference/incubator-milagro-crypto-rust/src")
M = pathlib.Path("/root/reference/src")

# ---- h2c vectors ---------------------------------------------------------------------------
src = json.loads((A / "test_utils/hash_to_curve_vectors/BLS12381G2_XMDSHA-256_SSWU_RO_.json").read_text())
out = {"source": "A/test_utils/hash_to_curve_vectors/BLS12381G2_XMDSHA-256_SSWU_RO_.json",
       "dst": src["dst"], "vectors": []}
for v in src["vectors"]:
    def pt(d):
        return {k: [c for c in d[k].split(",")] for k in ("x", "y")}
    out["vectors"].append({"msg": v["msg"], "u": [u.split(",") for u in v["u"]],
                           "Q0": pt(v["Q0"]), "Q1": pt(v["Q1"]), "P": pt(v["P"])})
(HERE / "h2c_g2_ro.json").write_text(json.dumps(out, indent=1) + "\n")

# ---- known compressed points ---------------------------------------------------------------
txt = (M / "amcl_utils.rs").read_text()
hexes = re.findall(r'hex::decode\("([0-9a-f]+)"\)', txt)
g1 = [h for h in hexes if len(h) == 96][:3]
g2_halves = [h for h in hexes if len(h) == 96][3:]
g2 = [g2_halves[2 * i] + g2_halves[2 * i + 1] for i in range(3)]
sig_txt = (M / "signature.rs").read_text()
m = re.search(r"let sk_bytes = vec!\[(.*?)\];", sig_txt, re.S)
readme_sk = bytes(int(x) for x in re.findall(r"\d+", m.group(1)))
agg_txt = (M / "aggregates.rs").read_text()
known = {
    "source": "M/src/amcl_utils.rs:83-144; M/src/signature.rs:105-108; M/src/keys.rs:250-350",
    "g1_compressed": g1, "g2_compressed": g2,
    "readme_sk": readme_sk.hex(), "readme_msg": "cats",
    "pk_not_in_subgroup_compressed": (bytes([128]) + bytes(47)).hex(),     # keys.rs:334-341, point (0,2)
    "pk_infinity_with_junk": (bytes([196]) + bytes(47)).hex(),             # keys.rs:344-350
    "pk_infinity": (bytes([192]) + bytes(47)).hex(),                       # keys.rs:250-259
    "uncompressed_bad_point_1_1": (bytes(47) + b"\x01" + bytes(47) + b"\x01").hex(),  # keys.rs:276-282
}
(HERE / "known_points.json").write_text(json.dumps(known, indent=1) + "\n")

# ---- derived (oracle) values ---------------------------------------------------------------
from oracle import bls_oracle as O
sk = int.from_bytes(readme_sk, "big")
pk = O.sk_to_pk(sk)
H = O.hash_to_curve_g2(b"cats")
sig = O.sign(sk, b"cats")
gt_gen = O.fexp(O.ate2(O.G2_GEN, O.G1_GEN, None, None))
ok, gt = O.signature_verify(sig, b"cats", pk, want_gt=True)
assert ok
derived = {
    "note": "derived with oracle/bls_oracle.py; NOT pinned by the reference (cross-checks only)",
    "readme_pk_compressed": O.serialize_g1(pk).hex(),
    "h_cats_compressed": O.serialize_g2(H).hex(),
    "readme_sig_cats_compressed": O.serialize_g2(sig).hex(),
    "gt_generator_bytes": O.f12_to_bytes(gt_gen).hex(),
}
(HERE / "derived.json").write_text(json.dumps(derived, indent=1) + "\n")
print("golden written", len(out["vectors"]), len(g1), len(g2))
