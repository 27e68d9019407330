"""Pins the C oracle (oracle/bls_oracle_c.c: checker at large sizes + CPU baseline) to the reference's golden vectors
and, bit for bit, to the Python oracle.  CPU only."""
import json
import os
import random

import numpy as np

from oracle import bls_oracle as O
from oracle import c_oracle as C

G = os.path.join(os.path.dirname(__file__), "golden")


def g1w(P):
    return O.serialize_uncompressed_g1(P)


def g2w(P):
    return O.serialize_uncompressed_g2(P)


def test_c_h2c_golden_vectors():
    vec = json.load(open(os.path.join(G, "h2c_g2_ro.json")))
    for v in vec["vectors"]:
        P = ((int(v["P"]["x"][0], 16), int(v["P"]["x"][1], 16)), (int(v["P"]["y"][0], 16), int(v["P"]["y"][1], 16)))
        assert C.hash_to_g2(v["msg"].encode(), vec["dst"].encode()) == g2w(P)
    for m in [b"", b"cats", b"x" * 100]:
        assert C.hash_to_g2(m) == g2w(O.hash_to_curve_g2(m))


def test_c_group_ops_and_subgroup():
    rnd = random.Random(2)
    pts = [O.g1_mul(O.G1_GEN, rnd.randrange(1, O.r)) for _ in range(5)]
    rc, out = C.g1_aggregate(b"".join(g1w(P) for P in pts + [pts[0], O.g1_neg(pts[1]), None]))
    assert rc == 0 and out == g1w(O.aggregate_public_keys(pts + [pts[0], O.g1_neg(pts[1]), None]))
    assert C.g1_aggregate(b"")[0] == -1
    k = rnd.randrange(1, O.r)
    assert C.g1_mul(g1w(O.G1_GEN), k) == g1w(O.g1_mul(O.G1_GEN, k))
    assert C.g2_mul(g2w(O.G2_GEN), k) == g2w(O.g2_mul(O.G2_GEN, k))
    assert C.subgroup_check_g2(g2w(O.G2_GEN)) and C.subgroup_check_g2(g2w(None))
    assert not C.subgroup_check_g2(g2w(O.map_to_curve_g2((5, 7))))
    assert C.subgroup_check_g1(g1w(O.G1_GEN)) and not C.subgroup_check_g1(g1w((0, 2)))


def test_c_pairing_gt_bytes():
    one, gt = C.pairing(g2w(O.G2_GEN), g1w(O.G1_GEN))
    assert not one and gt == O.f12_to_bytes(O.fexp(O.ate2(O.G2_GEN, O.G1_GEN, None, None)))
    assert gt[:48].hex() == "1250ebd871fc0a92a7b2d83168d0d727272d441befa15c503dd8e90ce98db3e7b6d194f60839c508a84305aaca1789b6"


def test_c_verify_functions_match_python_oracle():
    sets = []
    for j in range(3):
        sks = [7000 + 100 * j + i for i in range(3)]
        pks = [O.sk_to_pk(s) for s in sks]
        msg = bytes([j]) * 32
        sets.append((O.aggregate_signatures([O.sign(s, msg) for s in sks]), pks, msg))
    scal = [0x7fffffffffffffff, 12345, 1]
    it = iter(scal)
    ok_o, gt_o = O.verify_multiple_aggregate_signatures(lambda n: next(it).to_bytes(8, "big"),
                                                        [(s, O.aggregate_public_keys(p), m) for s, p, m in sets], want_gt=True)
    sigs = b"".join(g2w(s) for s, _, _ in sets)
    pks = b"".join(g1w(P) for _, p, _ in sets for P in p)
    msgs = b"".join(m for _, _, m in sets)
    ok, gt = C.verify_multiple(sigs, pks, [0, 3, 6, 9], msgs, [0, 32, 64, 96], np.array(scal, dtype=np.uint64))
    assert ok and ok_o and gt == O.f12_to_bytes(gt_o)
    apks = b"".join(g1w(O.aggregate_public_keys(p)) for _, p, _ in sets)
    ok, gt2 = C.verify_multiple(sigs, apks, None, msgs, [0, 32, 64, 96], np.array(scal, dtype=np.uint64))
    assert ok and gt2 == gt
    bad = bytearray(msgs); bad[40] ^= 1
    it = iter(scal)
    ok_o, gt_o = O.verify_multiple_aggregate_signatures(
        lambda n: next(it).to_bytes(8, "big"),
        [(s, O.aggregate_public_keys(p), bytes(bad[32 * j:32 * j + 32])) for j, (s, p, _) in enumerate(sets)], want_gt=True)
    ok, gt = C.verify_multiple(sigs, pks, [0, 3, 6, 9], bytes(bad), [0, 32, 64, 96], np.array(scal, dtype=np.uint64))
    assert not ok and not ok_o and gt == O.f12_to_bytes(gt_o)
    # non-subgroup signature rejects
    q = g2w(O.map_to_curve_g2((5, 7)))
    ok, _ = C.verify_multiple(q + sigs[192:], pks, [0, 3, 6, 9], msgs, [0, 32, 64, 96], np.array(scal, dtype=np.uint64))
    assert not ok
    # fast_aggregate_verify / aggregate_verify
    s0, p0, m0 = sets[0]
    ok, gt = C.fast_aggregate_verify(g2w(s0), b"".join(g1w(P) for P in p0), m0)
    assert ok
    ok, gt = C.fast_aggregate_verify(g2w(s0), b"".join(g1w(P) for P in p0[:2]), m0)
    ok_o, gt_o = O.fast_aggregate_verify(s0, m0, p0[:2], want_gt=True)
    assert ok == ok_o and gt == O.f12_to_bytes(gt_o)
    sks = [9001, 9002, 9003]
    msgs3 = [b"a" * 32, b"b" * 32, b"c" * 32]
    agg = O.aggregate_signatures([O.sign(s, m) for s, m in zip(sks, msgs3)])
    pk3 = [O.sk_to_pk(s) for s in sks]
    ok, gt = C.aggregate_verify(g2w(agg), b"".join(g1w(P) for P in pk3), msgs3)
    assert ok
    ok, gt = C.aggregate_verify(g2w(agg), b"".join(g1w(P) for P in pk3), msgs3[::-1])
    ok_o, gt_o = O.aggregate_verify(agg, msgs3[::-1], pk3, want_gt=True)
    assert ok == ok_o and gt == O.f12_to_bytes(gt_o)


def test_cpu_baseline_runs():
    from oracle import cpu_baseline
    r = cpu_baseline.run(2, 4, seed=1, threads=2)
    assert r["sets"] == 2 and r["kind"] == "port" and r["seconds"] > 0


def test_c_field_layer_variants_agree(tmp_path):
    """The macro-gated field-layer variants of the C oracle (-DORACLE_SQR_DEDICATED: dedicated squaring; -DORACLE_FP2_LAZY:
    lazily reduced Fp2 product; the Makefile builds the combination measured fastest) compute the same bytes: hash_to_G2 of
    the golden messages and a rejecting 3-set verify_multiple GT, for all four combinations."""
    import ctypes
    import subprocess
    src = os.path.join(os.path.dirname(C.HERE), "oracle", "bls_oracle_c.c")
    vec = json.load(open(os.path.join(G, "h2c_g2_ro.json")))
    sks = [11, 22, 33]
    msgs = [b"m0", b"m1-longer", b""]
    sig = b"".join(g2w(O.sign(s, m)) for s, m in zip(sks, msgs))
    pk = b"".join(g1w(O.sk_to_pk(s)) for s in sks)
    blob = b"".join(msgs[::-1])                                     # messages in the wrong order: GT != 1
    moff = np.array([0, 0, 9, 11], dtype=np.uint32)
    scal = np.array([3, 5, 7], dtype=np.uint64)
    outs = []
    for k, defs in enumerate(([], ["-DORACLE_SQR_DEDICATED"], ["-DORACLE_FP2_LAZY"], ["-DORACLE_SQR_DEDICATED", "-DORACLE_FP2_LAZY"])):
        so = str(tmp_path / f"liboc_{k}.so")
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-Wno-unused-function", "-shared", "-o", so, src] + defs)
        L = ctypes.CDLL(so)
        h = []
        for v in vec["vectors"]:
            out = np.zeros(192, dtype=np.uint8)
            m, d = np.frombuffer(v["msg"].encode() or b"\0", dtype=np.uint8), np.frombuffer(vec["dst"].encode(), dtype=np.uint8)
            L.oc_hash_to_g2(ctypes.c_void_p(m.ctypes.data), ctypes.c_size_t(len(v["msg"])), ctypes.c_void_p(d.ctypes.data), ctypes.c_size_t(len(d)),
                            ctypes.c_void_p(out.ctypes.data))
            P = ((int(v["P"]["x"][0], 16), int(v["P"]["x"][1], 16)), (int(v["P"]["y"][0], 16), int(v["P"]["y"][1], 16)))
            assert out.tobytes() == g2w(P)
            h.append(out.tobytes())
        gt = np.zeros(576, dtype=np.uint8)
        a = [np.frombuffer(x, dtype=np.uint8) for x in (sig, pk, blob)]
        rc = L.oc_verify_multiple(ctypes.c_void_p(a[0].ctypes.data), ctypes.c_void_p(a[1].ctypes.data), None, ctypes.c_void_p(a[2].ctypes.data),
                                  ctypes.c_void_p(moff.ctypes.data), ctypes.c_void_p(scal.ctypes.data), ctypes.c_size_t(3), ctypes.c_void_p(gt.ctypes.data))
        outs.append((h, rc, gt.tobytes()))
    assert all(o == outs[0] for o in outs[1:])
    assert outs[0][2] != O.f12_to_bytes(O.F12_ONE)
