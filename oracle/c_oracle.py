"""ctypes wrapper of the C oracle (oracle/bls_oracle_c.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libbls_oracle_c.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = [os.path.join(HERE, f) for f in ("bls_oracle_c.c", "ec_generic.inc", "bls_oracle_consts.h")]
        if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = ctypes.CDLL(LIB)
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def _u8(b):
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8) if len(b) else np.zeros(1, dtype=np.uint8)


def hash_to_g2(msg, dst=None):
    out = np.zeros(192, dtype=np.uint8)
    m = _u8(msg)
    if dst is None:
        lib().oc_hash_to_g2(_p(m), ctypes.c_size_t(len(msg)), None, ctypes.c_size_t(0), _p(out))
    else:
        d = _u8(dst)
        lib().oc_hash_to_g2(_p(m), ctypes.c_size_t(len(msg)), _p(d), ctypes.c_size_t(len(dst)), _p(out))
    return out.tobytes()


def g1_aggregate(pks96):
    out = np.zeros(96, dtype=np.uint8)
    a = _u8(pks96)
    rc = lib().oc_g1_aggregate(_p(a), ctypes.c_size_t(len(pks96) // 96), _p(out))
    return rc, out.tobytes()


def subgroup_check_g2(p192):
    a = _u8(p192)                      # (keep every buffer referenced until the call returns)
    return bool(lib().oc_subgroup_check_g2(_p(a)))


def subgroup_check_g1(p96):
    a = _u8(p96)
    return bool(lib().oc_subgroup_check_g1(_p(a)))


def g1_mul(p96, k):
    out = np.zeros(96, dtype=np.uint8)
    a, kb = _u8(p96), np.frombuffer(int(k).to_bytes(32, "big"), dtype=np.uint8).copy()
    lib().oc_g1_mul(_p(a), _p(kb), _p(out))
    return out.tobytes()


def g2_mul(p192, k):
    out = np.zeros(192, dtype=np.uint8)
    a, kb = _u8(p192), np.frombuffer(int(k).to_bytes(32, "big"), dtype=np.uint8).copy()
    lib().oc_g2_mul(_p(a), _p(kb), _p(out))
    return out.tobytes()


def pairing(q192, p96):
    gt = np.zeros(576, dtype=np.uint8)
    a, b = _u8(q192), _u8(p96)
    one = lib().oc_pairing(_p(a), _p(b), _p(gt))
    return bool(one), gt.tobytes()


def verify_multiple(sigs192, pks96, pk_off, msgs, msg_off, scalars):
    """Returns (accept, gt bytes).  pk_off None -> one (aggregate) key per set."""
    s, k, m = _u8(sigs192), _u8(pks96), _u8(msgs)
    mo = np.ascontiguousarray(msg_off, dtype=np.uint32)
    sc = np.ascontiguousarray(scalars, dtype=np.uint64)
    gt = np.zeros(576, dtype=np.uint8)
    if pk_off is None:
        po = None
    else:
        ko = np.ascontiguousarray(pk_off, dtype=np.uint32)
        po = _p(ko)
    ok = lib().oc_verify_multiple(_p(s), _p(k), po, _p(m), _p(mo), _p(sc), ctypes.c_size_t(len(sc)), _p(gt))
    return bool(ok), gt.tobytes()


def fast_aggregate_verify(sig192, pks96, msg, reject_inf=True):
    gt = np.zeros(576, dtype=np.uint8)
    a, b, c = _u8(sig192), _u8(pks96), _u8(msg)
    ok = lib().oc_fast_aggregate_verify(_p(a), _p(b), ctypes.c_size_t(len(pks96) // 96), _p(c),
                                        ctypes.c_size_t(len(msg)), ctypes.c_int(1 if reject_inf else 0), _p(gt))
    return bool(ok), gt.tobytes()


def aggregate_verify(sig192, pks96, msgs):
    gt = np.zeros(576, dtype=np.uint8)
    off = np.zeros(len(msgs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(x) for x in msgs])
    a, b, c = _u8(sig192), _u8(pks96), np.frombuffer(b"".join(msgs), dtype=np.uint8).copy()
    ok = lib().oc_aggregate_verify(_p(a), _p(b), _p(c), _p(off), ctypes.c_size_t(len(msgs)), _p(gt))
    return bool(ok), gt.tobytes()
