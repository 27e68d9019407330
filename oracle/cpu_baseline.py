"""CPU baseline = the C restatement of the reference's algorithm (oracle/bls_oracle_c.c) timed on the host cores.

The reference is single-threaded: "all host cores" means T independent single-threaded
instances of verify_multiple_aggregate_signatures on disjoint chunks of the same workload, throughputs summed.
kind = "port": the Rust crate cannot be compiled in this image (no rustc/cargo), so this is the oracle port, built
with gcc -O3 and the field-layer variant measured fastest on the GPU box's host CPU (oracle/Makefile).  TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never on the product path.
"""
import ctypes
import os
import threading
import time

import numpy as np

from . import c_oracle

R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _synth(n_sets, n_keys, seed):
    """Valid sets of the benchmark's shape, synthesised on the CPU with a small key pool (cheap: 64 G1 mults)."""
    rs = np.random.RandomState(seed & 0x7fffffff)
    pool_n = max(64, n_keys)
    gen = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb).to_bytes(48, "big") + \
          (0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1).to_bytes(48, "big")
    sks = [int.from_bytes(rs.bytes(32), "big") % (R_ORDER - 1) + 1 for _ in range(pool_n)]
    pool = [c_oracle.g1_mul(gen, s) for s in sks]
    sigs, pks, msgs = [], [], []
    for j in range(n_sets):
        idx = rs.choice(pool_n, size=n_keys, replace=False)
        msg = rs.bytes(32)
        agg = sum(sks[i] for i in idx) % R_ORDER
        sigs.append(c_oracle.g2_mul(c_oracle.hash_to_g2(msg), agg))
        pks.append(b"".join(pool[i] for i in idx))
        msgs.append(msg)
    scalars = np.array([int.from_bytes(rs.bytes(8), "big") >> 1 or 1 for _ in range(n_sets)], dtype=np.uint64)
    return sigs, pks, msgs, scalars


_CACHE = {}


def run(n_sets, n_keys, seed=0xB200, threads=None):
    threads = threads or os.cpu_count() or 1
    threads = max(1, min(threads, n_sets))
    key = (n_sets, n_keys, seed)
    if key not in _CACHE:                       # input synthesis (signing side) is not part of the timed region
        _CACHE.clear()
        _CACHE[key] = _synth(n_sets, n_keys, seed)
    sigs, pks, msgs, scalars = _CACHE[key]
    chunks = [list(range(t, n_sets, threads)) for t in range(threads)]
    results = [None] * threads

    def work(t):
        ids = chunks[t]
        s = b"".join(sigs[i] for i in ids)
        k = b"".join(pks[i] for i in ids)
        m = b"".join(msgs[i] for i in ids)
        koff = np.arange(0, len(ids) * n_keys + 1, n_keys, dtype=np.uint32)
        moff = np.arange(0, len(ids) * 32 + 1, 32, dtype=np.uint32)
        results[t] = c_oracle.verify_multiple(s, k, koff, m, moff, scalars[ids])[0]     # ctypes releases the GIL

    c_oracle.lib()
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    sec = time.perf_counter() - t0
    assert all(results), "CPU baseline: the valid synthetic batch must verify"
    return {"sets": n_sets, "seconds": sec, "threads": threads, "kind": "port"}


def run_hash_to_g2(n_msgs, threads=None, seed=0xB200):
    """Second headline metric on the CPU: hash_to_curve_g2 (A/bls381/core.rs:831-849) of n_msgs distinct 32-byte messages,
    T independent single-threaded instances on disjoint chunks."""
    threads = threads or os.cpu_count() or 1
    threads = max(1, min(threads, n_msgs))
    rs = np.random.RandomState(seed & 0x7fffffff)
    msgs = [rs.bytes(32) for _ in range(n_msgs)]
    L = c_oracle.lib()
    out = [np.zeros(192, dtype=np.uint8) for _ in range(threads)]

    def work(t):
        o = out[t]
        for i in range(t, n_msgs, threads):
            m = np.frombuffer(msgs[i], dtype=np.uint8)
            L.oc_hash_to_g2(c_oracle._p(m), ctypes.c_size_t(32), None, ctypes.c_size_t(0), c_oracle._p(o))      # releases the GIL

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    sec = time.perf_counter() - t0
    assert out[0].any()
    return {"msgs": n_msgs, "seconds": sec, "threads": threads, "kind": "port"}


if __name__ == "__main__":
    import json
    import sys
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    print(json.dumps(run(n, 128)))
