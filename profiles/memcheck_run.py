"""Small driver for `compute-sanitizer --tool memcheck`: touches every kernel family of the path once at small sizes
(bucket-method MSM path, the three per-item finishing kernels, decompression, aggregation, hash_to_G2).
usage: compute-sanitizer --tool memcheck python profiles/memcheck_run.py"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import milagro_bls_b200 as mb
from milagro_bls_b200 import _lib

eng = mb.Engine(0)
rnd = random.Random(7)
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
n = 520                                                   # >= 512: bucket-method MSM, several accumulation groups
sks = [rnd.randrange(1, R) for _ in range(n)]
pk = eng.g1_mul_gen(sks)
msgs = [rnd.getrandbits(256).to_bytes(32, "big") for _ in range(n)]
sig = eng.g2_mul(eng.hash_to_g2(msgs).reshape(-1), sks)
scal = np.array([rnd.randrange(1, 1 << 63) for _ in range(n)], dtype=np.uint64)
moff = list(range(0, 32 * n + 1, 32))
ok, fb = eng.verify_multiple(sig.reshape(-1), pk.reshape(-1), None, b"".join(msgs), moff, scal)
assert ok and fb == -1
offs = list(range(0, n + 1, 4))                           # 130 sets of 4 keys: aggregation path
ok2, _ = eng.verify_multiple(sig[:130].reshape(-1), pk.reshape(-1), offs, b"".join(msgs[:130]), moff[:131], scal[:130])
assert not ok2                                            # keys do not match: still a full pass over every kernel
# round 2: both forms of the chain kernels (plain / replicated lanes), whole and partial calls, ragged key sets through the
# staged aggregation kernel (1, 5 and 130 keys per set), the key table, the two-phase call, batch normalisation
import torch
for mode in (2, 1):
    eng.set_latency_mode(mode)
    ok, fb = eng.verify_multiple(sig.reshape(-1), pk.reshape(-1), None, b"".join(msgs), moff, scal)
    assert ok and fb == -1
    part = torch.zeros(592, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    eng.verify_multiple_partial(sig.reshape(-1), pk.reshape(-1), None, b"".join(msgs), moff, scal, 0, part.data_ptr())
    assert eng.combine_partials_dev(part.data_ptr(), 1) == (True, -1)
eng.set_latency_mode(0)
rag = [0, 1, 6, 136, 136, 141]                            # set sizes 1, 5, 130, 0, 5
st_keys = pk[:141].reshape(-1)
a1, s1 = eng.g1_aggregate(st_keys, rag)
assert list(s1) == [0, 0, 0, -1, 0]
c48all, st = eng.g1_compress(pk.reshape(-1)); assert not st.any()
tbl = mb.KeyTable(eng)
first, st = tbl.append(c48all, compressed=True, validate=True); assert first == 0 and not st.any()
a2, s2 = eng.g1_aggregate_indexed(tbl, np.arange(141, dtype=np.uint32), rag)
assert list(s2) == [0, 0, 0, -1, 0] and a1.tobytes() == a2.tobytes()
idx = np.arange(n, dtype=np.uint32)
assert eng.verify_multiple_indexed(tbl, sig.reshape(-1), idx, None, b"".join(msgs), moff, scal) == (True, -1)
assert eng.sig_precheck(sig.reshape(-1)) == -1
assert eng.verify_multiple_checked(pk.reshape(-1), None, b"".join(msgs), moff, scal)
H = eng.hash_to_g2(msgs[:40])                               # Montgomery-trick normalisation, partial last chunk
tbl.close()
for kern in (1, 3):
    eng.set_item_kernel(kern)
    acc, st, gt = eng.verify_batch(_lib.ITEM_PRE_AGGREGATED, sig[:9].reshape(-1), pk[:9].reshape(-1), None, msgs[:9], want_gt=True)
    assert acc.all() and not st.any()
c48, st = eng.g1_compress(pk[:64].reshape(-1)); assert not st.any()
back, st = eng.g1_decompress(c48, validate=True); assert not st.any() and back.tobytes() == pk[:64].tobytes()
c96, st = eng.g2_compress(sig[:64].reshape(-1)); assert not st.any()
back, st = eng.g2_decompress(c96); assert not st.any() and back.tobytes() == sig[:64].tobytes()
agg, st = eng.g2_aggregate(sig[:64].reshape(-1), [0, 10, 64]); assert not st.any()
print("memcheck driver ok")
