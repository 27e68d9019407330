"""Timing helper for b3_verify_batch (per-item accept bits): n items, device-resident inputs, CUDA-event time of the call.
usage: python profiles/run_batch.py [n_items] [keys_per_item] [reps] [item_kernel: 0 auto, 1 CTA per item, 2 thread per item, 3 lane pair per item]"""
import os
import sys
import random

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import milagro_bls_b200 as mb
from milagro_bls_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
eng = mb.Engine(0)
eng.set_item_kernel(int(sys.argv[4]) if len(sys.argv) > 4 else 0)
rnd = random.Random(1)
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
pool = 4096
sks = [rnd.randrange(1, R) for _ in range(pool)]
pk_pool = eng.g1_mul_gen(sks)
msgs = [rnd.randbytes(32) for _ in range(n)]
H = eng.hash_to_g2(msgs)
idx = np.array([[rnd.randrange(pool) for _ in range(k)] for _ in range(n)])
agg = [sum(sks[j] for j in row) % R for row in idx]
sig = eng.g2_mul(H.reshape(-1), agg)
pks = pk_pool[idx.reshape(-1)].reshape(-1)
mode = _lib.ITEM_FAST_AGGREGATE if k > 1 else _lib.ITEM_VERIFY
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_sig, d_pk = t(sig.reshape(-1)), t(pks)
d_off = t(np.arange(0, n * k + 1, k, dtype=np.uint32).view(np.int32))
d_msg = t(np.frombuffer(b"".join(msgs), dtype=np.uint8))
d_moff = t(np.arange(0, 32 * n + 1, 32, dtype=np.uint32).view(np.int32))
d_acc = torch.zeros(n, dtype=torch.int32, device=dev)
d_st = torch.zeros(n, dtype=torch.int32, device=dev)
for r in range(reps + 1):
    eng.verify_batch_dev(mode, d_sig.data_ptr(), d_pk.data_ptr(), d_off.data_ptr() if k > 1 else None, d_msg.data_ptr(), d_moff.data_ptr(), n,
                         d_acc.data_ptr(), d_st.data_ptr())
    ms = eng.last_kernel_ms(0)
    print(f"n={n} keys/item={k}: {ms:.3f} ms  -> {n / ms * 1e3:,.0f} items/s   items kernel {eng.last_kernel_ms(1):.3f} ms  accepted {int(d_acc.sum())}/{n}", flush=True)
print({s: round(v, 3) for s, v in eng.stage_ms().items() if v})
