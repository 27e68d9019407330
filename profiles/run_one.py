"""Short driver for ncu captures: synthesises the C4 batch (8192 sets x 128 keys) and runs `steps` resident
verify_multiple passes (python profiles/run_one.py [steps] [sets])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import milagro_bls_b200 as mb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
eng = mb.Engine(0)
bench.POOL = 16384                       # a small validator pool is enough for kernel captures
sks, pool = bench.synth_pool(eng, 0xB200)
lane = bench.Lane(0, dev, None, sks, pool, n, 128, 0xB200, 0, 0, 0)
part = torch.zeros(mb._lib.PARTIAL_BYTES, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for _ in range(steps):
    lane.partial_dev(n, 0, part.data_ptr())
    ok, fb = eng.combine_partials_dev(part.data_ptr(), 1)
    assert ok and fb == -1
print("ok", steps, "steps")
lane.close()
eng.close()
