"""One C4 call at a time (8192 sets x 128 keys resident in HBM), per latency mode: call time and per-stage spans.
usage: python profiles/run_latency.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import milagro_bls_b200 as mb

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
eng = mb.Engine(0)
bench.POOL = 65536
sks, pool = bench.synth_pool(eng, 0xB200)
c48 = eng.g1_compress(pool.reshape(-1))[0]
table = mb.KeyTable(eng, bench.POOL)
table.append(c48, compressed=True, validate=True)
lane = bench.Lane(0, dev, table, sks, pool, n, 128, 0xB200, 0, 0, 0)
part = torch.zeros(mb._lib.PARTIAL_BYTES, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for keys in ("table", "bytes"):
    for mode, name in ((2, "plain"), (1, "replicated")):
        lane.eng.set_latency_mode(mode)
        for serial in (False, True):
            lane.eng.set_serial(serial)
            best, st_best = None, None
            for r in range(reps + 1):
                flush.fill_(1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(lane.stream):
                    e0.record()
                    lane.partial_dev(n, 0, part.data_ptr(), keys)
                    st = dict(lane.eng.stage_ms())
                    ok, fb = lane.eng.combine_partials_dev(part.data_ptr(), 1)
                    st2 = lane.eng.stage_ms()
                    e1.record()
                torch.cuda.synchronize()
                assert ok and fb == -1
                for k, v in st2.items():
                    st[k] = st.get(k, 0) + v
                ms = e0.elapsed_time(e1)
                if r and (best is None or ms < best):
                    best, st_best = ms, st
            print(f"keys={keys:5s} chain kernels={name:10s} {'serialised' if serial else 'overlapped'}: {best:.3f} ms/call = {n / best * 1e3:,.0f} sets/s  " +
                  " ".join(f"{k}={v:.3f}" for k, v in st_best.items() if v), flush=True)
lane.eng.set_serial(False)
lane.close()
table.close()
eng.close()
