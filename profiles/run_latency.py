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
                    d = lane.d
                    ok, fb = lane.eng.verify_multiple_dev(table if keys == "table" else None, d["sigs"].data_ptr(), d["idx" if keys == "table" else "pks"].data_ptr(),
                                                          d["pk_off"].data_ptr(), d["msgs"].data_ptr(), d["msg_off"].data_ptr(), d["scal"].data_ptr(), n)
                    st = dict(lane.eng.stage_ms())
                    st2 = {}
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
# the whole call on HOST pointers (pinned): b3_verify_multiple_indexed, H2D + D2H inside, wall clock around the synchronous call
import time
h = lane.host("pinned")
for mode, name in ((2, "plain"), (1, "replicated"), (0, "auto")):
    lane.eng.set_latency_mode(mode)
    best = None
    for r in range(reps + 1):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ok, fb, gt = lane.eng.verify_multiple_indexed(table, h["sigs"], h["idx"], h["pk_off"], h["msgs"], h["msg_off"], h["scal"], want_gt=True)
        ms = (time.perf_counter() - t0) * 1e3
        assert ok and fb == -1
        if r and (best is None or ms < best):
            best, st_best = ms, dict(lane.eng.stage_ms())
    print(f"whole call, host pointers, keys=table chain kernels={name:10s}: {best:.3f} ms/call = {n / best * 1e3:,.0f} sets/s  " +
          " ".join(f"{k}={v:.3f}" for k, v in st_best.items() if v), flush=True)
lane.eng.set_latency_mode(0)
lane.close()
table.close()
eng.close()
