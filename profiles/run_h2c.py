"""hash_to_G2 alone: device-resident 32-byte messages, CUDA-event time of b3_hash_to_g2_dev per batch size and latency mode.
usage: python profiles/run_h2c.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import milagro_bls_b200 as mb

dev = torch.device("cuda", 0)
eng = mb.Engine(0)
peak = eng.imad_peak(wide=True)
for n in (8192, 65536):
    hm = torch.from_numpy(np.random.RandomState(5).randint(0, 256, size=n * 32, dtype=np.uint8)).to(dev)
    ho = torch.arange(0, n * 32 + 1, 32, dtype=torch.int32, device=dev)
    out = torch.empty(n * 192, dtype=torch.uint8, device=dev)
    for mode, name in ((2, "plain"),):
        eng.set_latency_mode(mode)
        best = None
        for r in range(5):
            eng.hash_to_g2_dev(hm.data_ptr(), ho.data_ptr(), n, out.data_ptr())
            st = eng.stage_ms()
            ms = st["hash_to_g2_affine"]
            best = ms if best is None or ms < best else best
        print(f"n={n:6d}: hash_to_curve_g2 + normalisation {best:.3f} ms = {n / best * 1e3:,.0f} hash_to_G2/s, "
              f"{n * 6700 * 300 / (best * 1e-3) / peak:.3f} of the IMAD.WIDE peak on 6700 M, {n * 1.255e6 / (best * 1e-3) / peak:.3f} on executed MACs")
