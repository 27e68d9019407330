"""Headline region (resident, key table) for several numbers of calls in flight per GPU.  usage: python profiles/sweep_inflight.py"""
import os
import subprocess
import sys
import re

for f in [int(x) for x in sys.argv[1:]] or (6, 8, 10, 12, 16):
    env = dict(os.environ, B3_BENCH_DIAG="1", B3_BENCH_PRINT_HEADLINE="1")
    out = subprocess.run([sys.executable, "bench.py", "--steps", "4", "--warmup", "3", "--inflight", str(f)], env=env, capture_output=True, text=True)
    for line in (out.stdout + out.stderr).splitlines():
        if line.startswith("headline_region"):
            print("inflight", f, line, flush=True)
