import os, sys, subprocess
for k in (16, 24, 32, 48, 64):
    env = dict(os.environ, B3_ACC_K=str(k))
    out = subprocess.run([sys.executable, "profiles/run_latency.py", "3"], env=env, capture_output=True, text=True).stdout
    for line in out.splitlines():
        if "keys=table chain kernels=replicated" in line and ("serialised" in line or "overlapped" in line):
            acc = [t for t in line.split() if t.startswith("miller_accumulate=")][0]
            print("K =", k, line.split(":")[0][-12:], line.split(":")[1].split("=")[0], acc, flush=True)
