#!/usr/bin/env python3
"""bench.py -- verified signature-sets/s of verify_multiple_aggregate_signatures on B200.

Contract (see the task statement):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  * workload at N = 1: BASELINE.json configs[3] -- 8192 attestation sets x 128 public keys, distinct random
    32-byte messages, 63-bit batch scalars (C4).  For N > 1 every rank keeps 8192 sets (weak scaling), so N = 8
    is BASELINE.json configs[4] (2^16 sets, NCCL combine of the 592-byte partial Miller products).
  * one call = one full verification of a C4 batch: G2 subgroup checks, G1 key aggregation, [c]apk, hash_to_G2,
    [c]sig sum, n+8 Miller loops, Fp12 product, (all-gather,) one final exponentiation, accept bit.
  * a step = one such call on EACH of --inflight (default 8) contexts per GPU, running concurrently: one b3_ctx + one
    host thread per call, the reference's own threading model (re-entrant types, callers parallelise externally).  A
    single 8192-set call is a chain of latency-bound kernels (~7 ms end to end) that cannot fill 148 SMs by itself;
    K steps = K x inflight full verifications of 8192 sets each.
  * value  = sets verified per second, inputs resident in HBM (device-pointer C-ABI entry points).
             `one_batch_in_flight` reports the same metric with a single call at a time (call latency).
  * e2e    = same metric through the host-pointer C-ABI call (b3_verify_multiple) with pinned HOST buffers:
             H2D of the step's inputs and D2H of accept + GT inside the timed region.
  * roofline: bound = integer multiply pipe (IMAD); the peak is measured live with a pure IMAD.WIDE carry-chain
    probe (b3_imad_peak); HBM figures are reported next to it to show that memory is not binding.
  * cpu_baseline / --impl reference: the reference's CPU algorithm (oracle/ C restatement; the Rust crate cannot be
    built here -- no rustc) on all host cores, on a bounded sample of the same workload.
Only this file's cpu_baseline / --impl reference legs touch oracle/ -- never the measured GPU path.
"""
import argparse
import json
import os
# every context owns four CUDA streams; the default of 8 hardware work queues would alias the streams of concurrent
# batches onto shared queues (false dependencies between independent batches)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SETS_PER_GPU = 8192
KEYS_PER_SET = 128
POOL = 16384
MSG_LEN = 32
R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
# algorithmic work per unit: 32x32->64 multiply-accumulates, 300 per Fp multiplication
MACS_PER_FP_MUL = 300
FP_MULS = {"g1_aggregate": 1400, "g2_parse_subgroup_check": 1170, "hash_to_g2_affine": 6700, "g1_scalar_mul_affine": 670 + 15,
           "g2_scalar_mul_sum": 1650, "miller_lines": 1800, "miller_accumulate": 3000, "miller_chain": 0, "final_exp": 0}
FP_MULS_PER_SET = 16400
# what this implementation actually executes per set (inversion-free maps, bucket-method sum, split
# Miller loop), in the same unit -- reported beside the SURVEY figure so the fraction cannot flatter the kernels
EXEC_FP_MULS_PER_SET = 12900
# DRAM bytes (read + write) per launch at the C4 shape from the committed `ncu --set full` captures (profiles/r1s_h2c_full.txt, r1t_accum_full.txt)
NCU_TRAFFIC_BYTES = {"hash_to_g2_affine": 534016 + 6311424, "miller_accumulate": 162358272 + 5568000}
CPU_PASSES = 4                # timed passes of the CPU baseline over its 2048-set sample (~10 s of CPU work per 4 cores)
B3_EXTRA_PAIRS = 8            # window sums of the bucket-method signature sum, each its own pair
BYTES_PER_SET = KEYS_PER_SET * 96 + 192 + MSG_LEN + 8          # algorithmic HBM bytes read per set


def splitmix64(seed):
    x = seed & (2**64 - 1)
    while True:
        x = (x + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        yield z ^ (z >> 31)


def synth_inputs(eng, n_sets, n_keys, seed, rank=0):
    """Synthetic, VALID signature sets.  Signing-side work (pk = [sk]G1,
    sig = [sum sk]H(m)) runs on the GPU helpers b3_g1_mul_gen / b3_g2_mul; it is input synthesis, not timed."""
    import numpy as np
    g = splitmix64(seed)
    sks = [1 + ((next(g) | (next(g) << 64) | (next(g) << 128) | (next(g) << 192)) % (R_ORDER - 1)) for _ in range(POOL)]
    pool = eng.g1_mul_gen(sks)                                   # (POOL, 96)
    rs = np.random.RandomState((seed + 7919 * rank) & 0x7fffffff)
    idx = np.stack([rs.choice(POOL, size=n_keys, replace=False) for _ in range(n_sets)])      # (n_sets, n_keys)
    msgs = rs.randint(0, 256, size=(n_sets, MSG_LEN), dtype=np.uint8)
    msgs[:, 0] = rank
    msgs[:, 1:5] = np.arange(n_sets, dtype=">u4").view(np.uint8).reshape(n_sets, 4)            # all distinct
    sk_arr = sks
    agg = [sum(sk_arr[i] for i in row) % R_ORDER for row in idx]
    H = eng.hash_to_g2([m.tobytes() for m in msgs])
    sigs = eng.g2_mul(H.reshape(-1), agg)                        # (n_sets, 192)
    pks = pool[idx.reshape(-1)]                                  # (n_sets*n_keys, 96)
    pk_off = np.arange(0, n_sets * n_keys + 1, n_keys, dtype=np.uint32)
    msg_off = np.arange(0, n_sets * MSG_LEN + 1, MSG_LEN, dtype=np.uint32)
    return {"sigs": np.ascontiguousarray(sigs.reshape(-1)), "pks": np.ascontiguousarray(pks.reshape(-1)), "pk_off": pk_off,
            "msgs": np.ascontiguousarray(msgs.reshape(-1)), "msg_off": msg_off, "n": n_sets, "sks": sks, "idx": idx}


def draw_scalars(n, seed):
    import numpy as np
    from milagro_bls_b200 import SeededRng, draw_scalar
    rng = SeededRng(seed)
    return np.array([draw_scalar(rng) for _ in range(n)], dtype=np.uint64)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_run(sample_sets, n_keys, seed, threads=None):
    """Times the CPU restatement of the reference's algorithm (oracle/) on `sample_sets` sets of the same shape,
    T independent single-threaded instances on disjoint chunks (the reference has no threads).  Returns a dict."""
    from oracle import cpu_baseline
    return cpu_baseline.run(sample_sets, n_keys, seed, threads)


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np  # noqa: F401
    cores = os.cpu_count() or 1
    sample = max(cores, args.ref_sets)
    times, last = [], None
    for it in range(args.warmup + args.steps):
        last = cpu_reference_run(sample, KEYS_PER_SET, 0xB200, cores)
        if it >= args.warmup:
            times.append(last["seconds"])
    sec = sum(times) / max(len(times), 1)
    value = sample / sec
    line = {"impl": "reference", "metric": "verified sig-sets/s (verify_multiple_aggregate_signatures, 128 keys/set)", "value": value,
            "unit": "sets/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (6x64-bit Montgomery limbs, __int128)",
            "data": "synthetic",
            "config": {"workload": f"verify_multiple_aggregate_signatures: {SETS_PER_GPU} sets x {KEYS_PER_SET} keys per GPU (C4/C5)",
                       "sample_sets_per_step": sample, "keys_per_set": KEYS_PER_SET},
            "cpu_baseline": {"value": value, "unit": "sets/s", "cores": last["threads"], "kind": last["kind"],
                             "sample": f"{sample} sets x {KEYS_PER_SET} keys per step, {last['threads']} independent single-threaded instances"},
            "e2e": {"value": value, "unit": "sets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--sets", type=int, default=SETS_PER_GPU, help="sets per GPU (default: the C4 shape)")
    ap.add_argument("--keys", type=int, default=KEYS_PER_SET)
    ap.add_argument("--ref-sets", type=int, default=2048,
                    help="sets per step of the CPU reference arm / cpu_baseline sample (~20 s of CPU work on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=8,
                    help="verification batches in flight per GPU (one b3_ctx + host thread each; 1 = one call at a time)")
    ap.add_argument("--h2c-msgs", type=int, default=65536, help="messages per hash_to_G2 batch of the second metric")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the SURVEY 8(f) rows (decompression, aggregation, per-item verification)")
    ap.add_argument("--breakdown", action="store_true", help="print the per-stage device times to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import milagro_bls_b200 as mb
    from milagro_bls_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    eng = mb.Engine(local_rank)
    try:
        _bench(eng, args, world, rank, local_rank, dev)          # every tensor made on the library's stream dies in here
    finally:
        torch.cuda.synchronize()
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        eng.close()                                              # ... before the stream itself is destroyed
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


class Lane:
    """One verification context (b3_ctx = its own CUDA streams and scratch) with its own synthetic batch.  The reference's
    types are re-entrant and its callers parallelise externally (threading): concurrent batches on
    separate contexts are the drop-in equivalent, and they are what keeps the GPU full while one batch is in its serial
    tail (closing Miller chain, final exponentiation)."""

    def __init__(self, local_rank, dev, n, nk, seed, rank, tag):
        import numpy as np
        import torch
        import milagro_bls_b200 as mb
        self.eng = mb.Engine(local_rank)
        self.stream = torch.cuda.ExternalStream(int(self.eng.L.b3_ctx_stream(self.eng.handle)), device=dev)
        inp = synth_inputs(self.eng, n, nk, seed, rank)
        scal = draw_scalars(n, b"bench-%d-%d" % (rank, tag))
        keys = ("sigs", "pks", "pk_off", "msgs", "msg_off")
        with torch.cuda.stream(self.stream):
            self.d = {k: torch.from_numpy(inp[k]).to(dev) for k in keys}
            self.d["scal"] = torch.from_numpy(scal.view(np.int64)).to(dev)
            self.stream.synchronize()
        self.pin = {k: torch.from_numpy(inp[k]).pin_memory() for k in keys}
        self.pin["scal"] = torch.from_numpy(scal.view(np.int64)).pin_memory()
        self.h2d_bytes = sum(int(t.numel() * t.element_size()) for t in self.pin.values())

    def partial_dev(self, n, base, d_partial):
        d = self.d
        self.eng.verify_multiple_partial_dev(d["sigs"].data_ptr(), d["pks"].data_ptr(), d["pk_off"].data_ptr(), d["msgs"].data_ptr(),
                                             d["msg_off"].data_ptr(), d["scal"].data_ptr(), n, base, d_partial)

    def close(self):
        self.d = self.pin = None
        self.eng.close()


def _bench(eng, args, world, rank, local_rank, dev):
    import numpy as np
    import torch
    import torch.distributed as dist
    import milagro_bls_b200 as mb
    from milagro_bls_b200 import sharding
    n, nk = args.sets, args.keys
    S = max(1, args.inflight)
    # `eng` is the combining context (all-gathered partials -> product -> final exponentiation); every torch op of the
    # benchmark (L2 flush, NCCL all-gather, timing events) runs on ITS stream
    lib_stream = torch.cuda.ExternalStream(int(eng.L.b3_ctx_stream(eng.handle)), device=dev)
    torch.cuda.set_stream(lib_stream)
    lanes = [Lane(local_rank, dev, n, nk, 0xB200 + 0x101 * t, rank, t) for t in range(S)]
    try:
        _bench_lanes(eng, lanes, args, world, rank, local_rank, dev)
    finally:
        torch.cuda.synchronize()
        for ln in lanes:
            ln.close()


def _bench_next_rows(eng, lane, dev, n, nk):
    """batched decompression + validation (PublicKey::from_bytes / Signature::from_bytes), signature aggregation,
    per-item verification.  Host-pointer calls are timed by wall clock around the synchronous C-ABI call (H2D/D2H inside);
    the per-item verification runs on the lane's device-resident C4 batch and is timed by the library's CUDA events.
    Each row carries a size-independent parity property checked here (round trip / all items accept); bit-exact parity with
    the oracle is in tests/test_gpu_parity.py."""
    import numpy as np
    import torch
    from milagro_bls_b200 import _lib
    peak = eng.imad_peak(wide=True)
    out = {}

    def wall(fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        return (time.perf_counter() - t0) / reps, r

    def row(name, units, secs, fp_muls, extra=None):
        d = {"value": units / secs, "unit": name.split(":")[1], "units_per_call": units, "ms_per_call": secs * 1e3,
             "algorithmic_fp_muls_per_unit": fp_muls, "imad_frac": units / secs * fp_muls * MACS_PER_FP_MUL / peak}
        d.update(extra or {})
        out[name.split(":")[0]] = d

    keys96 = lane.pin["pks"].numpy()[:96 * 65536]
    nkeys = len(keys96) // 96
    c48, st = eng.g1_compress(keys96)
    assert not st.any()
    secs, (back, st) = wall(lambda: eng.g1_decompress(c48, validate=True))
    assert not st.any() and back.tobytes() == keys96.tobytes(), "G1 compress -> decompress round trip"
    row("g1_decompress_validate:keys/s", nkeys, secs, 490 + 1000, {"timing": "wall clock, host pointers (48 B in, 96 B + status out per key)"})
    sig192 = lane.pin["sigs"].numpy()
    nsig = len(sig192) // 192
    c96, st = eng.g2_compress(sig192)
    assert not st.any()
    secs, (back, st) = wall(lambda: eng.g2_decompress(c96))
    assert not st.any() and back.tobytes() == sig192.tobytes(), "G2 compress -> decompress round trip"
    row("g2_decompress:signatures/s", nsig, secs, 1100, {"timing": "wall clock, host pointers (96 B in, 192 B + status out per signature)"})
    off = np.arange(0, nsig + 1, 4, dtype=np.uint32)
    secs, (agg, st) = wall(lambda: eng.g2_aggregate(sig192, off))
    assert not st.any()
    row("g2_aggregate:signatures/s", nsig, secs, 30, {"signatures_per_aggregate": 4, "timing": "wall clock, host pointers"})
    # per-item verification (fast_aggregate_verify of every set of the resident batch: nk keys per item, one final exponentiation each)
    d = lane.d
    acc = torch.zeros(n, dtype=torch.int32, device=dev)
    stt = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        lane.eng.verify_batch_dev(_lib.ITEM_FAST_AGGREGATE, d["sigs"].data_ptr(), d["pks"].data_ptr(), d["pk_off"].data_ptr(),
                                  d["msgs"].data_ptr(), d["msg_off"].data_ptr(), n, acc.data_ptr(), stt.data_ptr())
        ms = lane.eng.last_kernel_ms(0)
        best = ms if best is None else min(best, ms)
    torch.cuda.synchronize()
    assert int(acc.sum()) == n and not bool(stt.any()), "every set of the valid batch verifies on its own"
    row("verify_batch_fast_aggregate:items/s", n, best * 1e-3, 1400 + 6700 + 1170 + 12260 + 8000,
        {"keys_per_item": nk, "timing": "CUDA events of the call, inputs resident in HBM; best of 3"})
    return out


def _bench_lanes(eng, lanes, args, world, rank, local_rank, dev):
    import numpy as np
    import torch
    import torch.distributed as dist
    import milagro_bls_b200 as mb
    from milagro_bls_b200 import sharding
    n, nk = args.sets, args.keys
    S = len(lanes)
    PB = mb._lib.PARTIAL_BYTES
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    base = rank * n
    stage_acc, stage_lock = {}, threading.Lock()

    def add_stages(e):
        st = e.stage_ms()
        with stage_lock:
            for k, v in st.items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(steps, use_lanes, host_inputs, flush_l2):
        """`steps` steps; a step = ONE full verification of a C4-shaped batch ON EVERY LANE (len(use_lanes) concurrent calls, each
        on its own context and host thread), so steps * len(use_lanes) batches in all, batch i on lane i % L.
        The lane thread runs the whole call: its partial Miller product (device-pointer entry, or the host-pointer entry with the
        H2D copies inside), then product + final exponentiation + accept bit on its own context.  N > 1: between the two, this
        thread issues ONE all-gather per step for the L partials of the step (in step order on every rank) and hands the gathered buffer back to the lane, which finishes batch i after issuing its batch i + L."""
        L = len(use_lanes)
        B = steps * L
        partials = torch.zeros(max(B, 1), PB, dtype=torch.uint8, device=dev)
        torch.cuda.current_stream().synchronize()
        ready = [threading.Event() for _ in range(B)]
        gathered_ev = [threading.Event() for _ in range(B)]
        gathered = [None] * B
        results = [None] * B
        errors = []
        full_call = host_inputs and world == 1             # the reference-facing call: b3_verify_multiple on host pointers

        def lane_main(t):
            ln = use_lanes[t]
            pending = None

            def finish(i):
                gathered_ev[i].wait()
                if errors:
                    raise errors[0]
                results[i] = ln.eng.combine_partials_dev(gathered[i].data_ptr(), world)
                add_stages(ln.eng)

            try:
                with torch.cuda.stream(ln.stream):
                    for i in range(t, B, L):
                        if flush_l2:
                            flush.fill_(1)                          # evict L2 between iterations (single-lane mode only)
                            ln.stream.synchronize()
                        p = ln.pin
                        if full_call:
                            ok, fb, _gt = ln.eng.verify_multiple(p["sigs"].numpy(), p["pks"].numpy(), p["pk_off"].numpy(), p["msgs"].numpy(),
                                                                 p["msg_off"].numpy(), p["scal"].numpy().view(np.uint64), want_gt=True)
                            results[i] = (ok, fb)
                            add_stages(ln.eng)
                            continue
                        if host_inputs:
                            ln.eng.verify_multiple_partial(p["sigs"].numpy(), p["pks"].numpy(), p["pk_off"].numpy(), p["msgs"].numpy(),
                                                           p["msg_off"].numpy(), p["scal"].numpy().view(np.uint64), base, partials[i].data_ptr())
                        else:
                            ln.partial_dev(n, base, partials[i].data_ptr())
                        add_stages(ln.eng)
                        if world == 1:
                            results[i] = ln.eng.combine_partials_dev(partials[i].data_ptr(), 1)
                            add_stages(ln.eng)
                            continue
                        # N > 1: hand the partial to the gathering thread and finish the PREVIOUS batch of this lane, whose
                        # all-gather has had a whole batch time to complete (ranks may drift by up to one batch per lane)
                        ready[i].set()
                        if pending is not None:
                            finish(pending)
                        pending = i
                    if pending is not None:
                        finish(pending)
            except BaseException as ex:                              # noqa: BLE001
                errors.append(ex)
                for ev in ready + gathered_ev:
                    ev.set()

        threads = [threading.Thread(target=lane_main, args=(t,), daemon=True) for t in range(L)]
        for th in threads:
            th.start()
        if world > 1:
            for k in range(steps):
                for i in range(k * L, (k + 1) * L):
                    ready[i].wait()
                if errors:
                    break
                # the ONLY collective: one all-gather per step for the L calls in flight (world x L x 592 bytes over NCCL)
                g = sharding.all_gather_partial_batch(partials[k * L:(k + 1) * L], world)
                torch.cuda.current_stream().synchronize()
                for j in range(L):
                    gathered[k * L + j] = g[j]
                    gathered_ev[k * L + j].set()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        return results

    def total_launches():
        return eng.launches + sum(ln.eng.launches for ln in lanes)

    def timed(steps, warmup, use_lanes, host_inputs=False, flush_l2=False):
        for r in run_steps(warmup, use_lanes, host_inputs, flush_l2):
            assert r[0] and r[1] == -1, "verification of the valid synthetic batch must accept"
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = total_launches()
        with stage_lock:
            stage_acc.clear()
        e0.record()                                   # device idle (barrier above): the timestamp is "now"
        res = run_steps(steps, use_lanes, host_inputs, flush_l2)
        torch.cuda.synchronize()                      # every lane's streams
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for r in res:
            assert r[0] and r[1] == -1, "verification of the valid synthetic batch must accept"
        with stage_lock:
            st = {k: v / (steps * len(use_lanes)) for k, v in stage_acc.items()}
        return float(t.item()), total_launches() - launches0, st, res[-1]

    W = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # headline: S batches in flight, device-resident inputs
    ms_res, launches, stages_pipe, last = timed(args.steps, W, lanes)
    clocks = sampler.stop() if rank == 0 else None
    # e2e: the same, from pinned HOST buffers through the host-pointer C ABI (H2D + D2H inside the timed region)
    ms_e2e, _, _, _ = timed(args.steps, W, lanes, host_inputs=True)
    # one batch in flight (call latency), L2 flushed before every step: per-kernel spans for the roofline
    ms_one, _, stages, _ = timed(args.steps, W, lanes[:1], flush_l2=True)
    ms_one_e2e, _, _, _ = timed(args.steps, W, lanes[:1], host_inputs=True, flush_l2=True)
    # untimed diagnostic pass: the same step with the independent stages serialised, for a clean per-stage breakdown
    lanes[0].eng.set_serial(True)
    with stage_lock:
        stage_acc.clear()
    run_steps(2, lanes[:1], False, True)
    with stage_lock:
        stages_serial = {k: v / 2 for k, v in stage_acc.items()}
    lanes[0].eng.set_serial(False)
    # correctness inside the bench: a batch with one flipped message bit must reject
    if rank == 0:
        ln = lanes[0]
        keep = ln.d["msgs"]
        bad = keep.clone()
        bad[5] ^= 1
        torch.cuda.synchronize()
        ln.d["msgs"] = bad
        part = torch.zeros(PB, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ln.partial_dev(n, base, part.data_ptr())
        ln.d["msgs"] = keep
        ok_bad, _ = eng.combine_partials_dev(part.data_ptr(), 1)
        assert not ok_bad, "tampered batch must reject"

    # second headline metric: hash_to_G2/s (b3_hash_to_g2_dev: SHA-256 xmd, SSWU, 3-isogeny, cofactor clearing, affine
    # normalisation and 192-byte wire output), 32-byte messages resident in HBM
    nh = args.h2c_msgs
    hm = torch.from_numpy(np.random.RandomState(5 + rank).randint(0, 256, size=nh * MSG_LEN, dtype=np.uint8)).to(dev)
    ho = torch.arange(0, nh * MSG_LEN + 1, MSG_LEN, dtype=torch.int32, device=dev)
    hout = torch.empty(nh * 192, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for _ in range(2):
        eng.hash_to_g2_dev(hm.data_ptr(), ho.data_ptr(), nh, hout.data_ptr())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.hash_to_g2_dev(hm.data_ptr(), ho.data_ptr(), nh, hout.data_ptr())
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_h2c = float(t.item())
    h2c_rate = nh * world * args.steps / (ms_h2c * 1e-3)

    # rows SURVEY.md 8(f) marks "next" (the callers / data formats either side of the path), measured on rank 0 at N = 1
    next_rows = _bench_next_rows(eng, lanes[0], dev, n, nk) if (world == 1 and not args.no_next_rows) else None

    total_sets = n * world                               # per call across the ranks
    value = total_sets * S * args.steps / (ms_res * 1e-3)
    e2e = total_sets * S * args.steps / (ms_e2e * 1e-3)
    h2d_bytes = lanes[0].h2d_bytes

    if rank == 0:
        peak_mac = eng.imad_peak(wide=True)            # 32x32->64 MACs (IMAD.WIDE pairs) per second, measured live
        peak_imad = eng.imad_peak(wide=False)
        # dominant kernel = the stage with the largest device time when run alone (serialised pass); its duration for the
        # roofline is the CUDA-event span inside the ONE-BATCH-IN-FLIGHT timed region (where it still shares the GPU with
        # the overlapped stages of its own batch, but not with a second batch -- spans of co-running batches are not
        # attributable to one kernel)
        dom = max((k for k in stages_serial if k in FP_MULS and FP_MULS[k] > 0), key=lambda k: stages_serial[k])
        dom_ms = stages[dom]
        units = n + (B3_EXTRA_PAIRS if dom.startswith("miller") else 0)
        macs = FP_MULS[dom] * MACS_PER_FP_MUL * units
        achieved = macs / (dom_ms * 1e-3)
        per_stage = {k: {"ms_timed_region": stages[k], "ms_alone": stages_serial.get(k),
                         "ms_timed_region_pipelined": stages_pipe.get(k),
                         "frac_timed_region": FP_MULS[k] * MACS_PER_FP_MUL * n / (stages[k] * 1e-3) / peak_mac,
                         "frac_alone": FP_MULS[k] * MACS_PER_FP_MUL * n / (stages_serial[k] * 1e-3) / peak_mac if stages_serial.get(k) else None}
                     for k in stages if FP_MULS.get(k, 0) > 0 and stages[k] > 0}
        step_ms = ms_res / (args.steps * S)            # device time per 8192-set call in the headline region
        whole = FP_MULS_PER_SET * MACS_PER_FP_MUL * n / (step_ms * 1e-3)
        whole_one = FP_MULS_PER_SET * MACS_PER_FP_MUL * n / (ms_one / args.steps * 1e-3)
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = peaks["hbm_gbs"], "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        hbm_achieved = BYTES_PER_SET * n / (step_ms * 1e-3) / 1e9
        roofline = {"bound": "imad", "kernel": dom, "achieved": achieved / 1e9, "peak": peak_mac / 1e9, "unit": "GMAC/s (32x32->64 multiply-accumulate)",
                    "frac": achieved / peak_mac,
                    "traffic": NCU_TRAFFIC_BYTES.get(dom) if (n == SETS_PER_GPU and nk == KEYS_PER_SET) else None,
                    "traffic_note": "DRAM bytes read + written per launch, ncu --set full capture (profiles/r1s_h2c_full.txt); the kernel is "
                                    "bound by the integer multiply pipe, not HBM",
                    "kernel_ms": dom_ms, "algorithmic_fp_muls_per_unit": FP_MULS[dom], "macs_per_fp_mul": MACS_PER_FP_MUL, "units_per_launch": units,
                    "peak_source": "live probe b3_imad_peak(wide=1): IMAD.WIDE carry chains, all SMs",
                    "plain_imad_peak_gops": peak_imad / 1e9,
                    "frac_alone": per_stage[dom]["frac_alone"],
                    "whole_step": {"achieved": whole / 1e9, "frac": whole / peak_mac, "fp_muls_per_set": FP_MULS_PER_SET,
                                   "frac_of_executed_work": whole / peak_mac * EXEC_FP_MULS_PER_SET / FP_MULS_PER_SET,
                                   "executed_fp_muls_per_set_estimate": EXEC_FP_MULS_PER_SET,
                                   "frac_one_batch_in_flight": whole_one / peak_mac},
                    "per_stage": per_stage,
                    "hbm": {"achieved_gbs": hbm_achieved, "peak_gbs": hbm_peak, "frac": hbm_achieved / hbm_peak, "peak_source": hbm_src,
                            "algorithmic_bytes_per_set": BYTES_PER_SET},
                    "stage_ms": stages, "stage_ms_serialised": stages_serial, "stage_ms_pipelined": stages_pipe,
                    "note": "kernel spans (stage_ms, frac) come from the one-batch-in-flight timed region (L2 flushed before every step): "
                            "CUDA-event spans on the stream each stage runs on; independent stages of a batch overlap on separate streams, "
                            "so they do not add up to the step.  stage_ms_serialised: same step with the stages run one after another "
                            "(untimed pass).  stage_ms_pipelined: spans inside the headline region, where batches share the GPU.  "
                            "whole_step is the headline region."}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cpu_reference_run(max(os.cpu_count() or 1, args.ref_sets), nk, 0xB200, os.cpu_count() or 1)      # warm-up pass
                passes = [cpu_reference_run(max(os.cpu_count() or 1, args.ref_sets), nk, 0xB200, os.cpu_count() or 1) for _ in range(CPU_PASSES)]
                c = passes[-1]
                c_sets, c_sec = sum(x["sets"] for x in passes), sum(x["seconds"] for x in passes)
                cpu = {"value": c_sets / c_sec, "unit": "sets/s", "cores": c["threads"], "kind": c["kind"],
                       "sample": f"{CPU_PASSES} passes over {c['sets']} sets x {nk} keys, {c['threads']} independent single-threaded instances, "
                                 f"{c_sec:.1f} s wall = {c_sec * c['threads']:.0f} core-seconds"}
            except Exception as ex:                                        # noqa: BLE001
                cpu = {"value": None, "unit": "sets/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
        cache = (f"{S} batches in flight on {S} contexts, each with its own {h2d_bytes / 1e6:.0f} MB of inputs ({S * h2d_bytes / 1e6:.0f} MB > 126 MB L2) "
                 "plus 160 MB of Miller-line scratch written and re-read per step; no explicit flush" if S > 1 else
                 "L2 flushed (256 MiB write) before every step; inputs ~103 MB per GPU")
        out = {"metric": "verified sig-sets/s (verify_multiple_aggregate_signatures, 128 keys/set)", "value": value, "unit": "sets/s",
               "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms_res / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (12x32-bit Montgomery limbs, IMAD.WIDE)",
               "data": "synthetic",
               "config": {"workload": f"verify_multiple_aggregate_signatures: calls of {n} sets x {nk} keys per GPU "
                                      f"({'C4' if world == 1 else 'C5-style'}: {total_sets} sets per call over {world} GPU(s)), 32-byte distinct "
                                      f"messages, 63-bit scalars; a step = {S} such calls in flight per GPU on {S} contexts",
                          "sets_per_call_per_gpu": n, "keys_per_set": nk, "sets_per_call": total_sets, "calls_per_step": S,
                          "sets_per_step": total_sets * S, "parallelism": f"set-sharded x{world}",
                          "batches_in_flight": S, "cache": cache},
               "e2e": {"value": e2e, "unit": "sets/s", "h2d_bytes_per_step": h2d_bytes * world * S, "d2h_bytes_per_step": (576 + 16) * world * S,
                       "ms_per_step": ms_e2e / args.steps},
               "one_batch_in_flight": {"value": total_sets * args.steps / (ms_one * 1e-3), "ms_per_step": ms_one / args.steps,
                                       "e2e_value": total_sets * args.steps / (ms_one_e2e * 1e-3), "e2e_ms_per_step": ms_one_e2e / args.steps,
                                       "cache": "L2 flushed (256 MiB write) before every step"},
               "hash_to_g2": {"value": h2c_rate, "unit": "hash_to_G2/s", "messages_per_gpu": nh, "message_bytes": MSG_LEN,
                              "ms_per_batch": ms_h2c / args.steps,
                              "imad_frac": h2c_rate / world * FP_MULS["hash_to_g2_affine"] * MACS_PER_FP_MUL / peak_mac},
               "next_rows": next_rows,
               "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
               "accept": bool(last[0])}
        if args.breakdown:
            print(json.dumps({"overlapped": stages, "serialised": stages_serial, "pipelined": stages_pipe}, indent=1), file=sys.stderr)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    sys.exit(main())
