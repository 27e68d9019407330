"""The sharded form of verify_multiple_aggregate_signatures behind the C ABI (b3_comm_*, b3_sharded_begin / _finish,
b3_verify_multiple_sharded): SURVEY.md 8e, BASELINE config C5.

  * one GPU (always runs): a 1-rank communicator with two lanes (two contexts, two host threads) pipelining begin(k + 1) before
    finish(k); every call must reproduce the unsharded call bit for bit (accept, first_bad, GT bytes), which the oracle pins.
  * >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py`, skipped on a 1-GPU box): one process per GPU,
    NCCL all-gather of the 592-byte partials issued from inside the library; every rank's GT equals the 1-rank GT of the
    whole batch and the oracle's, a tampered set on the last rank rejects everywhere with the oracle's GT, and a non-subgroup
    signature on the last rank comes back as the GLOBAL first_bad index on every rank.
"""
import os
import sys
import tempfile
import threading
import time

import numpy as np
import pytest

from oracle import bls_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def g1w(P):
    return O.serialize_uncompressed_g1(P)


def g2w(P):
    return O.serialize_uncompressed_g2(P)


def _global_batch(n_sets=10, n_keys=3):
    """Deterministic global batch (every rank rebuilds the same one): sets, wire arrays, scalars in GLOBAL set order."""
    from milagro_bls_b200 import SeededRng, draw_scalar
    sets = []
    for j in range(n_sets):
        sks = [7000 + 131 * j + 17 * i for i in range(n_keys)]
        msg = bytes([j, 0xC5]) * 16
        sig = O.g2_mul(O.hash_to_curve_g2(msg), sum(sks) % O.r)
        sets.append((sig, [O.sk_to_pk(s) for s in sks], msg))
    rng = SeededRng(b"c5")
    scalars = np.array([draw_scalar(rng) for _ in sets], dtype=np.uint64)
    return sets, scalars


def _wire(sets, lo, hi, msgs=None, sig_override=None):
    sub = sets[lo:hi]
    sigs = b"".join(g2w(sig_override[lo + j]) if sig_override and (lo + j) in sig_override else g2w(s) for j, (s, _, _) in enumerate(sub))
    pks = b"".join(g1w(P) for _, p, _ in sub for P in p)
    nk = len(sets[0][1])
    offs = np.arange(0, nk * len(sub) + 1, nk, dtype=np.uint32)
    ms = [m for _, _, m in sub] if msgs is None else msgs[lo:hi]
    moff = np.cumsum([0] + [len(m) for m in ms]).astype(np.uint32)
    return sigs, pks, offs, b"".join(ms), moff


def _cases(sets):
    n = len(sets)
    msgs = [m for _, _, m in sets]
    tam = list(msgs)
    tam[n - 2] = b"tampered" * 4
    return [("valid", None, None), ("tampered", tam, None), ("non_subgroup", None, {n - 1: O.map_to_curve_g2((5, 7))})]


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as g
    g.build_cuda()
    import milagro_bls_b200 as mb
    return mb.Engine(0)


def test_one_rank_two_lanes_pipelined(eng):
    import milagro_bls_b200 as mb
    sets, scalars = _global_batch()
    n = len(sets)
    comm = mb.Comm(0, 1, 0, None, lanes=2)
    engines = [eng, mb.Engine(0)]
    want = {}
    for name, msgs, sigo in _cases(sets):
        w = _wire(sets, 0, n, msgs, sigo)
        want[name] = eng.verify_multiple(*w, scalars, want_gt=True)
    ok_o, gt_o = O.verify_multiple_aggregate_signatures(O.SeededRng(b"c5").fill, [(s, O.aggregate_public_keys(p), m) for s, p, m in sets], want_gt=True)
    assert want["valid"] == (True, -1, O.f12_to_bytes(gt_o)) and ok_o
    assert want["tampered"][0] is False and want["non_subgroup"][:2] == (False, n - 1)
    results, errors = {}, []

    def lane_main(lane):
        try:
            e = engines[lane]
            order = _cases(sets) if lane == 0 else _cases(sets)[::-1]         # the lanes run DIFFERENT calls in the same step
            pending = None
            for name, msgs, sigo in order:
                w = _wire(sets, 0, n, msgs, sigo)
                t = e.sharded_begin(comm, lane, None, *w, scalars, 0)
                if pending is not None:                                    # finish step k after beginning step k + 1
                    results[(lane, pending[0])] = e.sharded_finish(comm, lane, pending[1], want_gt=True)
                pending = (name, t)
            results[(lane, pending[0])] = e.sharded_finish(comm, lane, pending[1], want_gt=True)
        except BaseException as ex:                                           # noqa: BLE001
            errors.append(ex)

    ths = [threading.Thread(target=lane_main, args=(t,)) for t in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(120)
    assert not errors, errors
    assert comm.collectives == 3                                            # one per step, not one per call
    for lane in range(2):
        for name in want:
            k = 2 if name == "non_subgroup" else 3            # (GT of a batch holding a point outside G2 is not pinned: no bilinearity)
            assert results[(lane, name)][:k] == want[name][:k], (lane, name)
    # the key-table form of the sharded call
    tbl = mb.KeyTable(eng)
    keys = b"".join(g1w(P) for _, p, _ in sets for P in p)
    first, st = tbl.append(keys, compressed=False, validate=True)
    assert not st.any()
    w = _wire(sets, 0, n)
    idx = np.arange(first, first + len(keys) // 96, dtype=np.uint32)
    comm1 = mb.Comm(0, 1, 0, None, lanes=1)
    t = eng.sharded_begin(comm1, 0, tbl, w[0], idx, w[2], w[3], w[4], scalars, 0)
    assert eng.sharded_finish(comm1, 0, t, want_gt=True) == want["valid"]
    comm1.close()
    comm.close()
    engines[1].close()
    tbl.close()


# ----------------------------------------------------------------------------------------------- >= 2 GPUs
def _rank_main(rank, world, id_path, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import milagro_bls_b200 as mb
    torch.cuda.set_device(rank)
    eng = mb.Engine(rank)
    if rank == 0:
        uid = mb.nccl_unique_id()
        with open(id_path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(id_path + ".tmp", id_path)
    else:
        t0 = time.time()
        while not os.path.exists(id_path):
            if time.time() - t0 > 120:
                raise RuntimeError("no NCCL id from rank 0")
            time.sleep(0.05)
        uid = open(id_path, "rb").read()
    comm = mb.Comm(rank, world, rank, uid, lanes=1)
    sets, scalars = _global_batch(n_sets=4 * world + 1)
    n = len(sets)
    per = (n + world - 1) // world
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    lines = []
    for name, msgs, sigo in _cases(sets):
        whole = eng.verify_multiple(*_wire(sets, 0, n, msgs, sigo), scalars, want_gt=True)           # the 1-rank result, on this rank
        w = _wire(sets, lo, hi, msgs, sigo)
        t = eng.sharded_begin(comm, 0, None, *w, scalars[lo:hi], lo)
        got = eng.sharded_finish(comm, 0, t, want_gt=True)
        k = 2 if name == "non_subgroup" else 3                # (GT of a batch holding a point outside G2 is not pinned: no bilinearity)
        lines.append((name, got[:k] == whole[:k], got[0], got[1], got[2].hex()))
    with open(out_path % rank, "w") as f:
        for ln in lines:
            f.write("\t".join(str(x) for x in ln) + "\n")
    comm.close()
    eng.close()


def test_two_ranks_nccl_matches_one_rank_and_oracle():
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    import torch.multiprocessing as mp
    import __graft_entry__ as g
    g.build_cuda()
    d = tempfile.mkdtemp()
    id_path, out_path = os.path.join(d, "nccl_id"), os.path.join(d, "rank%d.tsv")
    mp.spawn(_rank_main, args=(world, id_path, out_path), nprocs=world, join=True)
    sets, scalars = _global_batch(n_sets=4 * world + 1)
    n = len(sets)
    oracle = {}
    for name, msgs, sigo in _cases(sets):
        if name == "non_subgroup":
            continue
        ms = [m for _, _, m in sets] if msgs is None else msgs
        ok_o, gt_o = O.verify_multiple_aggregate_signatures(O.SeededRng(b"c5").fill,
                                                            [(s, O.aggregate_public_keys(p), m) for (s, p, _), m in zip(sets, ms)], want_gt=True)
        oracle[name] = (ok_o, O.f12_to_bytes(gt_o).hex())
    for rank in range(world):
        rows = [ln.rstrip("\n").split("\t") for ln in open(out_path % rank)]
        assert [r[0] for r in rows] == ["valid", "tampered", "non_subgroup"]
        for name, same, ok, fb, gt in rows:
            assert same == "True", (rank, name)                              # every rank: sharded == 1-rank, bit for bit
            if name in oracle:
                assert (ok == "True") == oracle[name][0] and gt == oracle[name][1] and fb == "-1", (rank, name)
            else:
                assert ok == "False" and int(fb) == n - 1, (rank, name)      # GLOBAL index of the bad signature, on every rank
