"""Host-side logic of the multi-GPU path (milagro_bls_b200/sharding.py) on CPU: world_size 2 over gloo.

The GPU kernels cannot run here, so each rank's partial is produced by the ORACLE (Miller product of its shard in the
reference's 576-byte GT layout + first-bad word); what is under test is the sharding itself: contiguous ranges with
global indices, the single all-gather, the minimum over first-bad words, and that the product of the gathered
partials followed by ONE final exponentiation equals the unsharded result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from milagro_bls_b200 import sharding
from oracle import bls_oracle as O


def test_shard_range_covers_everything():
    for n in [0, 1, 7, 8, 8192, 65536, 65537]:
        for world in [1, 2, 3, 4, 8]:
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _f12_from_bytes(b):                 # inverse of O.f12_to_bytes (A/fp12.rs:784-856)
    v = [int.from_bytes(b[48 * i:48 * i + 48], "big") for i in range(12)]
    return tuple(tuple((v[4 * a + 2 * c], v[4 * a + 2 * c + 1]) for c in range(2)) for a in range(3))


def _sets(n):
    out = []
    for j in range(n):
        sk = 1000 + 7 * j
        msg = bytes([j]) * 32
        out.append((O.sign(sk, msg), O.sk_to_pk(sk), msg))
    return out


def _partial_bytes(sets, scalars, index_base, bad_local=None):
    """Oracle stand-in for b3_verify_multiple_partial_dev: prod_j miller(H_j, [c_j] pk_j) * miller(sum [c_j] sig_j, -G1)."""
    f = O.F12_ONE
    S = None
    for (sig, pk, msg), c in zip(sets, scalars):
        f = O.f12_mul(f, O.ate2(O.hash_to_curve_g2(msg), O.g1_mul(pk, c), None, None))
        S = O.g2_add(S, O.g2_mul(sig, c))
    f = O.f12_mul(f, O.ate2(S, O.NEG_G1, None, None))
    fb = sharding.NO_BAD if bad_local is None else index_base + bad_local
    return O.f12_to_bytes(f) + int(fb).to_bytes(8, "little") + bytes(8)


def _worker(rank, world, port, n, bad_at, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sets = _sets(n)
        scalars = [3 + 5 * j for j in range(n)]
        a, b = sharding.shard_range(n, rank, world)
        bad_local = bad_at - a if (bad_at is not None and a <= bad_at < b) else None
        part = torch.frombuffer(bytearray(_partial_bytes(sets[a:b], scalars[a:b], a, bad_local)), dtype=torch.uint8)
        gathered = sharding.all_gather_partials(part, world)
        assert gathered.numel() == world * sharding.PARTIAL_BYTES
        # rank-major order: slot r holds rank r's first-bad word
        raw = bytes(gathered.numpy().tobytes())
        f = O.F12_ONE
        for r in range(world):
            rec = raw[r * sharding.PARTIAL_BYTES:(r + 1) * sharding.PARTIAL_BYTES]
            f = O.f12_mul(f, _f12_from_bytes(rec[:576]))
        gt = O.f12_to_bytes(O.fexp(f))
        # several calls in flight: one all-gather for L = 3 calls; row j = the world partials of call j, rank-major
        tag = lambda r, j: bytes([16 * r + j]) * sharding.PARTIAL_BYTES
        mine = torch.frombuffer(bytearray(b"".join(tag(rank, j) for j in range(3))), dtype=torch.uint8).view(3, -1)
        both = sharding.all_gather_partial_batch(mine, world)
        assert tuple(both.shape) == (3, world, sharding.PARTIAL_BYTES) and both.is_contiguous()
        for j in range(3):
            for r in range(world):
                assert bytes(both[j, r].numpy().tobytes()) == tag(r, j)
        out_q.put((rank, sharding.first_bad_of(gathered), gt))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("bad_at", [None, 2])
def test_two_rank_gloo_matches_unsharded(bad_at):
    n, world = 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bad_at, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # unsharded oracle run with the same scalars
    sets = _sets(n)
    it = iter(3 + 5 * j for j in range(n))
    ok, gt = O.verify_multiple_aggregate_signatures(lambda k: int(next(it)).to_bytes(8, "big"), sets, want_gt=True)
    assert ok
    for rank, fb, gt_r in res:
        assert gt_r == O.f12_to_bytes(gt)                 # identical GT on every rank, equal to the one-process value
        assert fb == (-1 if bad_at is None else bad_at)   # global index of the first bad set
