"""Multi-GPU sharding of verify_multiple_aggregate_signatures.

One process per GPU.  Signature sets are independent up to one Fp12 product, so they are split contiguously by rank;
rank r verifies its shard with global set indices (so `first_bad` means the same as in the one-GPU run), producing a
592-byte partial (Miller-loop product + first failing index).  ONE all-gather of the partials is the only collective
on the path; every rank then multiplies the partials and runs the single final exponentiation.

The product's collective lives INSIDE the C library (b3_comm_create / b3_verify_multiple_sharded: ncclAllGather issued by the
library, include/milagro_bls_b200.h; bench.py and tests/test_gpu_multirank.py use that).  This module holds the host-side
shard arithmetic and the same exchange written over torch.distributed for hosts that run their own collective
(b3_verify_multiple_partial* + b3_combine_partials_dev): it only moves bytes, so it runs over NCCL (GPU tensors) and over
gloo (CPU tensors, tests/test_sharding.py).
"""
import torch
import torch.distributed as dist

PARTIAL_BYTES = 592
NO_BAD = 0x7FFFFFFFFFFFFFFF


def shard_range(n, rank, world):
    """Contiguous shard [start, stop) of n sets for `rank`; the first n % world ranks get one extra set."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_partials(partial, world, group=None):
    """partial: uint8 tensor of PARTIAL_BYTES on this rank -> uint8 tensor of world * PARTIAL_BYTES, rank-major."""
    if partial.dtype != torch.uint8 or partial.numel() != PARTIAL_BYTES:
        raise ValueError("partial must be a uint8 tensor of %d bytes" % PARTIAL_BYTES)
    if world == 1:
        return partial.clone()
    out = torch.empty(world * PARTIAL_BYTES, dtype=torch.uint8, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous(), group=group)
    return out


def all_gather_partial_batch(partials, world, group=None):
    """Several calls in flight per rank: partials = uint8 tensor (L, PARTIAL_BYTES), row j = this rank's partial of call j.
    ONE all-gather for all L calls -> uint8 tensor (L, world, PARTIAL_BYTES): row j holds the world partials of call j,
    rank-major, ready for b3_combine_partials_dev."""
    if partials.dtype != torch.uint8 or partials.dim() != 2 or partials.shape[1] != PARTIAL_BYTES:
        raise ValueError("partials must be a (L, %d) uint8 tensor" % PARTIAL_BYTES)
    L = partials.shape[0]
    if world == 1:
        return partials.reshape(L, 1, PARTIAL_BYTES).clone()
    out = torch.empty(world * L * PARTIAL_BYTES, dtype=torch.uint8, device=partials.device)
    dist.all_gather_into_tensor(out, partials.contiguous().view(-1), group=group)
    return out.view(world, L, PARTIAL_BYTES).transpose(0, 1).contiguous()


def first_bad_of(partials):
    """Minimum of the int64 first-bad words of a gathered partial buffer; -1 if no shard saw a bad signature."""
    t = partials.view(-1, PARTIAL_BYTES)[:, 576:584].contiguous().view(torch.int64).reshape(-1)
    m = int(t.min().item())
    return -1 if m == NO_BAD else m


def verify_multiple_sharded(eng, d, n_local, index_base, world, group=None, want_gt=False):
    """d: dict of device tensors of THIS rank's shard (sigs, pks, pk_off or None, msgs, msg_off, scal).
    Returns what Engine.combine_partials_dev returns: (accept, first_bad[, gt]) -- identical on every rank."""
    dev = d["sigs"].device
    partial = torch.empty(PARTIAL_BYTES, dtype=torch.uint8, device=dev)
    pk_off = d.get("pk_off")
    eng.verify_multiple_partial_dev(d["sigs"].data_ptr(), d["pks"].data_ptr(), pk_off.data_ptr() if pk_off is not None else 0,
                                    d["msgs"].data_ptr(), d["msg_off"].data_ptr(), d["scal"].data_ptr(), n_local, index_base,
                                    partial.data_ptr())
    gathered = all_gather_partials(partial, world, group)
    torch.cuda.current_stream(dev).synchronize()
    return eng.combine_partials_dev(gathered.data_ptr(), world, want_gt=want_gt)
