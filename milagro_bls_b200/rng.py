"""Caller-injected randomness for verify_multiple_aggregate_signatures.

The reference takes any `rand::Rng` (M/src/aggregates.rs:261) and draws one scalar per set with the rule of
M/src/aggregates.rs:278-287: 8 bytes -> i64::from_be_bytes -> abs(); zero is redrawn.  Any object with a
`fill(n) -> bytes` method can be injected here; SeededRng is a deterministic stream for tests and benchmarks."""
import hashlib


class SeededRng:
    """SHA-256 in counter mode over a seed: fill(n) returns the next n bytes of the stream."""

    def __init__(self, seed: bytes):
        self.seed, self.ctr, self.buf = bytes(seed), 0, b""

    def fill(self, n: int) -> bytes:
        while len(self.buf) < n:
            self.buf += hashlib.sha256(self.seed + self.ctr.to_bytes(8, "big")).digest()
            self.ctr += 1
        out, self.buf = self.buf[:n], self.buf[n:]
        return out


def draw_scalar(rng) -> int:
    """One batch scalar in [1, 2^63 - 1] (M/src/aggregates.rs:278-287).  i64::MIN, where the reference's abs()
    overflows (probability 2^-64), is treated as out of contract and redrawn."""
    while True:
        v = int.from_bytes(rng.fill(8), "big", signed=True)
        if v == -(1 << 63):
            continue
        v = abs(v)
        if v:
            return v
