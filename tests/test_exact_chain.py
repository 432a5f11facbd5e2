"""The claim behind DESIGN.md 9 item 3 (not on the device yet): a sequential fp64 sum of non-negative
terms can be evaluated block-wise with integer increments and parity transfer functions without
changing a bit (tools/exact_chain_study.py)."""
import os
import random
import struct
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import exact_chain_study as ec


def test_blocked_chain_equals_sequential_sum_bit_for_bit():
    rng = random.Random(11)
    for block in (4, 32, 100):
        for _ in range(40):
            n = rng.randint(1, 600)
            scale = 10.0 ** rng.uniform(-15, 3)
            terms = [0.0 if rng.random() < 0.1 else scale * 10.0 ** rng.uniform(-8, 0) * rng.random() for _ in range(n)]
            # exact ties: halves of a power of two added to integers
            if rng.random() < 0.3:
                terms = [float(rng.randint(1, 5))] + [0.5 * 2.0 ** -rng.randint(50, 52) * rng.randint(1, 3) for _ in range(n)]
            seq = 0.0
            for r in terms:
                seq = seq + r
            got, _, _ = ec.blocked_sum(terms, block)
            assert struct.pack("<d", seq) == struct.pack("<d", got), (block, n)


def test_transfer_functions_compose_associatively():
    rng = random.Random(3)
    for _ in range(200):
        f, g, h = [(rng.randint(0, 9), rng.randint(0, 9)) for _ in range(3)]
        assert ec.compose(ec.compose(f, g), h) == ec.compose(f, ec.compose(g, h))
