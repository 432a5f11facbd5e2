"""CPU study (pure Python, exact integers): can the sequential fp64 sum of a chain of non-negative
terms -- acc = fl(acc + r_n), the M step of the EM -- be evaluated in
parallel WITHOUT changing a bit?

While the accumulator stays in one binade [2^e, 2^(e+1)) with ulp u, fl(acc + r) = acc + u * k where
k = round-to-nearest-integer(r / u), ties (fraction exactly 1/2) to the even accumulator. So a block
of terms is an integer sum plus a 2-state transfer function of the incoming parity of acc / u:
    (increment if acc/u is even, increment if acc/u is odd)
and transfer functions compose associatively -> a parallel scan. A block is valid if the accumulator
does not leave the binade inside it; a block that does is replayed sequentially from where it leaves.

This script checks bit-equality of the blocked evaluation with the plain sequential sum on random
chains shaped like the EM's contributions, and reports how many blocks need the sequential replay."""
import math
import random
import struct
import sys


def ulp_exp(x):
    """e with ulp(x) = 2^e for a normal positive double"""
    m, ex = math.frexp(x)            # x = m * 2^ex, 0.5 <= m < 1
    return ex - 53


def term_transfer(r, e):
    """r / 2^e as (increment if acc is even, increment if odd) for round-to-nearest-even"""
    if r == 0.0:
        return (0, 0)
    m, ex = math.frexp(r)
    mi = int(m * (1 << 53))          # r = mi * 2^(ex - 53), exact
    sh = e - (ex - 53)               # r / 2^e = mi / 2^sh
    if sh <= 0:
        k = mi << (-sh)
        return (k, k)
    if sh > 60:
        return (0, 0)                # far below half an ulp
    q, rem = mi >> sh, mi & ((1 << sh) - 1)
    half = 1 << (sh - 1)
    if rem < half:
        return (q, q)
    if rem > half:
        return (q + 1, q + 1)
    # tie: the result acc + q or acc + q + 1, whichever is even
    return (q + (q & 1), q + 1 - (q & 1))


def compose(f, g):
    """first f, then g (both: parity of the incoming accumulator -> increment)"""
    out = []
    for p in (0, 1):
        a = f[p]
        out.append(a + g[(p + a) & 1])
    return tuple(out)


def blocked_sum(terms, block=32):
    acc = 0.0
    n_replay = n_blocks = 0
    i = 0
    # the first terms (acc == 0 or subnormal/short) are added sequentially
    while i < len(terms) and acc == 0.0:
        acc = acc + terms[i]; i += 1
    while i < len(terms):
        blk = terms[i:i + block]
        n_blocks += 1
        e = ulp_exp(acc)
        A = int(acc / math.ldexp(1.0, e))               # exact: acc is a multiple of its ulp
        tf = (0, 0)
        for r in blk:                                     # (a tree on the device)
            tf = compose(tf, term_transfer(r, e))
        A2 = A + tf[A & 1]
        if A2 < (1 << 53):                                # stayed in the binade: every partial sum did too
            acc = math.ldexp(float(A2), e)
        else:                                             # crosses into the next binade: replay
            n_replay += 1
            for r in blk:
                acc = acc + r
        i += len(blk)
    return acc, n_blocks, n_replay


def main():
    rng = random.Random(1)
    bad = tot_blocks = tot_replay = 0
    n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    for c in range(n_chains):
        n = rng.randint(8, 2000)
        scale = 10.0 ** rng.uniform(-12, 0)
        # contributions r = x * (count / sum): spread over a few decades, some exact repeats and zeros
        terms = []
        for _ in range(n):
            t = rng.random()
            if t < 0.1:
                terms.append(0.0)
            elif t < 0.2 and terms:
                terms.append(terms[rng.randrange(len(terms))])
            else:
                terms.append(scale * 10.0 ** rng.uniform(-6, 0) * rng.random())
        seq = 0.0
        for r in terms:
            seq = seq + r
        blk, nb, nr = blocked_sum(terms)
        tot_blocks += nb; tot_replay += nr
        if struct.pack("<d", seq) != struct.pack("<d", blk):
            bad += 1
    print("chains %d: %d differ; blocks of 32 terms: %d, replayed sequentially (binade crossing): %d (%.2f %%)" % (
        n_chains, bad, tot_blocks, tot_replay, 100.0 * tot_replay / max(tot_blocks, 1)))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
