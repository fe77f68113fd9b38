"""Model check of the peer-exchange buffer protocol (csrc/topk_exchange.cu) under arbitrary rank
skew — the part the single-device staged tests cannot reach.

Every rank owns two parity halves with one slot per peer.  P(n) stores epoch n's list into slot
`rank` of half n & 1 of EVERY rank's buffer and then raises its flag to n everywhere; M(n) may run
only once all G flags of the own buffer are >= n and then reads the G slots of half n & 1, which
must still hold epoch n.  A random scheduler interleaves the ranks' programs; the claim under test
is the header comment's: two halves are enough for the fused order (P n, M n) and for the lagged
order (M n-1, then P n), for any skew the flags allow.  The last test shows the model has teeth:
the order (P n, then M n-1) does clobber a half that is still to be read.
"""
import random

import pytest


def run(programs, G, seed):
    """programs[r] = list of ('P', n) / ('M', n).  Returns None or a description of the violation."""
    rng = random.Random(seed)
    slot = [[[0] * G for _ in range(2)] for _ in range(G)]     # slot[owner][half][writer] = epoch
    flag = [[0] * G for _ in range(G)]                         # flag[owner][writer]
    pc = [0] * G
    while True:
        ready = []
        for r in range(G):
            if pc[r] == len(programs[r]):
                continue
            op, n = programs[r][pc[r]]
            if op == "P" or all(flag[r][g] >= n for g in range(G)):
                ready.append(r)
        if not ready:
            if all(pc[r] == len(programs[r]) for r in range(G)):
                return None
            return f"deadlock at {pc}"
        r = rng.choice(ready)
        op, n = programs[r][pc[r]]
        if op == "P":
            for owner in range(G):
                slot[owner][n & 1][r] = n
            for owner in range(G):                              # flags after the data, like the kernel
                flag[owner][r] = n
        else:
            for g in range(G):
                if slot[r][n & 1][g] != n:
                    return f"rank {r} merging epoch {n} found epoch {slot[r][n & 1][g]} from rank {g}"
        pc[r] += 1


def fused(steps):
    return [op for n in range(1, steps + 1) for op in (("P", n), ("M", n))]


def lagged(steps, first=1):
    prog = [("P", first)]
    for n in range(first + 1, first + steps):
        prog += [("M", n - 1), ("P", n)]
    return prog + [("M", first + steps - 1)]


@pytest.mark.parametrize("G", [2, 3, 8])
def test_two_halves_suffice_for_fused_and_lagged_streams(G):
    for seed in range(300):
        assert run([fused(9)] * G, G, seed) is None
        assert run([lagged(9)] * G, G, seed) is None
        # warm-up with fused searches, then a lagged stream, then fused again (bench.py's sequence)
        mixed = fused(3) + lagged(6, first=4) + [op for n in (10, 11) for op in (("P", n), ("M", n))]
        assert run([mixed] * G, G, seed) is None


def test_model_detects_the_unsafe_order():
    """Push n BEFORE merging n-1 needs more than two halves: some schedule clobbers a half."""
    def unsafe(steps):
        prog = [("P", 1)]
        for n in range(2, steps + 1):
            prog += [("P", n), ("M", n - 1)]
        return prog + [("M", steps)]
    assert any(run([unsafe(9)] * 3, 3, seed) is not None for seed in range(300))
