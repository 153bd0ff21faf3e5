"""Searches a small exact median-of-25 selection network for column-presorted 5x5 windows.

Pipeline (all compare-exchanges, so the 0/1 principle applies and the network can be verified
EXHAUSTIVELY over all 2^25 binary inputs, bit-parallel):
  1. sort each of the 5 columns (shared between the horizontally adjacent outputs of a thread);
  2. sort rank-row r across the 5 columns;
  3. only 13 of the 25 positions can hold the median (SURVEY-style counting argument); a
     sorting network on those 13, read at its 7th output.
Comparators that do not influence the result are then removed greedily, re-verifying all
2^25 inputs after each removal.  Prints the network as C macro lines.
"""
import itertools
import sys

import numpy as np

N = 25
NBITS = 1 << N


def truth_wires():
    # wire i as a bit-vector over all 2^25 inputs: bit k of wire i = (k >> i) & 1, packed in uint64
    idx = np.arange(NBITS // 64, dtype=np.uint64)
    wires = []
    for i in range(N):
        if i < 6:
            pat = 0
            for b in range(64):
                if (b >> i) & 1:
                    pat |= 1 << b
            wires.append(np.full(NBITS // 64, pat, dtype=np.uint64))
        else:
            wires.append(np.where((idx >> np.uint64(i - 6)) & np.uint64(1), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0)))
    return wires


def popcount_ge13(wires):
    # bit-vector: number of ones among the 25 wires >= 13  (reference median of 0/1 inputs)
    # bit-sliced counter (5 bits)
    cnt = [np.zeros_like(wires[0]) for _ in range(5)]
    for w in wires:
        carry = w
        for b in range(5):
            t = cnt[b] & carry
            cnt[b] = cnt[b] ^ carry
            carry = t
    # value >= 13: 13 = 01101b.  ge = c4 | (c3 & ((c2 & (c1 | c0)) | ...)) -> compute via comparison
    c0, c1, c2, c3, c4 = cnt
    ge = c4 | (c3 & c2 & (c1 | c0))      # 8+4+{2|1} = >= 13 when c3,c2 set and (c1 or c0); 
    ge = ge | (c3 & c2 & c1)              # covered above
    # careful: c3&c2 = 12, need +1: (c1|c0). values 13,14,15 ok; >=16 via c4.
    return ge


SORT5 = [(0, 1), (3, 4), (2, 4), (2, 3), (1, 4), (0, 3), (0, 2), (1, 3), (1, 2)]


def batcher(n):
    # Batcher odd-even mergesort comparators for n a power of two
    comps = []
    p = 1
    while p < n:
        k = p
        while k >= 1:
            for j in range(k % p, n - k, 2 * k):
                for i in range(min(k, n - j - k)):
                    if (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        comps.append((i + j, i + j + k))
            k //= 2
        p *= 2
    return comps


def build():
    # element (r, c): index r*5 + c  (r = row in the window, c = column)
    comps = []
    for c in range(5):                       # 1. column sorts
        comps += [(a * 5 + c, b * 5 + c) for a, b in SORT5]
    ncol = len(comps)
    for r in range(5):                       # 2. rank-row sorts
        comps += [(r * 5 + a, r * 5 + b) for a, b in SORT5]
    cand = [(0, 3), (0, 4), (1, 2), (1, 3), (1, 4), (2, 1), (2, 2), (2, 3), (3, 0), (3, 1), (3, 2), (4, 0), (4, 1)]
    cidx = [r * 5 + c for r, c in cand]
    b16 = [(a, b) for a, b in batcher(16) if b < 13]
    comps += [(cidx[a], cidx[b]) for a, b in b16]
    out = cidx[6]
    return comps, ncol, out


def run(comps, wires0, skip=()):
    w = list(wires0)
    for k, (a, b) in enumerate(comps):
        if k in skip:
            continue
        lo = w[a] & w[b]
        hi = w[a] | w[b]
        w[a], w[b] = lo, hi
    return w


def main():
    wires0 = truth_wires()
    want = popcount_ge13(wires0)
    # sanity of the reference on a few random inputs
    rng = np.random.default_rng(0)
    for _ in range(200):
        k = int(rng.integers(0, NBITS))
        ones = bin(k).count("1")
        bit = (int(want[k // 64]) >> (k % 64)) & 1
        assert bit == (1 if ones >= 13 else 0), (k, ones, bit)
    comps, ncol, out = build()
    w = run(comps, wires0)
    assert np.array_equal(w[out], want), "pipeline is not a median network"
    print("full pipeline: %d comparators (%d column + %d rest)" % (len(comps), ncol, len(comps) - ncol), file=sys.stderr)
    # greedy removal of the non-column comparators (the column sorts are shared, keep them whole)
    skip = set()
    for k in range(len(comps) - 1, ncol - 1, -1):
        trial = skip | {k}
        w = run(comps, wires0, trial)
        if np.array_equal(w[out], want):
            skip = trial
    rest = [c for k, c in enumerate(comps) if k >= ncol and k not in skip]
    print("after pruning: %d comparators after the column sorts" % len(rest), file=sys.stderr)
    # min/max instruction count after dead-output elimination
    live = {out}
    ops = 0
    for a, b in reversed(rest):
        na = a in live
        nb = b in live
        if na or nb:
            ops += int(na) + int(nb)
            live |= {a, b}
    print("min/max instructions after dead-output elimination: %d (+ %d per shared column sort)" % (ops, 18), file=sys.stderr)
    print("OUT %d" % out)
    print("REST " + " ".join("%d,%d" % c for c in rest))


if __name__ == "__main__":
    main()
