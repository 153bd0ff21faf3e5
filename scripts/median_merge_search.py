"""Searches the comparator networks of the merge-based 5x5 median (k_median5, round 2).

Scheme (every step is min/max only, so the 0/1 principle applies, restricted to the inputs that
satisfy each step's precondition -- thresholding a real input keeps sortedness):
  1. each image column of a window is sorted once (SORT5, 9 exchanges), shared by the 5 windows
     that contain it;
  2. MERGE55: two adjacent sorted columns -> one sorted list of 10 (shared by the windows that
     contain both columns);
  3. CORE: two sorted 10-lists (4 adjacent columns = the core two horizontally adjacent windows
     share) -> the core's order statistics 7..12 (0-based) in order.  A core element with k core
     elements below it has rank k..k+5 in a window = core + one more column, so only those six can
     be the window's median (rank 12);
  4. FINAL: median = min_j max(core[12-j], c[j-1]) over j = 0..5 (c = the window's fifth column,
     sorted; the j = 0 term is core[12] alone): the rank-12 element of the union of two sorted lists.
Each network starts from a Batcher sorter and is pruned greedily (random orders, best kept),
re-verifying ALL admissible 0/1 inputs after every removal.  Prints C macro bodies and op counts
(min/max instructions after dropping the halves nobody reads).

usage: python scripts/median_merge_search.py [tries]
"""
import itertools
import random
import sys


def batcher(n):
    """Batcher odd-even mergesort on n wires (n any size: built for the next power of two, comparators
    touching missing wires dropped -- they would compare with +inf)."""
    m = 1
    while m < n:
        m *= 2
    comps = []
    p = 1
    while p < m:
        k = p
        while k >= 1:
            for j in range(k % p, m - k, 2 * k):
                for i in range(min(k, m - j - k)):
                    if (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        comps.append((i + j, i + j + k))
            k //= 2
        p *= 2
    return [(a, b) for a, b in comps if a < n and b < n]


def oe_merge(A, B):
    """Batcher's odd-even merge of two sorted wire lists of any lengths: (comparators, wires of the
    result in ascending order).  A comparator (x, y) leaves the minimum on x, the maximum on y."""
    if not A:
        return [], list(B)
    if not B:
        return [], list(A)
    if len(A) == 1 and len(B) == 1:
        return [(A[0], B[0])], [A[0], B[0]]
    c1, E = oe_merge(A[0::2], B[0::2])
    c2, O = oe_merge(A[1::2], B[1::2])
    comps, out = c1 + c2, [E[0]]
    k = min(len(O), len(E) - 1)
    for i in range(k):
        comps.append((O[i], E[i + 1]))
        out += [O[i], E[i + 1]]
    return comps, out + E[k + 1:] + O[k:]


def run(comps, v):
    v = list(v)
    for a, b in comps:
        if v[a] > v[b]:
            v[a], v[b] = v[b], v[a]
    return v


def sorted01(n):
    return [[0] * (n - k) + [1] * k for k in range(n + 1)]


def admissible(groups):
    """all 0/1 inputs whose wire groups (lists of wire indices) are each sorted ascending"""
    n = sum(len(g) for g in groups)
    out = []
    for combo in itertools.product(*[sorted01(len(g)) for g in groups]):
        v = [0] * n
        for g, vals in zip(groups, combo):
            for w, x in zip(g, vals):
                v[w] = x
        out.append(v)
    return out


def ok(comps, inputs, want):
    """want: list of (wire, rank): after the network, `wire` must hold the rank-th smallest input"""
    for v in inputs:
        r = run(comps, v)
        s = sorted(v)
        for w, k in want:
            if r[w] != s[k]:
                return False
    return True


def prune(comps, inputs, want, tries, seed=0):
    rng = random.Random(seed)
    best = None
    for t in range(tries):
        cur = list(comps)
        order = list(range(len(cur)))
        if t > 0:
            rng.shuffle(order)
        else:
            order.reverse()
        removed = set()
        changed = True
        while changed:
            changed = False
            for idx in order:
                if idx in removed:
                    continue
                trial = [c for k, c in enumerate(cur) if k not in removed and k != idx]
                if ok(trial, inputs, want):
                    removed.add(idx)
                    changed = True
        res = [c for k, c in enumerate(cur) if k not in removed]
        cost = ops(res, [w for w, _ in want])
        if best is None or cost < best[0]:
            best = (cost, res)
    return best[1]


def ops(comps, outs):
    """min/max instructions after dead-half elimination: a comparator (a, b) writes min -> a, max -> b"""
    live = set(outs)
    n = 0
    for a, b in reversed(comps):
        la, lb = a in live, b in live
        n += la + lb
        if la or lb:
            live.add(a)
            live.add(b)
        # a dead half is simply not computed; the wire keeps being "not live" above this point
        if not la:
            pass
        if not lb:
            pass
    return n


def emit(name, comps):
    body = " ".join("X(%d, %d)" % c for c in comps)
    print("#define %s(X) \\\n    %s" % (name, body))


def main():
    tries = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    # 2. MERGE55: wires 0..4 = column A (sorted), 5..9 = column B (sorted)
    A, B = list(range(5)), list(range(5, 10))
    m55, out = oe_merge(A, B)
    inp = admissible([A, B])
    assert ok(m55, inp, [(w, k) for k, w in enumerate(out)])
    print("// MERGE55: %d exchanges, %d min/max; result ascending on wires %s" % (len(m55), ops(m55, out), out))
    emit("TVL1_MERGE55", m55)
    # 3. CORE: wires 0..9 and 10..19 sorted -> ranks 7..12
    A, B = list(range(10)), list(range(10, 20))
    full, out = oe_merge(A, B)
    inp = admissible([A, B])
    want = [(out[k], k) for k in range(7, 13)]
    assert ok(full, inp, want)
    core = prune(full, inp, want, tries)
    print("// CORE (ranks 7..12 of two sorted 10-lists, on wires %s): %d exchanges, %d min/max" % (
        [w for w, _ in want], len(core), ops(core, [w for w, _ in want])))
    emit("TVL1_CORE20", core)
    # 4. FINAL check of the closed form on all admissible 0/1 inputs (6 core values sorted + 5 column values sorted;
    #    the 14 other core values are 7 below and 7 above by construction)
    bad = 0
    for c6 in sorted01(6):
        for col in sorted01(5):
            allv = [0] * 7 + c6 + [1] * 7 + col          # core ranks 0..6 <= c6 <= ranks 13..19 -- as 0/1: consistent cases only
            if c6[0] == 0 and False:
                pass
            med = sorted(allv)[12]
            terms = [c6[5]] + [max(c6[5 - j], col[j - 1]) for j in range(1, 6)]
            if min(terms) != med:
                bad += 1
    print("// FINAL closed form: %d mismatches on the admissible 0/1 inputs" % bad)


if __name__ == "__main__":
    main()
