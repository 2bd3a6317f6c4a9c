"""Enumerates the thread -> shared-memory mappings of the three-stage block-Jacobi kernels (csrc/jacobi_uniform_q3p.cuh,
csrc/jacobi_uniform_q4p.cuh) and checks (a) every slot of the tile is touched exactly once per stage and (b) no access
instruction has a bank conflict (64-bit: the 16 lanes of a half warp on 16 distinct 8-byte banks; 128-bit: the 8 lanes of a
quarter warp on 8 distinct 16-byte units).  The formulas below restate the index arithmetic of the kernels; run this after
changing a layout constant or a lane assignment, before spending GPU time."""
import numpy as np


def distinct(addr, group, mod):
    """worst number of lanes sharing a bank within any group of `group` consecutive lanes"""
    worst = 0
    for s in range(0, len(addr), group):
        lanes = addr[s:s + group]
        worst = max(worst, len(lanes) - len(set(int(a) % mod for a in lanes)))
    return worst


def q3():
    RS = 258                                   # kQ3jRStride
    tid = np.arange(256)
    # stages A / C: plane (fixed x-node i) of element (ex, row)
    row = 2 * ((tid >> 2) & 3) + ((tid >> 4) & 1) + 8 * ((tid >> 5) & 1)
    ex, i = tid >> 6, tid & 3
    seen, worst = set(), 0
    for k in range(4):
        for j in range(4):
            addr = RS * row + 64 * ex + i + 4 * j + 16 * k
            seen.update(int(a) for a in addr)
            worst = max(worst, distinct(addr, 16, 16))
    assert len(seen) == 4096 and worst == 0, ("q3 A/C", len(seen), worst)
    # the bulk stores of stage C: lanes 0..7 of a warp send exactly the elements the warp owns
    for w in range(8):
        own = {(int(r), int(x)) for r, x in zip(row[w * 32:w * 32 + 32], ex[w * 32:w * 32 + 32])}
        sent = {((l & 7) + 8 * ((l >> 5) & 1), l >> 6) for l in range(w * 32, w * 32 + 8)}
        assert own == sent, ("q3 store ownership", w)
    # stage B: x-line (j, k) of the 4 elements of a row, 128-bit accesses
    j, k, rowb = tid & 3, (tid >> 3) & 3, ((tid >> 2) & 1) + 2 * (tid >> 5)
    seen, worst = set(), 0
    for e in range(4):
        word = RS * rowb + 64 * e + 4 * j + 16 * k
        assert np.all(word % 2 == 0)
        for half in range(2):
            unit = word // 2 + half
            seen.update(int(u) for u in unit)
            worst = max(worst, distinct(unit, 8, 8))
    assert len(seen) == 2048 and worst == 0, ("q3 B", len(seen), worst)
    # prefetch: lanes 0, 1 of every warp fetch all 16 rows
    rows = sorted((t & 1) + 2 * (t >> 5) for t in tid if (t & 31) < 2)
    assert rows == list(range(16))


def q4():
    RS, N3 = 500, 125
    tid = np.arange(160)
    w, lane = tid >> 5, tid & 31
    ebase = RS * (lane >> 2) + N3 * (lane & 3)
    for name, addr_of in (("A/C", lambda a, b: ebase + w + 5 * a + 25 * b), ("B", lambda a, b: ebase + 25 * w + 5 * a + b)):
        seen, worst = set(), 0
        for a in range(5):
            for b in range(5):
                addr = addr_of(a, b)
                seen.update(int(x) for x in addr)
                worst = max(worst, distinct(addr, 16, 16))
        assert len(seen) == 4000 and worst == 0, ("q4 " + name, len(seen), worst)
    rows = sorted(r for t in tid if (t & 31) == 0 for r in range(t >> 5, 8, 5))
    assert rows == list(range(8))


if __name__ == "__main__":
    q3()
    q4()
    print("jacobi kernel mappings: complete and bank-conflict free")
