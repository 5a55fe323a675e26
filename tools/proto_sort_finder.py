"""CPU prototype for the next round's match finder (notes): a per-frame stable sort of
(32-bit hash, position) keys yields, for every position, the nearest earlier position with the same hash --
the candidate a collision-free "latest occurrence" table would return -- as a streaming, data-parallel computation.
Checks the sort formulation against the sequential dictionary formulation and reports how often the candidate is a real match."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests")]
import numpy as np
import __graft_entry__ as ge

PRIME8 = np.uint64(0xCF1BBCDCB7A56463); PRIME5 = np.uint64(889523592379)
pkg = ge.load_package()
frame = pkg.corpus.g2(1 << 20)
n = frame.size - 8
v = np.zeros(n, dtype=np.uint64)
for k in range(8):                                           # little-endian 8-byte words at every position
    v |= frame[k:k + n].astype(np.uint64) << np.uint64(8 * k)
with np.errstate(over="ignore"):
    h8 = ((v * PRIME8) >> np.uint64(32)).astype(np.uint32)
    h5 = (((v << np.uint64(24)) * PRIME5) >> np.uint64(32)).astype(np.uint32)

def by_sort(h):
    order = np.argsort(h, kind="stable")                     # positions grouped by hash, ascending position inside a group
    hs = h[order]
    prev = np.full(n, -1, dtype=np.int64)
    same = hs[1:] == hs[:-1]
    prev[order[1:][same]] = order[:-1][same]                 # predecessor in the group = nearest earlier occurrence
    return prev

def by_dict(h):
    last = {}; prev = np.full(n, -1, dtype=np.int64)
    for p, x in enumerate(h.tolist()):
        prev[p] = last.get(x, -1); last[x] = p
    return prev

for name, h, k in (("8-byte", h8, 8), ("5-byte", h5, 5)):
    a = by_sort(h); b = by_dict(h)
    assert np.array_equal(a, b)
    has = a >= 0
    q = a[has]; p = np.nonzero(has)[0]
    real = np.ones(p.size, dtype=bool)
    for j in range(k):
        real &= frame[p + j] == frame[q + j]
    print(f"{name} hash: {has.mean() * 100:.1f} % of positions have an earlier occurrence; {real.mean() * 100:.2f} % of those are true {k}-byte matches "
          f"(the rest are 32-bit hash collisions); sort == sequential dictionary: True")
