# chain3 tiles hold float4 (16 B) per (bin, lane): a 128-bit shared access is served per quarter-warp
# (8 lanes); conflict-free iff the 8 slots (16-byte units) are distinct mod 8.
import itertools
def worst(slots):
    w = 1
    for qd in range(4):
        s = [x % 8 for x in slots[8*qd:8*qd+8]]
        w = max(w, max(s.count(v) for v in set(s)))
    return w
for LT in (3, 4):
    D = LT + 1
    for RS in range(32, 41):
        rd = worst([RS * (-D * l) + l for l in range(32)])
        best = None
        for perm in itertools.permutations(range(5)):
            # lane bits perm[0..2] -> fillI bits 0..2 ; perm[3..4] -> fillF bits 0..1
            def fI(l): return sum(((l >> perm[i]) & 1) << i for i in range(3))
            def fF(l): return sum(((l >> perm[3 + i]) & 1) << i for i in range(2))
            wf = max(worst([RS * (fI(l) - D * (fF(l) + 4 * it)) + (fF(l) + 4 * it) for l in range(32)]) for it in range(8))
            # global coalescing score: how many lanes of a quarter-warp share a row (more = better)
            rows_per_q = max(len(set(fF(l) for l in range(8 * qd, 8 * qd + 8))) for qd in range(4))
            key = (wf, rows_per_q)
            if best is None or key < best[0]: best = (key, perm)
        print("LT", LT, "RS", RS, "read", rd, "fill(conflict, rows/quarter)", best)
