# bank-conflict check for the chain tiles: float2 tiles [row][RS], 8-byte accesses are served per half-warp,
# a half-warp access is conflict-free iff the 16 float2 slots (addr mod 16) are distinct.
def conflicts(slots):
    worst = 1
    for h in (0, 1):
        s = [x % 16 for x in slots[16*h:16*h+16]]
        worst = max(worst, max(s.count(v) for v in set(s)))
    return worst
for CT in (1, 2):
    for LT in range(1, 9):
        D = LT + 1
        res = []
        for RS in range(32, 41):
            # read in step(): lane -> [ (k - D*(lane//CT)) ][lane]
            rd = conflicts([RS * (-D * (l // CT)) + l for l in range(32)])
            # fill A (old): fillI = lane&7, fillF = lane>>3, rows fl = fillF + 4*it
            fa = max(conflicts([RS * ((l & 7) - D * (((l >> 3) + 4 * it) // CT)) + ((l >> 3) + 4 * it) for l in range(32)]) for it in range(8))
            # fill B: fillI = (lane&3) + 4*(lane>>4), fillF = (lane>>2)&3
            fb = max(conflicts([RS * (((l & 3) + 4 * (l >> 4)) - D * ((((l >> 2) & 3) + 4 * it) // CT)) + (((l >> 2) & 3) + 4 * it) for l in range(32)]) for it in range(8))
            res.append((RS, rd, fa, fb))
        good = [r for r in res if r[1] == 1 and min(r[2], r[3]) == 1]
        print("CT", CT, "LT", LT, "good:", good[:4], "| best otherwise:", sorted(res, key=lambda r: (r[1], min(r[2], r[3])))[:2])
