#!/usr/bin/env python
"""List-scheduling model of the column pipeline: a GPU's range is nrow x ncol block tasks of D steps each,
claimed in a fixed order by `grid` persistent CTAs; a claimed task starts when its top / left producers are hj / hk steps
in (the CTA is blocked meanwhile); column 0 of GPU g waits for the last column of GPU g-1 (+hr steps).  Prints the
finish time of each GPU relative to the ideal (tasks * D / grid) for the claim orders fast_long.cu offers.  The measured
8192^3 numbers it explains: one range alone 67.3 ms row-major / 65.9 ms earliest-start (model 1.046 / 1.026 x 61.6 ms),
8 GPUs 78.3 / 96 ms (model 1.21 / 1.50)."""
import heapq, sys
def simulate(nrow, ncol, order, grid, D, hj, hk, ngpu=1, hr=164):
    """list scheduling: claims in `order` (list of (r,c)), each claim by the earliest-free CTA; a claimed task starts
    when its deps allow (CTA blocked meanwhile). Multi-GPU chain: (r,0) of gpu g needs (r,ncol-1) of gpu g-1 + hr."""
    start = [dict() for _ in range(ngpu)]
    end_all = 0
    ends = []
    for g in range(ngpu):
        free = [0.0] * grid
        heapq.heapify(free)
        for (r, c) in order:
            f = heapq.heappop(free)
            s = f
            if r > 0: s = max(s, start[g][(r - 1, c)] + hj)
            if c > 0: s = max(s, start[g][(r, c - 1)] + hk)
            elif g > 0: s = max(s, start[g - 1][(r, ncol - 1)] + hr)
            start[g][(r, c)] = s
            heapq.heappush(free, s + D)
        ends.append(max(free))
    return ends
nrow, ncol, D, hj, hk = 128, 8, 8238, 27, 52
rm = [(r, c) for r in range(nrow) for c in range(ncol)]
es = sorted(rm, key=lambda t: hj * t[0] + hk * t[1])
ideal = nrow * ncol * D / 148
for name, order in (("row-major", rm), ("earliest-start", es)):
    for grid in (144, 148):
        e1 = simulate(nrow, ncol, order, grid, D, hj, hk)
        e8 = simulate(nrow, ncol, order, grid, D, hj, hk, ngpu=8)
        print(name, grid, "1gpu %.3f" % (e1[0] / ideal), "8gpu last %.3f" % (e8[-1] / ideal), ["%.2f" % (x / ideal) for x in e8])
print("---- tilt sweep (key = hj*r + kappa*c), grid 148")
for kappa in (0, 4, 8, 16, 27, 40, 52):
    order = sorted(rm, key=lambda t: (hj * t[0] + kappa * t[1], t[0], t[1]))
    for hr in (56, 164):
        e8 = simulate(nrow, ncol, order, 148, D, hj, hk, ngpu=8, hr=hr)
        print("kappa", kappa, "hr", hr, "1gpu %.3f" % (e8[0] / ideal), "8gpu %.3f" % (e8[-1] / ideal))
print("---- row partition: 16 rows x 64 cols per gpu, chain via J faces")
def simulate_rows(nrow, ncol, order, grid, D, hj, hk, ngpu, hr):
    start = [dict() for _ in range(ngpu)]
    ends = []
    for g in range(ngpu):
        free = [0.0] * grid
        heapq.heapify(free)
        for (r, c) in order:
            f = heapq.heappop(free)
            s = f
            if c > 0: s = max(s, start[g][(r, c - 1)] + hk)
            if r > 0: s = max(s, start[g][(r - 1, c)] + hj)
            elif g > 0: s = max(s, start[g - 1][(nrow - 1, c)] + hr)
            start[g][(r, c)] = s
            heapq.heappush(free, s + D)
        ends.append(max(free))
    return ends
rr = [(r, c) for c in range(64) for r in range(16)]   # column-major within the row strip
for kappa in (0, 8, 27):
    order = sorted(rr, key=lambda t: (hk * t[1] + kappa * t[0], t[1], t[0]))
    e8 = simulate_rows(16, 64, order, 148, D, hj, hk, 8, 164)
    print("rows kappa", kappa, "1gpu %.3f" % (e8[0] / ideal), "8gpu %.3f" % (e8[-1] / ideal))
