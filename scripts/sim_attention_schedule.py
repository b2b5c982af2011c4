#!/usr/bin/env python
"""Replay the launch order of the attention kernels on the bench's synthetic C2 batches: how many (sequence, 64-row
tile) CTAs are live, how much work (64-key chunks) each has, and how long the kernel must last on 148 SMs x 2 resident
CTAs when a CTA cannot be split — the explanation of the idle SM-time ncu reports.  CPU only."""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_batches  # noqa: E402

B, T, TILE, CH, SLOTS = 128, 200, 64, 64, 296
ntile = (T + TILE - 1) // TILE
for bi, batch in enumerate(synth_batches(4, B, T, 3416, seed=20191019)):
    L = (batch[0] != 0).sum(1)
    tasks = []
    for tile in reversed(range(ntile)):          # grid order: last (heaviest) row tile first
        for Ls in L:
            first = T - int(Ls)
            q0 = T - TILE * (ntile - tile)
            if q0 + TILE <= first:
                continue                          # padding-only tile: exits at once
            kbeg = (first // CH) * CH
            tasks.append((q0 + TILE - kbeg + CH - 1) // CH)
    slots = [0.0] * SLOTS
    heapq.heapify(slots)
    for n in tasks:
        heapq.heappush(slots, heapq.heappop(slots) + n)
    total, mk = sum(tasks), max(slots)
    # the same rows as 32-row tiles, four resident CTAs per SM (a tile's chunk then costs half a unit)
    fine = []
    nt32 = (T + 31) // 32
    for tile in reversed(range(nt32)):
        for Ls in L:
            first = T - int(Ls)
            q0 = T - 32 * (nt32 - tile)
            if q0 + 32 <= first:
                continue
            kbeg = (first // CH) * CH
            fine.append(0.5 * ((q0 + 32 - kbeg + CH - 1) // CH))
    s4 = [0.0] * (2 * SLOTS)
    heapq.heapify(s4)
    for n in fine:
        heapq.heappush(s4, heapq.heappop(s4) + n)
    # (a slot of the 4-CTA configuration runs at half the speed of a slot of the 2-CTA one)
    print(f"batch {bi}: live rows {L.sum() / (B * T):.2f}, mean length {L.mean():.0f}, full-length {np.mean(L == T):.2f}; "
          f"live tiles {len(tasks)} of {B * ntile}, chunk units {total}, per slot {total / SLOTS:.2f}, heaviest {max(tasks)}, "
          f"makespan {mk:.1f} => efficiency {total / SLOTS / mk:.2f}; as 32-row tiles on 4 CTAs/SM: {len(fine)} live, "
          f"makespan {2 * max(s4):.1f} => {sum(fine) / SLOTS / (2 * max(s4)):.2f}")
