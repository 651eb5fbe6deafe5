"""Goal-directed scripted players used to drive trajectories into every reward branch
(onion placement, useful dish pickup, soup pickup, delivery, counter hand-overs) on all
layouts.  Uniform random play almost never cooks a soup on the larger layouts
(SURVEY.md section 8c, KAT-3), so golden trajectories mix this planner with noise.

Works on the packed per-world state rows of include/ocb.h (test utility only).
"""
from __future__ import annotations

from collections import deque

import numpy as np

NONE, TOMATO, ONION, DISH, SOUP = 0, 1, 2, 3, 4
AIR, POT, COUNTER, ONION_SOURCE, DISH_SOURCE, SERVING, TOMATO_SOURCE = range(7)
NORTH, SOUTH, EAST, WEST, STAY, INTERACT = range(6)


class ScriptedTeam:
    def __init__(self, params, rng: np.random.Generator, noise: float = 0.2):
        self.p = params
        self.W, self.H, self.S, self.P = params.width, params.height, params.size, params.num_players
        self.T = list(params.terrain)
        self.times = list(params.recipe_times)
        self.rng = rng
        self.noise = noise
        self.cell0 = 1 + 6 * self.P
        self.delta = (-self.W, self.W, 1, -1)

    def _cells_of(self, kind):
        return [c for c in range(self.S) if self.T[c] == kind]

    def _obj(self, row, c):
        o = self.cell0 + 4 * c
        return row[o], row[o + 1], row[o + 2], row[o + 3]

    def _bfs(self, row, me):
        """distance over AIR cells from player `me`, other players are obstacles"""
        occupied = {row[1 + 6 * j] for j in range(self.P) if j != me}
        src = row[1 + 6 * me]
        dist, first = {src: 0}, {src: None}
        q = deque([src])
        while q:
            c = q.popleft()
            for d in range(4):
                n = c + self.delta[d]
                if 0 <= n < self.S and self.T[n] == AIR and n not in dist and n not in occupied:
                    dist[n] = dist[c] + 1
                    first[n] = d if first[c] is None else first[c]
                    q.append(n)
        return dist, first

    def _go_use(self, row, me, targets):
        """action that walks next to one of `targets` (non-AIR cells), faces it and interacts"""
        if not targets:
            return None
        dist, first = self._bfs(row, me)
        best = None
        for tcell in targets:
            for d in range(4):
                stand = tcell - self.delta[d]  # standing here and facing d looks at tcell
                if stand in dist and (best is None or dist[stand] < best[0]):
                    best = (dist[stand], stand, d)
        if best is None:
            return None
        _, stand, face = best
        pos, ori = row[1 + 6 * me], row[1 + 6 * me + 1]
        if pos == stand:
            return INTERACT if ori == face else face
        return first[stand]

    def act(self, row, me):
        if self.rng.random() < self.noise:
            return int(self.rng.integers(0, 6))
        held = row[1 + 6 * me + 2]
        pots = self._cells_of(POT)
        pot_objs = {c: self._obj(row, c) for c in pots}
        fillable = [c for c, o in pot_objs.items() if o[0] == NONE or (o[3] < 0 and o[1] + o[2] < 3)]
        cooking_or_ready = [c for c, o in pot_objs.items() if o[0] == SOUP and o[3] >= 0]
        ready = [c for c, o in pot_objs.items()
                 if o[0] == SOUP and o[3] >= 0 and o[3] >= self.times[4 * o[1] + o[2]]]
        counters = self._cells_of(COUNTER)
        empty_counters = [c for c in counters if self._obj(row, c)[0] == NONE]
        a = None
        if held == NONE:
            want = []
            others_dish = any(row[1 + 6 * j + 2] == DISH for j in range(self.P) if j != me)
            if cooking_or_ready and not others_dish:
                want.append(self._cells_of(DISH_SOURCE) + [c for c in counters if self._obj(row, c)[0] == DISH])
            if fillable:
                src = self._cells_of(ONION_SOURCE)
                if self.rng.random() < 0.3:
                    src = src + self._cells_of(TOMATO_SOURCE)
                want.append(src + [c for c in counters if self._obj(row, c)[0] in (ONION, TOMATO)])
            want.append([c for c in counters if self._obj(row, c)[0] == SOUP])
            for tg in want:
                a = self._go_use(row, me, tg)
                if a is not None:
                    break
        elif held in (ONION, TOMATO):
            a = self._go_use(row, me, fillable)
            if a is None:
                a = self._go_use(row, me, empty_counters)
        elif held == DISH:
            a = self._go_use(row, me, ready)
            if a is None and cooking_or_ready:
                a = self._go_use(row, me, cooking_or_ready)  # waits in front of the pot
                if a == INTERACT:
                    a = STAY
            if a is None:
                a = self._go_use(row, me, empty_counters)
        elif held == SOUP:
            a = self._go_use(row, me, self._cells_of(SERVING))
            if a is None:
                a = self._go_use(row, me, empty_counters)
        if a is None:
            a = int(self.rng.integers(0, 6))
        return int(a)

    def joint(self, row):
        row = [int(v) for v in row]
        return [self.act(row, i) for i in range(self.P)]
